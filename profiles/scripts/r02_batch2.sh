#!/bin/bash
# round 2, GPU batch 2 (2 GPUs): multi-GPU parity through the public API (peer-memory exchange and NCCL), N=2 bench
# lines for both exchanges, N=1 launch list of the training rounds only.  Outputs under gpurun_out/b2/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b2; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
nvidia-smi -L > $O/gpus.txt 2>&1
nvidia-smi topo -m >> $O/gpus.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_public_api.py -q -p no:cacheprovider --timeout 900 > $O/pytest_multi.txt 2>&1; echo "exit $?" >> $O/pytest_multi.txt
tail -15 $O/pytest_multi.txt
run2() {  # $1 = tag, rest = env
  tag=$1; shift
  timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2_$tag.json 2> $O/bench_n2_$tag.err
  echo "exit $?" >> $O/bench_n2_$tag.err
}
run2 p2p B2_DUMMY=1
run2 nccl B2_EXCHANGE=nccl
run2 p2p_nograph B2_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1450 -c 500 --csv --log-file $O/launches_n1.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_bench.txt 2>&1
for t in p2p nccl p2p_nograph; do echo $t; tail -1 $O/bench_n2_$t.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['parity'], d['config'].get('exchange'))" ; tail -3 $O/bench_n2_$t.err; done
