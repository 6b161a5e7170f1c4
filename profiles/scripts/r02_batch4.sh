#!/bin/bash
# round 2, GPU batch 4 (2 GPUs): multi-GPU + public-API parity on the final host layer, N=1 / N=2 bench lines with the
# public-API e2e arm and its timing breakdown, N=1 launch list.  Outputs under gpurun_out/b4/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b4; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_public_api.py -q -p no:cacheprovider --timeout 240 > $O/pytest_multi.txt 2>&1; echo "exit $?" >> $O/pytest_multi.txt
tail -6 $O/pytest_multi.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider --timeout 240 -k "hist_kernel or auc or num_parallel or trees_identical or base_score" > $O/pytest_parity.txt 2>&1; echo "exit $?" >> $O/pytest_parity.txt
tail -4 $O/pytest_parity.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "exit $?" >> $O/bench_n1.err
CUDA_VISIBLE_DEVICES=0 B2_SHARD_TRANSPORT=shm timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > $O/bench_n1_shm.json 2> $O/bench_n1_shm.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "exit $?" >> $O/bench_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1450 -c 300 --csv --log-file $O/launches_n1.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_bench.txt 2>&1
for t in n1 n1_shm n2; do echo $t; tail -1 $O/bench_$t.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4)); print(json.dumps(d['e2e'])[:1500]); print(d.get('parity'))" ; tail -2 $O/bench_$t.err; done
