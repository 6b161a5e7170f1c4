#!/bin/bash
# round 2, GPU batch 5 (8 GPUs): C3 at N=8 and N=4 (peer-memory exchange over NVSwitch, public-API e2e with 8 actors),
# C5 (50M x 200, 50 categorical, 10 classes) at N=8.  Data generated on the GPUs (--gen gpu).  Outputs gpurun_out/b5/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b5; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1; free -g >> $O/gpus.txt; nproc >> $O/gpus.txt; df -h /dev/shm >> $O/gpus.txt
run() {  # $1 tag, $2 ngpu, rest: bench args
  tag=$1; n=$2; shift; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $n --steps 20 --warmup 3 "$@" > $O/bench_$tag.json 2> $O/bench_$tag.err
  echo "exit $?" >> $O/bench_$tag.err
  tail -1 $O/bench_$tag.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$tag', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4), d['config'].get('quantise_seconds')); print(json.dumps(d['e2e'])[:1200]); print(d.get('parity'))
except Exception as e: print('$tag', 'no json', e)"
  tail -2 $O/bench_$tag.err
}
run c3_n8 8 --gen gpu
run c5_n8 8 --workload C5
