#!/bin/bash
# round 2, GPU batch 11 (2 GPUs): NVLink shard redistribution with a persistent staging block and cached peer mappings:
# phase timing, the ingest tests, the N = 2 bench line with the public-API e2e arm.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b11; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider --timeout 280 -k "interleaved or identical_to_one_gpu_and_oracle or killed" > $O/pytest_multi.txt 2>&1; echo "exit $?" >> $O/pytest_multi.txt
tail -4 $O/pytest_multi.txt
B2_INGEST_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err; echo "exit $?" >> $O/bench_n2.err
grep "b2 ingest rank 0" $O/bench_n2.err | tail -24
for t in n2; do tail -1 $O/bench_$t.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4)); print(json.dumps(d['e2e'])[:1500]); print(d.get('parity'))
except Exception as e: print('$t', 'no json', e)"; tail -2 $O/bench_$t.err; done
