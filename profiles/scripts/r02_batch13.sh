#!/bin/bash
# round 2, GPU batch 13 (2 GPUs): public-API e2e at N = 2 with warm memory pools, strided host read (default) against the
# NVLink redistribution of INTERLEAVED shards (B2_INTERLEAVED_INGEST=1).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b13; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-parity --no-cpu-baseline > $O/bench_n2_host.json 2> $O/bench_n2_host.err
B2_INTERLEAVED_INGEST=1 timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 --no-parity --no-cpu-baseline > $O/bench_n2_nvlink.json 2> $O/bench_n2_nvlink.err
for t in n2_host n2_nvlink; do tail -1 $O/bench_$t.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step']); print(json.dumps(d['e2e'])[:1300])
except Exception as e: print('$t', 'no json', e)"; done
