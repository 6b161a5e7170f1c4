#!/bin/bash
# round 2, GPU batch 8 (2 GPUs): shared-memory atomic microbenchmark (+ ncu counters), NVLink redistribution of INTERLEAVED
# shards (tests + N = 2 bench with the public-API e2e arm, A/B against the strided host read).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b8; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/smem_atomics profiles/microbench/smem_atomics.cu > $O/microbench.txt 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 120 /tmp/smem_atomics 4000 >> $O/microbench.txt 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --clock-control none --metrics smsp__inst_executed_op_shared_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum,gpu__time_duration.sum \
  --csv --log-file $O/microbench_ncu.csv /tmp/smem_atomics 400 > /dev/null 2>&1
cat $O/microbench.txt
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/b8/microbench_ncu.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]),r[ki][:20]),{})[r[mi]]=float(r[vi].replace(',',''))
for k in sorted(d):
    m=d[k]; ins=m.get('smsp__inst_executed_op_shared_atom.sum',0) or 1
    print(k, 'instr %.3g wavefronts/instr %.3f conflicts/instr %.3f us %.1f'%(ins, m.get('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',0)/ins, m.get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum',0)/ins, m.get('gpu__time_duration.sum',0)/1e3))
PY
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_public_api.py -m gpu -q -p no:cacheprovider --timeout 280 > $O/pytest_multi.txt 2>&1; echo "exit $?" >> $O/pytest_multi.txt
tail -6 $O/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "exit $?" >> $O/bench_n2.err
B2_INTERLEAVED_INGEST=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 --no-parity --no-cpu-baseline > $O/bench_n2_hostread.json 2> $O/bench_n2_hostread.err; echo "exit $?" >> $O/bench_n2_hostread.err
for t in n2 n2_hostread; do tail -1 $O/bench_$t.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4)); print(json.dumps(d['e2e'])[:1200]); print(d.get('parity'))
except Exception as e: print('$t', 'no json', e)"; tail -2 $O/bench_$t.err; done
