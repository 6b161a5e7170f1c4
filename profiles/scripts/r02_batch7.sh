#!/bin/bash
# round 2, GPU batch 7 (1 GPU): 64 KiB-aligned histogram (PRMT-merged cell address) A/B, full GPU suite, ncu --set full of
# one round of histogram launches (-> hist_traffic.json).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b7; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -6 $O/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err; echo "exit $?" >> $O/bench_c3_n1.err
echo "B2_HIST_ALIGNED=0" > $O/bench_ab1.txt
B2_HIST_ALIGNED=0 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --no-e2e >> $O/bench_ab1.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_build -s 8 -c 8 -o $O/hist_full \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_full.txt 2>&1
python profiles/scripts/ncu_hist_summary.py $O/hist_full.ncu-rep $O/hist_traffic.json 10000000 100 > $O/hist_summary.txt 2>&1
for f in $O/bench_c3_n1.json $O/bench_ab1.txt; do tail -1 $f | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4), (d.get('e2e') or {}).get('value'))"; done
cat $O/hist_summary.txt
