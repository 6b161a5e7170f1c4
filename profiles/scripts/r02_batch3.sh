#!/bin/bash
# round 2, GPU batch 3 (1 GPU): narrow-group histogram kernel (v3), memory pool, auc, public-API e2e.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b3; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "exit $?" >> $O/smoke.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -8 $O/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "exit $?" >> $O/bench_default.err
i=0
for v in "B2_HIST_VARIANT=2" "B2_HIST_NARROW=0" "B2_POOL_MAX_GB=0"; do
  i=$((i+1)); echo "$v" > $O/bench_ab$i.txt
  timeout 400 env $v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --no-public-e2e >> $O/bench_ab$i.txt 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1450 -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_bench.txt 2>&1
# one boosting round of histogram launches (root + 7 gathered levels), full metric set
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_build -s 8 -c 8 -o $O/hist_full \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_full.txt 2>&1
python profiles/scripts/ncu_hist_summary.py $O/hist_full.ncu-rep $O/hist_traffic.json 10000000 100 > $O/hist_summary.txt 2>&1
ls -la $O | head -30
for f in $O/bench_ab*.txt; do head -1 $f; tail -1 $f | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('seconds_quantise'))"; done
python -c "
import json; d=json.loads(open('$O/bench_default.json').read().strip().split(chr(10))[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['parity'], d['cpu_baseline'])"
