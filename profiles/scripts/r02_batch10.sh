#!/bin/bash
# round 2, GPU batch 10 (1 GPU): histogram kernel v4 (all groups of a row in one CTA): bit-exactness of every variant,
# bench A/B against v3 (same box), the full GPU suite with v4 active, ncu --set full of one round of v4 launches.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b10; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 400 -k "variants or hist_kernel" > $O/pytest_variants.txt 2>&1; echo "exit $?" >> $O/pytest_variants.txt
tail -4 $O/pytest_variants.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_v3.json 2> $O/bench_v3.err; echo "exit $?" >> $O/bench_v3.err
B2_HIST_VARIANT=4 timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_v4.json 2> $O/bench_v4.err; echo "exit $?" >> $O/bench_v4.err
for t in v3 v4; do tail -1 $O/bench_$t.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4), (d.get('e2e') or {}).get('value')); print(d.get('parity'))
except Exception as e: print('$t', 'no json', e)"; tail -2 $O/bench_$t.err; done
B2_HIST_VARIANT=4 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 > $O/pytest_gpu_v4.txt 2>&1; echo "exit $?" >> $O/pytest_gpu_v4.txt
tail -6 $O/pytest_gpu_v4.txt
B2_HIST_VARIANT=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_build -s 8 -c 8 -o $O/hist_full_v4 \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_full.txt 2>&1
python profiles/scripts/ncu_hist_summary.py $O/hist_full_v4.ncu-rep $O/hist_traffic_v4.json 10000000 100 > $O/hist_summary.txt 2>&1
cat $O/hist_summary.txt
