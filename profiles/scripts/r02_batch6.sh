#!/bin/bash
# round 2, GPU batch 6 (1 GPU): full GPU suite on the final build, C3 headline line, launch list, C2 and C4 (1 GPU).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b6; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "exit $?" >> $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -6 $O/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err; echo "exit $?" >> $O/bench_c3_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1450 -c 300 --csv --log-file $O/launches_n1.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_bench.txt 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 --workload C2 > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err; echo "exit $?" >> $O/bench_c2_n1.err
timeout 900 python bench.py --steps 20 --warmup 3 --workload C4 > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err; echo "exit $?" >> $O/bench_c4_n1.err
for t in c3_n1 c2_n1 c4_n1; do tail -1 $O/bench_$t.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4), d['config'].get('quantise_seconds')); print(json.dumps(d['e2e'])[:900]); print(d.get('parity')); print(d.get('cpu_baseline'))
except Exception as e: print('$t', 'no json', e)"; tail -2 $O/bench_$t.err; done
