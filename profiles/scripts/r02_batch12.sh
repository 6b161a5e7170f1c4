#!/bin/bash
# round 2, GPU batch 12 (1 GPU, final): kernel v4 with whole-row loads against v3 on the same box, the winner runs the
# full GPU suite and the ncu --set full capture; upload worker count A/B for the e2e arm.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b12; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 400 -k "variants" > $O/pytest_variants.txt 2>&1; echo "exit $?" >> $O/pytest_variants.txt
tail -3 $O/pytest_variants.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_v3.json 2> $O/bench_v3.err; echo "exit $?" >> $O/bench_v3.err
B2_HIST_VARIANT=4 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_v4.json 2> $O/bench_v4.err; echo "exit $?" >> $O/bench_v4.err
B2_UPLOAD_WORKERS=16 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > $O/bench_v3_w16.json 2> $O/bench_v3_w16.err
for t in v3 v4 v3_w16; do tail -1 $O/bench_$t.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4), 'e2e', (d.get('e2e') or {}).get('value'), 'engine-level', ((d.get('e2e') or {}).get('engine_level') or {}).get('value')); print(d.get('parity'))
except Exception as e: print('$t', 'no json', e)"; done
WIN=$(python - <<'PY'
import json
def val(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception: return None
a, b = val('gpurun_out/b12/bench_v3.json'), val('gpurun_out/b12/bench_v4.json')
ok = a and b and (b.get('parity') or {}).get('oracle_match') and (b['parity']['model_sha256'] == a['parity']['model_sha256']) and b['value'] > a['value'] * 1.01
print(4 if ok else 3)
PY
)
echo "winner variant $WIN" | tee $O/winner.txt
export B2_HIST_VARIANT=$WIN
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "exit $?" >> $O/smoke.txt; tail -2 $O/smoke.txt
if [ "$WIN" = "4" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hist_build -s 8 -c 8 -o $O/hist_full \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_full.txt 2>&1
python profiles/scripts/ncu_hist_summary.py $O/hist_full.ncu-rep $O/hist_traffic.json 10000000 100 > $O/hist_summary.txt 2>&1
cat $O/hist_summary.txt
fi
