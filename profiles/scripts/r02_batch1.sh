#!/bin/bash
# round 2, GPU batch 1 (1 GPU): parity suite + public-API tests on the new level loop (CUDA graph, fused leaf pass,
# contiguous histogram chunks), A/B bench lines, ncu launch list.  Outputs under gpurun_out/b1/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b1; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
nvidia-smi -L > $O/gpus.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "exit $?" >> $O/smoke.txt
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > $O/pytest_gpu.txt 2>&1; echo "exit $?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "exit $?" >> $O/bench_default.err
i=0
for v in "B2_GRAPH=0" "B2_LEAF_FUSED=0" "B2_HIST_DEBUG_MODE=4" "B2_GRAPH=0 B2_LEAF_FUSED=0 B2_HIST_DEBUG_MODE=4"; do
  i=$((i+1))
  echo "$v" > $O/bench_ab$i.txt
  timeout 600 env $v python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-parity >> $O/bench_ab$i.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $O/ncu_bench.txt 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider \
  -k "test_trees_identical and not sampling" -x > $O/sanitizer.txt 2>&1; echo "exit $?" >> $O/sanitizer.txt
tail -3 $O/sanitizer.txt
cat $O/bench_default.json | head -c 3000
for f in $O/bench_ab*.txt; do head -1 $f; tail -1 $f | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"; done
