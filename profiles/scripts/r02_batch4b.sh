#!/bin/bash
# quick 2-GPU check of the single IPC arena and of the strided remote shard read before the 8-GPU batch
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/b4b; mkdir -p $O
export B2_BENCH_CACHE=/tmp/b2cache
timeout 400 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider --timeout 200 -k "exchange_paths or categorical" > $O/pytest_multi.txt 2>&1; echo "exit $?" >> $O/pytest_multi.txt
tail -4 $O/pytest_multi.txt
timeout 200 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider --timeout 200 -k "num_parallel" > $O/pytest_npt.txt 2>&1; tail -2 $O/pytest_npt.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 3 --gen gpu > $O/bench_n2.json 2> $O/bench_n2.err; echo "exit $?" >> $O/bench_n2.err
tail -1 $O/bench_n2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], round(d['roofline']['frac'],4)); print(json.dumps(d['e2e'])[:1500]); print(d.get('parity'))"; tail -2 $O/bench_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:partition_kernel -s 9 -c 2 -o $O/part_full \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --gen gpu > $O/ncu_part.txt 2>&1
ncu -i $O/part_full.ncu-rep --page raw --csv > $O/part_raw.csv 2>/dev/null
ls -la $O
