"""How does the CPU oracle scale with OpenMP threads on the GPU box's host? (picks the baseline's thread count)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
X, y = bench.synth_shard(n, 100, 0, 1)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
O.lib().or_set_num_threads(32)
t0 = time.time(); c = O.Cuts.from_data(X); b = c.bin(X); print("quantise(32 thr) %.2fs" % (time.time() - t0), flush=True)
for nt in (8, 16, 32, 64, 128):
    O.lib().or_set_num_threads(nt)
    bst = O.Booster({"objective": "reg:squarederror", "max_depth": 8, "hist_qbits": 0, "base_score": 0.5}, c); bst.init_margin(n)
    bst.boost(b, y)
    t0 = time.time(); bst.boost(b, y); bst.boost(b, y)
    print("threads %3d: %.3f s/round (%d rows)" % (nt, (time.time() - t0) / 2, n), flush=True)
