"""Launch list of the quantisation path (sketch + binning) of a C3 matrix: run under
ncu --metrics gpu__time_duration.sum --clock-control none --csv (profiles/r01_summary.md "Quantisation")."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xgboost_ray_b200 import engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
X, y = bench.synth_shard(n, 100, 0, 1)
dm = E.DMatrix(X, label=y)
t0 = time.time(); dm._ensure_quantized(256); print("quantise %.3fs" % (time.time() - t0), flush=True)
