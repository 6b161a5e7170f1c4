"""Wall time of upload + quantisation of a C3 matrix, tiled vs element-wise bin kernel (B2_BIN_TILED=0/1)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xgboost_ray_b200 import engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
X, y = bench.synth_shard(n, 100, 0, 1)
for rep in range(3):
    dm = E.DMatrix(X, label=y)
    t0 = time.time(); dm._ensure_quantized(256); dt = time.time() - t0
    print("B2_BIN_TILED=%s quantise %.4fs" % (os.environ.get("B2_BIN_TILED", "1"), dt), flush=True)
    del dm
