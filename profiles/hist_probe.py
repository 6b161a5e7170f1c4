"""Kernel-level probe of hist_build_kernel through B2_HistBuildRaw (root pass and a gathered pass)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgboost_ray_b200 import engine as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
f = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = np.random.default_rng(0)
bins = rng.integers(0, 256, size=(n, f), dtype=np.uint8)
qg = rng.integers(-1000, 1000, size=n, dtype=np.int32); qh = rng.integers(0, 1000, size=n, dtype=np.int32)
sel = np.sort(rng.choice(n, n // 2, replace=False)).astype(np.int32)
for name, ridx in (("root", None), ("gather-half-ascending", sel)):
    for rep in range(2):
        _, ms = E.hist_build_raw(bins, qg, qh, ridx=ridx, window_rows=8191, chunk_rows=4096)
    rows = n if ridx is None else len(ridx)
    print("mode=%s %s: %.3f ms  %.2f G rows/s  %.0f GB/s algorithmic" % (os.environ.get("B2_HIST_DEBUG_MODE", "0"), name, ms,
          rows / ms * 1e-6, rows * (f + 8 + (4 if ridx is not None else 0)) / ms * 1e-6), flush=True)
