"""hist kernel probe: chunk / window sweep (values are small, so huge windows are exact)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgboost_ray_b200 import engine as E
n, f = 10_000_000, 100
rng = np.random.default_rng(0)
bins = rng.integers(0, 256, size=(n, f), dtype=np.uint8)
qg = rng.integers(-1000, 1000, size=n, dtype=np.int32); qh = rng.integers(0, 1000, size=n, dtype=np.int32)
for window, chunk in ((8191, 4096), (8191, 8191), (1 << 20, 4096), (1 << 20, 16384), (1 << 20, 65536), (32767, 32767)):
    for rep in range(2):
        _, ms = E.hist_build_raw(bins, qg, qh, window_rows=window, chunk_rows=chunk)
    print("mode=%s window=%d chunk=%d root: %.3f ms %.2f G rows/s" % (os.environ.get("B2_HIST_DEBUG_MODE", "0"), window, chunk, ms, n / ms * 1e-6), flush=True)
