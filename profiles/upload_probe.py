import os, sys, time, subprocess
import numpy as np
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from xgboost_ray_b200 import engine as E
    X = np.random.default_rng(0).random((10_000_000, 100), dtype=np.float32)
    E.DMatrix(X[:1000])  # context
    for rep in range(3):
        t0 = time.perf_counter(); d = E.DMatrix(X); dt = time.perf_counter() - t0
        print("workers=%s upload %.3f s  %.1f GB/s" % (os.environ.get("B2_UPLOAD_WORKERS"), dt, X.nbytes / dt / 1e9), flush=True)
        del d
else:
    for w in (1, 4, 8, 16):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, B2_UPLOAD_WORKERS=str(w)))
