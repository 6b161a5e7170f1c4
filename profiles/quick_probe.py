"""Early probe (round 1): quantise + a few boosting rounds on a C3-shaped synthetic matrix, print timers."""
import json, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xgboost_ray_b200 import engine as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
f = int(sys.argv[2]) if len(sys.argv) > 2 else 100
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 6
qbits = int(sys.argv[5]) if len(sys.argv) > 5 else 18
rng = np.random.default_rng(1234)
t0 = time.time()
X = rng.random((n, f), dtype=np.float32) * 10
a = rng.normal(size=10).astype(np.float32)
y = (X[:, :10] @ a + np.sin(X[:, 10]) + rng.normal(scale=0.1, size=n).astype(np.float32)).astype(np.float32)
print("gen %.1fs" % (time.time() - t0), flush=True)
t0 = time.time(); dm = E.DMatrix(X, label=y); print("upload %.2fs" % (time.time() - t0), flush=True)
t0 = time.time(); dm._ensure_quantized(256); print("quantize %.2fs" % (time.time() - t0), flush=True)
params = {"objective": "reg:squarederror", "max_depth": depth, "eta": 0.3, "base_score": 0.5, "hist_qbits": qbits}
bst = E.Booster(params, cache=[dm])
for r in range(rounds):
    t0 = time.time(); bst.update(dm, r); dt = time.time() - t0
    t = bst.get_timers(reset=True)
    gbs = t["hist_bytes"] / (t["hist_ms"] * 1e-3) / 1e9 if t["hist_ms"] else 0
    print("round %d wall %.1f ms dev %.1f ms hist %.2f ms (%d launches, %.2f GB, %.0f GB/s) launches %d rmse %s" % (
        r, dt * 1e3, t["round_ms"], t["hist_ms"], t["hist_launches"], t["hist_bytes"] / 1e9, gbs, t["kernel_launches"],
        bst.eval_set([(dm, "train")], r)), flush=True)
print([len(t["left"]) for t in bst.get_trees()])
