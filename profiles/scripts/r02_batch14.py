"""round 2, GPU batch 14 (2 GPUs): where does a public train() call at N = 2 spend its time?  Three consecutive calls on
the same 10M x 100 host matrix (strided host read of the INTERLEAVED shards), wall-clock stamps of both actors."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from xgboost_ray_b200 import RayDMatrix, RayParams, train  # noqa: E402
import xgboost_ray_b200.main as M  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    n, f = 10_000_000, 100
    X = rng.random((n, f), dtype=np.float32)
    y = (X[:, 0] * 2 + X[:, 1] - X[:, 2] + rng.random(n, dtype=np.float32)).astype(np.float32)
    params = {"objective": "reg:squarederror", "max_depth": 8, "eta": 0.1, "base_score": 0.5, "max_bin": 256}
    train(params, RayDMatrix(X[:200000], y[:200000]), num_boost_round=3, ray_params=RayParams(num_actors=2))
    out = []
    for call in range(3):
        t0 = time.time()
        d = RayDMatrix(X, y)
        extra = {}
        train(params, d, num_boost_round=20, evals=[(d, "train")], additional_results=extra, verbose_eval=False,
              ray_params=RayParams(num_actors=2))
        t1 = time.time()
        tm = extra.get("timing", {})
        rel = lambda v: None if v is None else round(v - t0, 4)   # noqa: E731
        row = {"call": call, "wall": round(t1 - t0, 4), "dispatch": rel(tm.get("t_dispatch")),
               "futures_done": [rel(v) for v in tm.get("t_future_done", [])],
               "actors": [{k: (rel(v) if k.startswith("t_") else (round(v, 4) if isinstance(v, float) else v)) for k, v in (a or {}).items()}
                          for a in tm.get("actors", [])]}
        out.append(row)
        print(json.dumps(row), flush=True)
        del d
    M.shutdown_actors()
    json.dump(out, open("gpurun_out/b14/calls.json", "w"), indent=1)


if __name__ == "__main__":
    main()
