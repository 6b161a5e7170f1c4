"""bench.py's synthetic workload: deterministic, and the rank shards tile the global matrix exactly
(INTERLEAVED sharding, xgboost_ray/matrix.py:1100), so the bench trains the SAME dataset at every N."""
import numpy as np

import bench


def test_shards_tile_the_global_matrix():
    n, f = 2500, 12
    X, y = bench.synth_shard(n, f, 0, 1, block_rows=1000)
    X2, y2 = bench.synth_shard(n, f, 0, 1, block_rows=1000)
    assert np.array_equal(X, X2) and np.array_equal(y, y2)           # deterministic
    assert X.shape == (n, f) and X.dtype == np.float32 and 0 <= X.min() and X.max() < 10
    for world in (2, 3, 8):
        parts = [bench.synth_shard(n, f, r, world, block_rows=1000) for r in range(world)]
        for r, (xs, ys) in enumerate(parts):
            assert np.array_equal(xs, X[r::world]) and np.array_equal(ys, y[r::world])
        assert sum(len(p[0]) for p in parts) == n


def test_label_is_learnable_signal():
    X, y = bench.synth_shard(4000, 20, 0, 1)
    a = np.random.default_rng(1234).normal(size=10).astype(np.float32)
    resid = y - (X[:, :10] @ a + np.sin(X[:, 10]))
    assert abs(resid.mean()) < 0.02 and 0.05 < resid.std() < 0.2     # y = sum a_j x_j + sin(x_10) + N(0, 0.1)
