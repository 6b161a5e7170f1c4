"""Sharding / data plane tests (port of the hot-path subset of xgboost_ray/tests/test_matrix.py)."""
import os

import numpy as np
import pytest

from xgboost_ray_b200.matrix import (RayDMatrix, RayShardingMode, _get_sharding_indices, combine_data)


def _reference_indices(sharding, rank, num_actors, n):
    """Index sets of xgboost_ray/matrix.py:1088-1110 restated (list form) to pin the slice form."""
    if sharding == RayShardingMode.BATCH:
        per, extras = divmod(n, num_actors)
        div = np.array([0] + extras * [per + 1] + (num_actors - extras) * [per]).cumsum()
        return list(range(div[rank], div[rank + 1]))
    return list(range(rank, n, num_actors))


@pytest.mark.parametrize("n", [1, 2, 7, 32, 33, 1000])
@pytest.mark.parametrize("w", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("mode", [RayShardingMode.INTERLEAVED, RayShardingMode.BATCH])
def test_sharding_indices_match_reference(n, w, mode):
    for r in range(w):
        sl = _get_sharding_indices(mode, r, w, n)
        assert list(range(n))[sl] == _reference_indices(mode, r, w, n)


@pytest.mark.parametrize("n", [8, 9, 31])
@pytest.mark.parametrize("w", [1, 2, 4])
@pytest.mark.parametrize("mode", [RayShardingMode.INTERLEAVED, RayShardingMode.BATCH])
def test_shard_and_combine_round_trip(n, w, mode):
    x = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    y = np.arange(n, dtype=np.float32)
    m = RayDMatrix(x, y, sharding=mode)
    m.load_data(w)
    parts = [m.get_data(r, w) for r in range(w)]
    assert sum(len(p["data"]) for p in parts) == n
    assert np.array_equal(combine_data(mode, [p["label"] for p in parts]), y)        # 1-D predictions
    assert np.array_equal(combine_data(mode, [p["data"] for p in parts]), x)         # softprob-style 2-D
    if mode == RayShardingMode.BATCH and n >= w:
        assert all(len(p["data"]) >= 1 for p in parts)                                # test_matrix.py:406-409


def test_sources_pandas_csv_parquet(tmp_path):
    import pandas as pd
    rng = np.random.RandomState(0)
    df = pd.DataFrame(rng.uniform(size=(20, 4)), columns=["a", "b", "c", "label"])
    csv, pq = str(tmp_path / "d.csv"), str(tmp_path / "d.parquet")
    df.to_csv(csv, index=False)
    df.to_parquet(pq)
    for src in (df, csv, pq):
        m = RayDMatrix(src, label="label")
        m.load_data(2)
        s0, s1 = m.get_data(0, 2), m.get_data(1, 2)
        assert s0["data"].shape == (10, 3) and s0["data"].dtype == np.float32
        assert np.allclose(combine_data(m.sharding, [s0["label"], s1["label"]]), df["label"].values.astype(np.float32))
        assert np.allclose(s0["data"], df[["a", "b", "c"]].values[0::2].astype(np.float32))
        assert m._columns == ["a", "b", "c"]                                          # column order preserved
    # list of files -> distributed (per-actor) loading, FIXED sharding
    df.iloc[:10].to_parquet(str(tmp_path / "p0.parquet"))
    df.iloc[10:].to_parquet(str(tmp_path / "p1.parquet"))
    m = RayDMatrix([str(tmp_path / "p0.parquet"), str(tmp_path / "p1.parquet")], label="label")
    assert m.distributed
    assert np.allclose(m.get_data(1, 2)["label"], df["label"].values[10:].astype(np.float32))
    with pytest.raises(RuntimeError):
        RayDMatrix([str(tmp_path / "p0.parquet"), str(tmp_path / "p1.parquet")], label="label").load_data(3)


def test_errors():
    x = np.zeros((2, 2), np.float32)
    with pytest.raises(RuntimeError):
        RayDMatrix(x, np.zeros(2)).load_data(3)                                       # more actors than rows
    with pytest.raises(ValueError):
        RayDMatrix(x, group=[1, 1])
    m = RayDMatrix(x, np.zeros(2), num_actors=2)
    with pytest.raises(ValueError):
        m.load_data(1)                                                                # actor count is fixed
    with pytest.raises(ValueError):
        RayDMatrix({"not": "supported"})
    a, b = RayDMatrix(x), RayDMatrix(x)
    assert a != b and hash(a) != hash(b) and a == a                                    # identity by uuid


def test_remote_shard_descriptors_carry_the_interleave_layout():
    """transport='remote': a shard is a description of rows inside THIS process.  INTERLEAVED shards also say where the
    whole matrix lives (address, rows, shard rank, number of shards) so that W GPU actors can each read one contiguous
    1/W block and exchange rows on the device (B2_MatrixCreateFromProcessInterleaved); BATCH shards are contiguous and
    carry no such hint.  Reading a descriptor back (same process here) reproduces the shard."""
    import os
    from xgboost_ray_b200 import RayDMatrix, RayShardingMode
    from xgboost_ray_b200.main import _attach_shared
    rng = np.random.RandomState(0)
    x = rng.normal(size=(1003, 7)).astype(np.float32)
    y = rng.normal(size=1003).astype(np.float32)
    for mode in (RayShardingMode.INTERLEAVED, RayShardingMode.BATCH):
        d = RayDMatrix(x, y, sharding=mode)
        d.load_data(3, transport="remote")
        for r in range(3):
            desc = d.get_shared(r, 3)["data"]
            assert desc[0] == "remote" and desc[1] == os.getpid() and len(desc) == 7
            block = _attach_shared(desc)
            want = d.get_data(r, 3)["data"]
            assert block.shape == want.shape and np.array_equal(np.asarray(block), want)
            if mode == RayShardingMode.INTERLEAVED:
                addr, n_total, rank, world = block.interleave
                assert (n_total, rank, world) == (1003, r, 3) and addr == d._keep_alive.ctypes.data
                assert block.row_stride == 3 * 7 * 4 and np.array_equal(want, x[r::3])
            else:
                assert block.interleave is None and block.row_stride == 7 * 4
        d.unload_data()
