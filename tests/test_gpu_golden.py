"""CUDA path against the committed golden fixtures (tests/golden/*.json) -- no oracle execution."""
import json
import os

import numpy as np
import pytest

from tests.golden.make_golden import CASES, FEATURE_TYPES, case_data

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", CASES)
def test_engine_reproduces_golden(name):
    from xgboost_ray_b200 import engine as E
    want = json.load(open(os.path.join(GOLD, name + ".json")))
    x, y, w, params, rounds = case_data(name)
    kw = {"feature_types": FEATURE_TYPES[name], "enable_categorical": True} if name in FEATURE_TYPES else {}
    dm = E.DMatrix(x, label=y, weight=w, **kw)
    bst = E.train(params, dm, num_boost_round=rounds, verbose_eval=False)
    ptrs, vals, mins, hm = dm.get_cuts()
    assert [int(v) for v in ptrs] == want["cut_ptrs"]
    assert [int(v) for v in vals.view(np.uint32)] == want["cut_vals_bits"]
    assert [int(v) for v in mins.view(np.uint32)] == want["min_vals_bits"]
    assert [int(v) for v in hm] == want["has_missing"]
    assert [int(v) for v in dm.get_bins().astype(np.int64).sum(axis=0)] == want["bins_sum_per_feature"]
    trees = bst.get_trees()
    assert len(trees) == len(want["trees"])
    for t, g in zip(trees, want["trees"]):
        for k in ("left", "right", "split_feature", "split_bin", "default_left"):
            assert [int(v) for v in t[k]] == g[k], k
        if "split_type" in g:
            assert [int(v) for v in t["split_type"]] == g["split_type"]
            cats = {str(i): [c for c in range(256) if (int(t["cat_bits"][i][c >> 5]) >> (c & 31)) & 1]
                    for i in range(len(t["left"])) if t["split_type"][i]}
            assert cats == g["categories"]
        leaf = np.asarray(g["split_feature"]) < 0
        assert np.max(np.abs(t["value"][leaf] - np.asarray(g["value"], np.float32)[leaf])) <= 1e-5
    pred = np.asarray(bst.predict(E.DMatrix(x[:64], **kw)), np.float64).reshape(-1)
    assert np.max(np.abs(pred - np.asarray(want["pred_head"]))) <= 1e-5
