"""CPU tests of the oracle: golden fixtures + the reference's relational known-answers."""
import json
import os

import numpy as np
import pytest

from tests.golden.make_golden import CASES, run_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle, name):
    want = json.load(open(os.path.join(GOLD, name + ".json")))
    got = run_case(name)
    for k in ("cut_ptrs", "cut_vals_bits", "min_vals_bits", "has_missing", "bins_sum_per_feature"):
        assert got[k] == want[k], k
    assert len(got["trees"]) == len(want["trees"])
    for a, b in zip(got["trees"], want["trees"]):
        for k in ("left", "right", "split_feature", "split_bin", "default_left"):
            assert a[k] == b[k], k
        assert np.allclose(a["value"], b["value"], rtol=0, atol=1e-7)
        assert a.get("split_type") == b.get("split_type") and a.get("categories") == b.get("categories")
    assert np.allclose(got["pred_head"], want["pred_head"], rtol=0, atol=1e-6)


def test_cut_semantics_binary_feature(oracle):
    """A.2: 0/1 feature -> cuts [1, 2.00001], bins 0/1, min_val -1e-5 (pinned by the toy tests)."""
    x = np.array([[0.0], [1.0], [1.0], [0.0]], np.float32)
    c = oracle.Cuts.from_data(x)
    assert list(c.ptrs) == [0, 2]
    assert c.vals[0] == np.float32(1.0) and abs(c.vals[1] - 2.00001) < 1e-6
    assert abs(c.mins[0] + 1e-5) < 1e-9
    assert list(c.bin(x)[:, 0]) == [0, 1, 1, 0]


def test_cuts_cap_and_missing(oracle):
    rng = np.random.RandomState(0)
    x = rng.normal(size=(5000, 3)).astype(np.float32)
    x[::7, 1] = np.nan
    x[:, 2] = 5.0
    c = oracle.Cuts.from_data(x, 256)
    n = np.diff(c.ptrs)
    assert n[0] == 256 and n[1] == 255 and n[2] == 1      # missing caps a feature at 255 real bins
    assert list(c.has_missing) == [0, 1, 0]
    b = c.bin(x)
    assert (b[::7, 1] == 255).all() and b[:, 0].max() == 255 and (b[:, 2] == 0).all()
    # every cut value is strictly increasing and the last one exceeds the column maximum
    for f in range(3):
        v = c.vals[c.ptrs[f]:c.ptrs[f + 1]]
        assert (np.diff(v) > 0).all() and v[-1] > np.nanmax(x[:, f])
    # quantile bins are balanced
    counts = np.bincount(b[:, 0], minlength=256)
    assert counts.max() <= 3 * counts.mean()


def test_bin_upper_bound_rule(oracle):
    x = np.arange(100, dtype=np.float32).reshape(-1, 1)
    c = oracle.Cuts.from_data(x, 256)
    b = c.bin(np.array([[-5.0], [0.0], [0.5], [1.0], [98.5], [99.0], [1e6]], np.float32))
    assert list(b[:, 0]) == [0, 0, 0, 1, 98, 99, 99]       # bin = #cuts <= x, clamped to the last bin


def test_toy_matrix_known_answers(oracle):
    """xgboost_ray/tests/test_end_to_end.py:72-139."""
    x = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1], [0, 0, 1, 0]] * 8, np.float32)
    y = np.array([0, 1, 2, 3] * 8, np.float32)
    p = {"max_depth": 2, "objective": "multi:softmax", "num_class": 4}
    bst, _ = oracle.train(p, x, y, 2)
    assert list(bst.predict(x)) == list(y)
    t = np.array([[0, 0, 1, 1], [0, 0, 1, 0]], np.float32)
    assert list(oracle.train(p, x[::2], y[::2], 2)[0].predict(t)) == [2, 2]
    assert list(oracle.train(p, x[1::2], y[1::2], 2)[0].predict(t)) == [3, 3]
    # 38 rounds x 4 classes = 152 trees (test_end_to_end.py:238)
    assert oracle.train(p, x, y, 38)[0].num_trees == 152


def test_row_order_independence_is_the_allreduce_property(oracle):
    """Fixed-point histograms make the model independent of row order / sharding: training on any
    permutation (= any INTERLEAVED / BATCH shard concatenation) yields the identical model."""
    rng = np.random.RandomState(3)
    x = rng.uniform(0, 10, size=(4000, 6)).astype(np.float32)
    y = (x[:, 0] * x[:, 1] + rng.normal(size=4000)).astype(np.float32)
    p = {"objective": "reg:squarederror", "max_depth": 5}
    a, _ = oracle.train(p, x, y, 4)
    perm = np.concatenate([np.arange(0, 4000, 2), np.arange(1, 4000, 2)])
    b, _ = oracle.train(p, x[perm], y[perm], 4)
    for i in range(4):
        ta, tb = a.tree(i), b.tree(i)
        assert np.array_equal(ta.split_feature, tb.split_feature) and np.array_equal(ta.split_bin, tb.split_bin)
        assert np.array_equal(ta.value, tb.value)


def test_boost_from_prediction(oracle):
    """4 rounds + 4 rounds from base_margin == 8 rounds (tests/test_sklearn.py:1155-1197)."""
    from sklearn.datasets import load_breast_cancer
    X, y = load_breast_cancer(return_X_y=True)
    X = X.astype(np.float32)
    p = {"objective": "binary:logistic", "max_depth": 6, "eta": 0.3, "base_score": 0.5}
    full, _ = oracle.train(p, X, y, 8)
    first, _ = oracle.train(p, X, y, 4)
    margin = first.predict(X, output_margin=True)
    second, _ = oracle.train(p, X, y, 4, base_margin=margin)
    pred = second.predict(X, base_margin=margin)
    assert np.allclose(pred, full.predict(X), rtol=0, atol=1e-6)


def test_fixed_point_agrees_with_float64(oracle):
    from sklearn.datasets import load_breast_cancer
    X, y = load_breast_cancer(return_X_y=True)
    X = X.astype(np.float32)
    for qb in (14, 18, 22):
        a, _ = oracle.train({"objective": "binary:logistic", "hist_qbits": qb}, X, y, 8)
        b, _ = oracle.train({"objective": "binary:logistic", "hist_qbits": 0}, X, y, 8)
        for i in range(8):
            ta, tb = a.tree(i), b.tree(i)
            assert np.array_equal(ta.split_feature, tb.split_feature) and np.array_equal(ta.split_bin, tb.split_bin)
            leaf = ta.split_feature < 0
            assert np.max(np.abs(ta.value[leaf] - tb.value[leaf])) <= 1e-5   # north_star leaf tolerance


def test_accuracy_ballpark_vs_sklearn(oracle):
    """Independent sanity (not parity): same ballpark as sklearn's HistGradientBoosting on digits 0/1."""
    from sklearn.datasets import load_digits
    from sklearn.ensemble import HistGradientBoostingClassifier
    d = load_digits(n_class=2)
    X, y = d.data.astype(np.float32), d.target.astype(np.float32)
    bst, _ = oracle.train({"objective": "binary:logistic", "max_depth": 4}, X[::2], y[::2], 10)
    err = np.mean((bst.predict(X[1::2]) > 0.5) != y[1::2])
    sk = HistGradientBoostingClassifier(max_iter=10, max_depth=4).fit(X[::2], y[::2])
    assert err < 0.1 and err <= (1 - sk.score(X[1::2], y[1::2])) + 0.05   # test_sklearn.py:115-141 bar


def test_deterministic_exp(oracle):
    xs = np.linspace(-80, 80, 4001).astype(np.float32)
    got = np.array([oracle.lib().or_expf(float(v)) for v in xs], np.float64)
    ref = np.exp(xs.astype(np.float64))
    assert np.max(np.abs(got - ref) / ref) < 4e-7


def test_empty_and_tiny_inputs(oracle):
    x = np.zeros((1, 2), np.float32)
    bst, _ = oracle.train({"objective": "reg:squarederror", "max_depth": 3, "base_score": 0.5}, x, np.ones(1, np.float32), 2)
    assert bst.num_trees == 2 and all(t.n_nodes == 1 for t in bst.trees())
    assert abs(bst.predict(x)[0] - (0.5 + 0.075 + 0.06375)) < 1e-6    # -G/(H+lambda)*eta twice


def test_scale_pos_weight_known_answer(oracle):
    """A.4: positive rows' gradient pairs are multiplied by scale_pos_weight (binary:logistic)."""
    x = np.zeros((10, 1), np.float32)                 # constant feature: the tree is a single leaf
    y = np.array([1, 1, 1, 0, 0, 0, 0, 0, 0, 0], np.float32)
    spw, lam = 3.0, 1.0
    b, _ = oracle.train({"objective": "binary:logistic", "max_depth": 2, "eta": 1.0, "base_score": 0.5, "lambda": lam,
                         "scale_pos_weight": spw}, x, y, 1)
    G = 0.5 * (7 - spw * 3); H = 0.25 * (7 + spw * 3)
    assert b.tree(0).n_nodes == 1 and abs(b.tree(0).value[0] - (-G / (H + lam))) < 1e-6


def test_max_delta_step_known_answer(oracle):
    """A.7: the leaf weight is clipped to +-max_delta_step (before eta); the split gain uses the clipped weights."""
    rng = np.random.RandomState(0)
    x = rng.uniform(0, 1, size=(200, 1)).astype(np.float32)
    y = np.where(x[:, 0] > 0.5, 10.0, -10.0).astype(np.float32)
    p = {"objective": "reg:squarederror", "max_depth": 1, "eta": 0.5, "base_score": 0.0, "lambda": 0.0}
    free, _ = oracle.train(p, x, y, 1)
    clip, _ = oracle.train(dict(p, max_delta_step=2.0), x, y, 1)
    tf, tc = free.tree(0), clip.tree(0)
    assert tf.split_bin[0] == tc.split_bin[0]
    assert sorted(tf.value[1:3]) == [-5.0, 5.0] and sorted(tc.value[1:3]) == [-1.0, 1.0]
    # gain with clipped weights: -(2 G w + H w^2); children: G = -+10 n_c, H = n_c, w = +-2 -> 36 n_c each;
    # root: G = -sum(y), H = 200, w = clip(-G/H, +-2)
    G = -float(y.sum()); w = float(np.clip(-G / 200.0, -2.0, 2.0))
    root_gain = -(2.0 * G * w + 200.0 * w * w)
    assert abs(tc.loss_chg[0] - (36.0 * 200 - root_gain)) < 1e-2


def test_weighted_sketch_known_answers(oracle):
    """A.2 "(weighted) quantiles": sample weights are the rank increments of the sketch.  Integer weights are
    equivalent to repeating rows, a uniform weight changes nothing, invalid weights are rejected."""
    rng = np.random.RandomState(0)
    n = 5000
    X = rng.normal(size=(n, 3)).astype(np.float32)
    X[::11, 1] = np.nan
    base = oracle.Cuts.from_data(X, 16)
    for w in (np.ones(n, np.float32), np.full(n, 0.37, np.float32)):
        c = oracle.Cuts.from_data(X, 16, weight=w)
        assert np.array_equal(c.vals, base.vals) and np.array_equal(c.ptrs, base.ptrs)
    heavy = X[:, 0] > 0
    dup = np.concatenate([X, X[heavy], X[heavy]])
    c_dup = oracle.Cuts.from_data(dup, 16)
    c_w = oracle.Cuts.from_data(X, 16, weight=np.where(heavy, 3.0, 1.0).astype(np.float32))
    assert np.array_equal(c_dup.vals, c_w.vals) and np.array_equal(c_dup.ptrs, c_w.ptrs)
    assert not np.array_equal(c_w.vals, base.vals)
    for bad in (-1.0, np.nan, np.inf):
        w = np.ones(n, np.float32); w[7] = bad
        with pytest.raises(ValueError):
            oracle.Cuts.from_data(X, 16, weight=w)


def _used_features(bst):
    return [sorted({int(f) for f in t.split_feature if f >= 0}) for t in bst.trees()]


def test_sampling_known_answers(oracle):
    """subsample / colsample_* (sampling section of the oracle): fractions, nesting, determinism, weights."""
    rng = np.random.RandomState(3)
    n, F = 4000, 20
    X = rng.uniform(0, 10, size=(n, F)).astype(np.float32)
    y = (X.sum(axis=1) + rng.normal(size=n)).astype(np.float32)
    base = {"objective": "reg:squarederror", "max_depth": 5, "eta": 0.3}
    plain, _ = oracle.train(base, X, y, 3)
    same, _ = oracle.train(dict(base, subsample=1.0, colsample_bytree=1.0, colsample_bylevel=1.0, colsample_bynode=1.0, seed=9), X, y, 3)
    assert all(np.array_equal(a.split_feature, b.split_feature) for a, b in zip(plain.trees(), same.trees()))
    # subsample: the root's hessian sum (= kept rows for squared error) is binomial around subsample * n, per tree
    sub, _ = oracle.train(dict(base, subsample=0.25, seed=1), X, y, 4)
    kept = [t.sum_hess[0] for t in sub.trees()]
    assert all(abs(k - 0.25 * n) < 5 * np.sqrt(n * 0.25 * 0.75) for k in kept) and len(set(kept)) > 1
    # bytree: int(0.5 * 20) = 10 features per tree, different trees draw different sets
    bt, _ = oracle.train(dict(base, colsample_bytree=0.5, seed=1), X, y, 6)
    sets = _used_features(bt)
    assert all(len(s) <= 10 for s in sets) and len({tuple(s) for s in sets}) > 1
    # bynode = 1/F: every node sees exactly one feature; seeds matter, same seed repeats
    bn1, _ = oracle.train(dict(base, colsample_bynode=0.05, seed=1), X, y, 2)
    bn2, _ = oracle.train(dict(base, colsample_bynode=0.05, seed=1), X, y, 2)
    bn3, _ = oracle.train(dict(base, colsample_bynode=0.05, seed=2), X, y, 2)
    assert all(np.array_equal(a.split_feature, b.split_feature) for a, b in zip(bn1.trees(), bn2.trees()))
    assert not all(np.array_equal(a.split_feature, b.split_feature) for a, b in zip(bn1.trees(), bn3.trees()))
    # nesting: with bytree = 0.25 (5 features) every level / node set is a subset of the tree's 5
    nest, _ = oracle.train(dict(base, colsample_bytree=0.25, colsample_bylevel=0.6, colsample_bynode=0.7, seed=4), X, y, 5)
    assert all(len(s) <= 5 for s in _used_features(nest))


def test_feature_weights_known_answer(oracle):
    """test_end_to_end.py:429-467 (xgboost feature_weights demo): weights 0..9, colsample_bynode=0.1 -> feature 0 is
    never used and feature 9 is used most."""
    rng = np.random.RandomState(1994)
    X = rng.randn(1000, 10).astype(np.float32)
    y = rng.randn(1000).astype(np.float32)
    bst, _ = oracle.train({"objective": "reg:squarederror", "colsample_bynode": 0.1, "max_depth": 6}, X, y, 60,
                          feature_weights=np.arange(10, dtype=np.float32))
    cnt = np.zeros(10, int)
    for t in bst.trees():
        for f in t.split_feature:
            if f >= 0:
                cnt[f] += 1
    assert cnt[0] == 0 and cnt.argmax() == 9
    with pytest.raises(ValueError):
        oracle.train({"objective": "reg:squarederror"}, X, y, 1, feature_weights=-np.ones(10, np.float32))


def test_base_score_estimation_known_answers(oracle):
    """A.3 (xgboost >= 2.0, no base_score given): Newton step of a stump at margin 0, then the inverse link:
    weighted mean(y) for squared error, sigmoid(4 (mean(y) - 0.5)) for logistic; 0.5 for multi-class."""
    rng = np.random.RandomState(0)
    X = rng.uniform(0, 1, size=(1000, 3)).astype(np.float32)
    y = (rng.uniform(size=1000) < 0.2).astype(np.float32)
    w = rng.uniform(0.5, 2.0, size=1000).astype(np.float32)
    b, _ = oracle.train({"objective": "binary:logistic", "max_depth": 2}, X, y, 1)
    assert abs(b.params["base_score"] - 1 / (1 + np.exp(-4 * (y.mean() - 0.5)))) < 1e-6
    yr = (X[:, 0] * 3 + 1).astype(np.float32)
    b, _ = oracle.train({"objective": "reg:squarederror", "max_depth": 2}, X, yr, 1, weight=w)
    assert abs(b.params["base_score"] - float((w.astype(np.float64) * yr).sum() / w.sum())) < 1e-5
    b, _ = oracle.train({"objective": "multi:softprob", "num_class": 3, "max_depth": 2}, X, (y * 2), 1)
    assert b.params.get("base_score", 0.5) == 0.5
    b, _ = oracle.train({"objective": "reg:squarederror", "max_depth": 2, "base_score": 0.5}, X, yr, 1)
    assert b.params["base_score"] == 0.5                       # an explicit value is never replaced


def test_custom_objective_known_answer(oracle):
    """test_xgboost_api.py:77-102: ten rounds of the squared-log-error objective on the 4-pattern toy matrix; the
    rounded predictions are the labels."""
    x = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1], [0, 0, 1, 0]] * 8, np.float32)
    y = np.array([0, 1, 0, 1] * 8, np.float32)
    cuts = oracle.Cuts.from_data(x, 256)
    bins = cuts.bin(x)
    b = oracle.Booster({"max_depth": 2, "objective": "reg:squarederror", "seed": 1000}, cuts)
    assert b.estimate_base_score(y) == 0.5
    b.init_margin(len(y))
    for _ in range(10):
        predt = b.margin[:, 0].astype(np.float64)
        predt[predt < -1] = -1 + 1e-6
        g = (np.log1p(predt) - np.log1p(y)) / (predt + 1)
        h = (-np.log1p(predt) + np.log1p(y) + 1) / np.power(predt + 1, 2)
        b.boost(bins, y, custom_g=g, custom_h=h)
    assert list(np.round(b.predict(x))) == list(y)
