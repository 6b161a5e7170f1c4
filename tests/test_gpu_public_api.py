"""The public surface -- xgboost_ray_b200.train / predict with RayDMatrix and RayParams -- on ONE GPU (these run on the
driver's single-GPU box): one spawned actor process, the real sm_100a engine underneath, results against the oracle.
Ports the hot-path subset of xgboost_ray/tests/test_end_to_end.py, test_fault_tolerance.py:401-444 and the actor
stop path of xgboost_ray/main.py:628-652, 774-785."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(seed=3, n=20011, f=17):
    rng = np.random.RandomState(seed)
    x = rng.uniform(0, 10, size=(n, f)).astype(np.float32)
    x[rng.uniform(size=x.shape) < 0.04] = np.nan
    y = (np.nan_to_num(x[:, 0]) + 0.5 * np.nan_to_num(x[:, 3]) + rng.normal(size=n) > 7).astype(np.float32)
    return x, y


def _same_trees(bst, ob, n_trees):
    trees = bst.get_trees()
    assert len(trees) == ob.num_trees == n_trees
    for i, t in enumerate(trees):
        o = ob.tree(i)
        assert np.array_equal(t["split_feature"], o.split_feature), "tree %d" % i
        assert np.array_equal(t["split_bin"], o.split_bin), "tree %d" % i
        assert np.array_equal(t["default_left"], o.default_left), "tree %d" % i
        leaf = o.split_feature < 0
        assert np.max(np.abs(t["value"][leaf] - o.value[leaf])) <= 1e-5


@pytest.mark.timeout(600)
@pytest.mark.parametrize("sharding", ["INTERLEAVED", "BATCH"])
def test_train_predict_one_actor_matches_oracle(oracle, sharding):
    from xgboost_ray_b200 import RayDMatrix, RayParams, RayShardingMode, predict, train
    x, y = _data()
    params = {"objective": "binary:logistic", "tree_method": "hist", "max_depth": 6, "eta": 0.3, "base_score": 0.5,
              "eval_metric": ["logloss", "error"]}
    mode = getattr(RayShardingMode, sharding)
    res, extra = {}, {}
    d = RayDMatrix(x, y, sharding=mode)
    bst = train(params, d, num_boost_round=6, evals=[(d, "train")], evals_result=res, additional_results=extra,
                ray_params=RayParams(num_actors=1, checkpoint_frequency=2))
    ob, _ = oracle.train(params, x, y, 6)
    _same_trees(bst, ob, 6)
    assert extra["total_n"] == len(y) and extra["training_time_s"] > 0
    assert len(res["train"]["logloss"]) == 6
    assert abs(res["train"]["logloss"][-1] - ob.metric("logloss", ob.margin, y)) < 1e-6
    assert abs(res["train"]["error"][-1] - ob.metric("error", ob.margin, y)) < 1e-6   # the eval line carries 6 decimals
    p = predict(bst, RayDMatrix(x, sharding=mode), ray_params=RayParams(num_actors=1))
    assert p.shape == (len(y),)
    assert np.max(np.abs(p - ob.predict(x))) <= 1e-5
    m = predict(bst, RayDMatrix(x, sharding=mode), ray_params=RayParams(num_actors=1), output_margin=True)
    assert np.max(np.abs(m - ob.predict(x, output_margin=True))) <= 1e-5


@pytest.mark.timeout(600)
def test_multiclass_toy_matrix_through_public_api():
    """test_end_to_end.py:72-103, 238-254: 4-class toy matrix is learned exactly; softprob predictions are [n, K]."""
    from xgboost_ray_b200 import RayDMatrix, RayParams, predict, train
    x = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1], [0, 0, 1, 0]] * 8, np.float32)
    y = np.array([0, 1, 2, 3] * 8, np.float32)
    bst = train({"max_depth": 2, "objective": "multi:softmax", "num_class": 4}, RayDMatrix(x, y), num_boost_round=2,
                ray_params=RayParams(num_actors=1))
    assert bst.num_trees() == 8
    assert list(predict(bst, RayDMatrix(x), ray_params=RayParams(num_actors=1))) == list(y)
    bst = train({"max_depth": 2, "objective": "multi:softprob", "num_class": 4}, RayDMatrix(x, y), num_boost_round=2,
                ray_params=RayParams(num_actors=1))
    p = predict(bst, RayDMatrix(x), ray_params=RayParams(num_actors=1))
    assert p.shape == (32, 4) and list(np.argmax(p, axis=1)) == list(y)


@pytest.mark.timeout(420)
def test_actor_killed_restart_from_checkpoint_equals_uninterrupted(oracle, tmp_path):
    """test_fault_tolerance.py:401-444 on the GPU: kill -9 the actor at round 7, the driver restarts it and training
    continues from checkpoint 5; the trees equal those of an uninterrupted run and of the oracle.  No base_score is
    given: the restart must keep the intercept that was estimated from the labels in the first attempt."""
    from tests.fault_injection import DieOnceCallback
    from xgboost_ray_b200 import RayDMatrix, RayParams, train
    x, y = _data(seed=11, n=12007, f=9)
    params = {"objective": "binary:logistic", "max_depth": 4, "eta": 0.3}
    ref = train(params, RayDMatrix(x, y), num_boost_round=10, ray_params=RayParams(num_actors=1, checkpoint_frequency=5))
    lock = str(tmp_path / "lock")
    bst = train(params, RayDMatrix(x, y), num_boost_round=10,
                ray_params=RayParams(num_actors=1, max_actor_restarts=1, checkpoint_frequency=5),
                callbacks=[DieOnceCallback(lock, rank=0, at=7)])
    assert os.path.exists(lock)
    assert bst.num_boosted_rounds() == 10
    assert bst.get_dump(dump_format="json", with_stats=True) == ref.get_dump(dump_format="json", with_stats=True)
    assert abs(float(bst.params["base_score"]) - float(ref.params["base_score"])) == 0.0
    ob, _ = oracle.train(params, x, y, 10)
    _same_trees(bst, ob, 10)
    with pytest.raises(RuntimeError, match="maximum number of retries"):
        train(params, RayDMatrix(x, y), num_boost_round=10, ray_params=RayParams(num_actors=1, max_actor_restarts=0),
              callbacks=[DieOnceCallback(str(tmp_path / "lock2"), rank=0, at=2)])


@pytest.mark.timeout(600)
def test_stop_event_interrupts_actor_training():
    """main.py:628-652, 774-785: setting the stop event ends the actor's train() call with RayXGBoostTrainingStopped
    (the engine stops at the next round boundary; the communicator is aborted first when there is one)."""
    import multiprocessing as mp
    from xgboost_ray_b200 import RayDMatrix, RayParams
    from xgboost_ray_b200 import main as M
    x, y = _data(seed=5, n=50000, f=20)
    stop_event = M._StopFlag(mp.get_context("spawn"))
    rp = M._validate_ray_params(RayParams(num_actors=1))
    _, rp.gpus_per_actor = M._autodetect_resources(rp)
    actors = M._create_actors(rp, stop_event)
    try:
        d = RayDMatrix(x, y)
        d.load_data(1)
        actors[0].call("load_data", d._uid, d.get_shared(0, 1), M._matrix_meta(d)).result()
        fut = actors[0].call("train", {"b2_uid": b"", "b2_rank": 0, "b2_world": 1}, True,
                             {"objective": "binary:logistic", "max_depth": 6}, d._uid, [], 1_000_000)
        time.sleep(3.0)
        assert not fut.done(0.0)                       # still training
        stop_event.set()
        t0 = time.time()
        with pytest.raises(M.RayXGBoostTrainingStopped):
            fut.result(timeout=60)
        assert time.time() - t0 < 30
        assert actors[0].is_alive()                    # the actor survives a stop and can be reused
    finally:
        M._shutdown(actors, force=True)


@pytest.mark.timeout(600)
def test_eval_matrix_freed_and_reallocated_is_not_served_from_a_stale_cache():
    """The evaluation-margin cache of a Booster is keyed by a matrix id, not by its address."""
    from xgboost_ray_b200 import engine as E
    x, y = _data(seed=7, n=4000, f=6)
    bst = E.train({"objective": "binary:logistic", "max_depth": 3, "base_score": 0.5}, E.DMatrix(x, label=y), 3, verbose_eval=False)
    seen = []
    for i in range(6):
        xi, yi = _data(seed=100 + i, n=1500, f=6)
        dm = E.DMatrix(xi, label=yi)
        got = float(bst.eval(dm).split(":")[-1])
        p = np.clip(bst.predict(E.DMatrix(xi)), 1e-16, 1 - 1e-16)
        want = float(np.mean(-(yi * np.log(p) + (1 - yi) * np.log(1 - p))))
        assert abs(got - want) < 1e-4, (i, got, want)
        seen.append(got)
        del dm
    assert len(set(round(v, 6) for v in seen)) > 1


@pytest.mark.timeout(420)
def test_sklearn_random_forest_estimators_on_the_gpu():
    """xgboost_ray/sklearn.py:602-637, 880-914 (RayXGBRFRegressor / RayXGBRFClassifier): one round of n_estimators
    parallel trees through the public fit / predict surface."""
    from sklearn.datasets import load_digits
    from xgboost_ray_b200 import RayParams
    from xgboost_ray_b200.sklearn import RayXGBRFClassifier, RayXGBRFRegressor
    rng = np.random.RandomState(0)
    X = rng.uniform(0, 10, size=(4000, 6)).astype(np.float32)
    y = (X[:, 0] * 2 + X[:, 1] + rng.normal(scale=0.2, size=4000)).astype(np.float32)
    rf = RayXGBRFRegressor(n_estimators=16, max_depth=6, random_state=5).fit(X, y, ray_params=RayParams(num_actors=1))
    assert len(rf.get_booster().get_dump()) == 16 and rf.get_booster().num_boosted_rounds() == 1
    assert np.mean((rf.predict(X, ray_params=RayParams(num_actors=1)) - y) ** 2) < 0.2 * np.var(y)
    d = load_digits(n_class=2)
    clf = RayXGBRFClassifier(n_estimators=8, max_depth=4, random_state=1).fit(d.data.astype(np.float32)[::2], d.target[::2],
                                                                            ray_params=RayParams(num_actors=1))
    pred = clf.predict(d.data.astype(np.float32)[1::2], ray_params=RayParams(num_actors=1))
    assert np.mean(pred != d.target[1::2]) < 0.1
