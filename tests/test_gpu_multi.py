"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2`): the public train()/predict() with
one actor process per GPU and the NCCL histogram allreduce produce the SAME model as one GPU and as
the oracle (integer histograms make the model independent of the world size)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    from xgboost_ray_b200 import engine
    return engine.device_count()


def _dump(bst):
    return bst.get_dump(dump_format="json", with_stats=True)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("sharding", ["INTERLEAVED", "BATCH"])
def test_two_gpu_model_identical_to_one_gpu_and_oracle(oracle, sharding):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from xgboost_ray_b200 import RayDMatrix, RayParams, RayShardingMode, predict, train
    rng = np.random.RandomState(5)
    n, f = 30001, 20   # odd row count: shards of different size
    x = rng.uniform(0, 10, size=(n, f)).astype(np.float32)
    x[rng.uniform(size=x.shape) < 0.05] = np.nan
    y = (np.nan_to_num(x[:, 0]) + np.nan_to_num(x[:, 1]) * 0.5 + rng.normal(size=n) > 7).astype(np.float32)
    params = {"objective": "binary:logistic", "max_depth": 6, "eta": 0.3, "base_score": 0.5, "eval_metric": ["logloss", "error"]}
    mode = getattr(RayShardingMode, sharding)
    res1, res2 = {}, {}
    d1 = RayDMatrix(x, y, sharding=mode)
    b1 = train(params, d1, num_boost_round=6, evals=[(d1, "train")], evals_result=res1, ray_params=RayParams(num_actors=1))
    d2 = RayDMatrix(x, y, sharding=mode)
    b2 = train(params, d2, num_boost_round=6, evals=[(d2, "train")], evals_result=res2, ray_params=RayParams(num_actors=2))
    assert _dump(b1) == _dump(b2)                                   # byte-identical trees
    assert np.allclose(res1["train"]["logloss"], res2["train"]["logloss"], rtol=0, atol=1e-9)
    ob, _ = oracle.train(params, x, y, 6)
    for i, t in enumerate(b2.get_trees()):
        o = ob.tree(i)
        assert np.array_equal(t["split_feature"], o.split_feature) and np.array_equal(t["split_bin"], o.split_bin)
        assert np.array_equal(t["default_left"], o.default_left)
        leaf = o.split_feature < 0
        assert np.max(np.abs(t["value"][leaf] - o.value[leaf])) <= 1e-5
    p2 = predict(b2, RayDMatrix(x, sharding=mode), ray_params=RayParams(num_actors=2))
    assert np.max(np.abs(p2 - ob.predict(x))) <= 1e-5               # recombined in original row order


@pytest.mark.timeout(300)
def test_two_gpu_toy_matrix(oracle):
    """test_end_to_end.py:162-211 on real GPUs: halves over-fit alone, two actors are exact."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from xgboost_ray_b200 import RayDMatrix, RayParams, predict, train
    x = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1], [0, 0, 1, 0]] * 8, np.float32)
    y = np.array([0, 1, 2, 3] * 8, np.float32)
    params = {"max_depth": 2, "objective": "multi:softmax", "num_class": 4}
    bst = train(params, RayDMatrix(x, y), num_boost_round=2, ray_params=RayParams(num_actors=2))
    assert list(predict(bst, RayDMatrix(x), ray_params=RayParams(num_actors=2))) == list(y)
    assert bst.num_trees() == 8


@pytest.mark.timeout(300)
def test_config_c1_breast_cancer_two_actors(oracle):
    """BASELINE config C1: breast_cancer, binary:logistic, tree_method=hist, RayParams(num_actors=2) -- here on
    two GPU actors; the model must equal the oracle's and the committed golden fixture's split sequence."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import os
    from sklearn.datasets import load_breast_cancer
    from xgboost_ray_b200 import RayDMatrix, RayParams, predict, train
    X, y = load_breast_cancer(return_X_y=True)
    X = X.astype(np.float32)
    params = {"objective": "binary:logistic", "tree_method": "hist", "max_depth": 6, "eta": 0.3, "base_score": 0.5}
    res = {}
    d = RayDMatrix(X, y.astype(np.float32))
    bst = train(params, d, num_boost_round=5, evals=[(d, "train")], evals_result=res,
                ray_params=RayParams(num_actors=2, cpus_per_actor=1))
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "breast_cancer_logistic.json")))
    for t, g in zip(bst.get_trees(), gold["trees"]):
        assert [int(v) for v in t["split_feature"]] == g["split_feature"]
        assert [int(v) for v in t["split_bin"]] == g["split_bin"]
        leaf = np.asarray(g["split_feature"]) < 0
        assert np.max(np.abs(t["value"][leaf] - np.asarray(g["value"], np.float32)[leaf])) <= 1e-5
    p = predict(bst, RayDMatrix(X), ray_params=RayParams(num_actors=2))
    assert np.mean((p > 0.5) == (y > 0.5)) > 0.98
    assert res["train"]["logloss"][-1] < res["train"]["logloss"][0]


@pytest.mark.timeout(300)
def test_two_gpu_categorical_identical_to_one_gpu_and_oracle(oracle):
    """Categorical splits (config C5's feature kind): the category-set candidates travel through the candidate
    allgather; 1 GPU, 2 GPUs and the oracle agree."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from xgboost_ray_b200 import RayDMatrix, RayParams, predict, train
    rng = np.random.RandomState(9)
    n = 40001
    xn = rng.uniform(0, 10, size=(n, 30))
    cats = np.column_stack([rng.randint(0, c, size=n) for c in (3, 4, 16, 64, 250)])
    x = np.column_stack([xn, cats]).astype(np.float32)
    x[rng.uniform(size=x.shape) < 0.03] = np.nan
    y = ((np.nan_to_num(x[:, 0]) > 5) * 1.0 + (np.nan_to_num(x[:, 32]) % 3 == 0) * 2.0 + (np.nan_to_num(x[:, 34]) % 5 < 2) * 1.0
         + rng.normal(scale=0.2, size=n)).astype(np.float32)
    types = ["q"] * 30 + ["c"] * 5
    params = {"objective": "reg:squarederror", "max_depth": 6, "eta": 0.3, "base_score": 0.5}
    kw = dict(feature_types=types, enable_categorical=True)
    b1 = train(params, RayDMatrix(x, y, **kw), num_boost_round=4, ray_params=RayParams(num_actors=1))
    b2 = train(params, RayDMatrix(x, y, **kw), num_boost_round=4, ray_params=RayParams(num_actors=2))
    assert _dump(b1) == _dump(b2)
    ob, _ = oracle.train(params, x, y, 4, is_cat=[0] * 30 + [1] * 5)
    n_cat_nodes = 0
    for i, t in enumerate(b2.get_trees()):
        o = ob.tree(i)
        assert np.array_equal(t["split_feature"], o.split_feature) and np.array_equal(t["split_bin"], o.split_bin)
        assert np.array_equal(t["split_type"], o.split_type) and np.array_equal(t["cat_bits"], o.cat_bits)
        n_cat_nodes += int(t["split_type"].sum())
    assert n_cat_nodes > 0
    p2 = predict(b2, RayDMatrix(x, **kw), ray_params=RayParams(num_actors=2))
    assert np.max(np.abs(p2 - ob.predict(x))) <= 1e-5


@pytest.mark.timeout(300)
def test_two_gpu_weighted_rows_identical_to_one_gpu_and_oracle(oracle):
    """Sample weights: weighted quantile sketch (integer rank sums, merged over the ranks) + weighted gradients."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from xgboost_ray_b200 import RayDMatrix, RayParams, train
    rng = np.random.RandomState(13)
    n, f = 30011, 12
    x = rng.normal(size=(n, f)).astype(np.float32)
    x[rng.uniform(size=x.shape) < 0.05] = np.nan
    w = rng.gamma(2.0, 1.0, size=n).astype(np.float32)
    y = (np.nan_to_num(x[:, 0]) - np.nan_to_num(x[:, 2]) + rng.normal(scale=0.3, size=n)).astype(np.float32)
    params = {"objective": "reg:squarederror", "max_depth": 5, "eta": 0.3, "base_score": 0.5, "max_bin": 64}
    b1 = train(params, RayDMatrix(x, y, weight=w), num_boost_round=4, ray_params=RayParams(num_actors=1))
    b2 = train(params, RayDMatrix(x, y, weight=w), num_boost_round=4, ray_params=RayParams(num_actors=2))
    assert _dump(b1) == _dump(b2)
    ob, _ = oracle.train(params, x, y, 4, weight=w)
    for i, t in enumerate(b2.get_trees()):
        o = ob.tree(i)
        assert np.array_equal(t["split_feature"], o.split_feature) and np.array_equal(t["split_bin"], o.split_bin)
        assert np.array_equal(t["split_cond"].view(np.uint32), o.split_cond.view(np.uint32))   # same (weighted) cuts


@pytest.mark.timeout(300)
def test_two_gpu_custom_objective_and_metric():
    """test_xgboost_api.py:77-152 on two GPU actors: custom objective through B2_BoosterBoostOneIter, custom metric
    averaged over the actors (B2_CommAllReduce, xgboost's _allreduce_metric)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.fault_injection import rmsle, squared_log
    from xgboost_ray_b200 import RayDMatrix, RayParams, predict, train
    x = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1], [0, 0, 1, 0]] * 8, np.float32)
    y = np.array([0, 1, 0, 1] * 8, np.float32)
    params = {"booster": "gbtree", "tree_method": "hist", "max_depth": 2, "seed": 1000}
    res1, res2 = {}, {}
    d1, d2 = RayDMatrix(x, y), RayDMatrix(x, y)
    b1 = train(params, d1, evals=[(d1, "dtrain")], evals_result=res1, obj=squared_log, feval=rmsle, ray_params=RayParams(num_actors=1))
    b2 = train(params, d2, evals=[(d2, "dtrain")], evals_result=res2, obj=squared_log, feval=rmsle, ray_params=RayParams(num_actors=2))
    assert _dump(b1) == _dump(b2)
    p2 = np.round(predict(b2, RayDMatrix(x), ray_params=RayParams(num_actors=2)))
    assert list(p2) == list(y)
    assert np.allclose(res1["dtrain"]["PyRMSLE"], res2["dtrain"]["PyRMSLE"], atol=0.1)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_exchange_paths_identical(oracle, exchange, monkeypatch):
    """The NVLink peer-memory exchange (default) and the NCCL reduce-scatter / allgather path give the same model as
    one GPU and as the oracle (deep trees, missing values, uneven shards)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import xgboost_ray_b200.main as M
    from xgboost_ray_b200 import RayDMatrix, RayParams, train
    M.shutdown_actors()                       # the exchange kind is read by the actor processes when they start
    if exchange == "nccl":
        monkeypatch.setenv("B2_EXCHANGE", "nccl")
    else:
        monkeypatch.delenv("B2_EXCHANGE", raising=False)
    rng = np.random.RandomState(21)
    n, f = 60013, 37
    x = rng.uniform(0, 10, size=(n, f)).astype(np.float32)
    x[rng.uniform(size=x.shape) < 0.03] = np.nan
    y = (np.nan_to_num(x[:, 0]) * 0.7 + np.sin(np.nan_to_num(x[:, 5])) * 3 + rng.normal(scale=0.5, size=n)).astype(np.float32)
    params = {"objective": "reg:squarederror", "max_depth": 8, "eta": 0.3, "base_score": 0.5}
    b1 = train(params, RayDMatrix(x, y), num_boost_round=8, ray_params=RayParams(num_actors=1))
    b2 = train(params, RayDMatrix(x, y), num_boost_round=8, ray_params=RayParams(num_actors=2))
    b3 = train(params, RayDMatrix(x, y), num_boost_round=8, ray_params=RayParams(num_actors=2))   # pooled actors, kept communicator
    assert _dump(b1) == _dump(b2) == _dump(b3)
    ob, _ = oracle.train(params, x, y, 8)
    for i, t in enumerate(b2.get_trees()):
        o = ob.tree(i)
        assert np.array_equal(t["split_feature"], o.split_feature) and np.array_equal(t["split_bin"], o.split_bin)
        leaf = o.split_feature < 0
        assert np.max(np.abs(t["value"][leaf] - o.value[leaf])) <= 1e-5
    M.shutdown_actors()


@pytest.mark.timeout(300)
def test_two_gpu_actor_killed_restart_and_elastic_continuation(oracle, tmp_path):
    """test_fault_tolerance.py:401-444 with two GPU actors: rank 1 is killed at round 7 while rank 0 waits for it in the
    histogram exchange -> the stop event aborts the communicator, the dead actor is restarted, training continues from
    checkpoint 5 and the trees equal an uninterrupted run.  With elastic_training the run finishes on the surviving GPU
    alone, quantised with the cut points of the first attempt: the model equals the oracle's for the same schedule."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import xgboost_ray_b200.main as M
    from tests.fault_injection import DieOnceCallback
    from xgboost_ray_b200 import RayDMatrix, RayParams, train
    rng = np.random.RandomState(17)
    n, f = 24001, 12
    x = rng.uniform(0, 10, size=(n, f)).astype(np.float32)
    y = (x[:, 0] + 0.5 * x[:, 3] + rng.normal(size=n) > 7).astype(np.float32)
    params = {"objective": "binary:logistic", "max_depth": 5, "eta": 0.3, "base_score": 0.5}
    ref = train(params, RayDMatrix(x, y), num_boost_round=10, ray_params=RayParams(num_actors=2, checkpoint_frequency=5))
    bst = train(params, RayDMatrix(x, y), num_boost_round=10,
                ray_params=RayParams(num_actors=2, max_actor_restarts=1, checkpoint_frequency=5),
                callbacks=[DieOnceCallback(str(tmp_path / "lock"), rank=1, at=7)])
    assert bst.num_boosted_rounds() == 10 and _dump(bst) == _dump(ref)
    # ---- elastic: rank 0 (the checkpointing rank) dies at round 7; slot 1 finishes rounds 6..9 alone on its shard
    extra = {}
    eb = train(params, RayDMatrix(x, y), num_boost_round=10, additional_results=extra,
               ray_params=RayParams(num_actors=2, elastic_training=True, max_failed_actors=1, max_actor_restarts=1,
                                    checkpoint_frequency=5),
               callbacks=[DieOnceCallback(str(tmp_path / "lock2"), rank=0, at=7)])
    assert eb.num_boosted_rounds() == 10 and extra["total_n"] == len(range(1, n, 2))
    cuts = oracle.Cuts.from_data(x, 256)
    ob = oracle.Booster(params, cuts)
    ob.init_margin(n)
    bins = cuts.bin(x)
    for _ in range(6):
        ob.boost(bins, y)
    xs, ys = np.ascontiguousarray(x[1::2]), np.ascontiguousarray(y[1::2])
    ob.margin = np.ascontiguousarray(ob.predict_margin(xs))
    bins_s = cuts.bin(xs)
    for _ in range(4):
        ob.boost(bins_s, ys)
    for i, t in enumerate(eb.get_trees()):
        o = ob.tree(i)
        assert np.array_equal(t["split_feature"], o.split_feature) and np.array_equal(t["split_bin"], o.split_bin), "tree %d" % i
        leaf = o.split_feature < 0
        assert np.max(np.abs(t["value"][leaf] - o.value[leaf])) <= 1e-5
    M.shutdown_actors()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("ingest", ["nvlink", "host"])
def test_two_gpu_interleaved_shards_redistributed_over_nvlink(ingest, monkeypatch):
    """INTERLEAVED shards of a driver-resident matrix: every rank reads one contiguous half and the rows are exchanged
    on the device (B2_MatrixCreateFromProcessInterleaved); the strided host read (B2_INTERLEAVED_INGEST=0) gives the same
    device matrix, hence byte-identical trees -- also with a row count that is not a multiple of the world size, an eval
    matrix of a different size, and on the pooled second call."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    import xgboost_ray_b200.main as M
    from xgboost_ray_b200 import RayDMatrix, RayParams, train
    M.shutdown_actors()
    monkeypatch.setenv("B2_INTERLEAVED_INGEST", "1" if ingest == "nvlink" else "0")
    rng = np.random.RandomState(33)
    n, f = 70001, 23
    x = rng.normal(size=(n, f)).astype(np.float32)
    x[rng.uniform(size=x.shape) < 0.02] = np.nan
    y = (np.nan_to_num(x[:, 2]) - np.nan_to_num(x[:, 7]) ** 2 + rng.normal(scale=0.3, size=n)).astype(np.float32)
    xe, ye = x[:9999], y[:9999]
    params = {"objective": "reg:squarederror", "max_depth": 7, "eta": 0.2, "base_score": 0.0, "eval_metric": "rmse"}
    r1, r2, add = {}, {}, {}
    b1 = train(params, RayDMatrix(x, y), num_boost_round=5, evals=[(RayDMatrix(xe, ye), "e")], evals_result=r1,
               ray_params=RayParams(num_actors=1))
    b2 = train(params, RayDMatrix(x, y), num_boost_round=5, evals=[(RayDMatrix(xe, ye), "e")], evals_result=r2,
               additional_results=add, ray_params=RayParams(num_actors=2))
    b3 = train(params, RayDMatrix(x, y), num_boost_round=5, ray_params=RayParams(num_actors=2))
    assert add["timing"]["actor0"]["ingest"] == ("nvlink-redistributed" if ingest == "nvlink" else "host")
    assert _dump(b1) == _dump(b2) == _dump(b3)
    assert np.allclose(r1["e"]["rmse"], r2["e"]["rmse"], rtol=0, atol=1e-9)
    M.shutdown_actors()
