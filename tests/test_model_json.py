"""f1 -- XGBoost-loadable model JSON: what Booster.save_model writes validates against XGBoost's
model schema (restated in tests/model_schema.py), is strict JSON (no NaN / Infinity literals), keeps this engine's
private data under `attributes`, and a file written the way XGBoost 2.x writes it loads back.  CPU only: the model
object is host-side Python; no kernel is involved."""
import json
import os

import jsonschema
import numpy as np
import pytest

from tests.model_schema import MODEL

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _strict_loads(raw):
    def bad(c):
        raise ValueError("non-JSON constant %r in the model file" % c)
    return json.loads(raw, parse_constant=bad)


def _booster_from_golden(name):
    """A host-side Booster filled with the trees of a committed golden fixture."""
    from xgboost_ray_b200 import engine as E
    g = json.load(open(os.path.join(GOLD, name)))
    params = dict(g["params"])
    b = E.Booster(params)
    b.n_features = len(g["cut_ptrs"]) - 1
    for t in g["trees"]:
        n = len(t["left"])
        left = np.asarray(t["left"], np.int32)
        parent = np.full(n, -1, np.int32)
        for i in range(n):
            if left[i] >= 0:
                parent[left[i]] = i
                parent[t["right"][i]] = i
        st = np.asarray(t.get("split_type", [0] * n), np.uint8)
        bits = np.zeros((n, 8), np.uint32)
        for nid, cats in (t.get("categories") or {}).items() if isinstance(t.get("categories"), dict) else []:
            for c in cats:
                bits[int(nid), c >> 5] |= np.uint32(1 << (c & 31))
        b._trees.append(dict(left=left, right=np.asarray(t["right"], np.int32), parent=parent,
                             split_feature=np.asarray(t["split_feature"], np.int32), split_bin=np.asarray(t["split_bin"], np.int32),
                             split_cond=np.nan_to_num(np.asarray(t["split_cond"], np.float32)) if False else np.asarray(t["split_cond"], np.float32),
                             default_left=np.asarray(t["default_left"], np.uint8), value=np.asarray(t["value"], np.float32),
                             base_weight=np.asarray(t["value"], np.float32), loss_chg=np.asarray(t["loss_chg"], np.float32),
                             sum_hess=np.ones(n, np.float64), split_type=st, cat_bits=bits))
    return b, g


@pytest.mark.parametrize("fixture", ["breast_cancer_logistic.json", "synthetic_categorical_softprob.json",
                                     "synthetic_missing_regression.json", "toy_softmax.json"])
def test_saved_model_validates_against_the_xgboost_schema(fixture, tmp_path):
    b, g = _booster_from_golden(fixture)
    raw = bytes(b.save_raw())
    d = _strict_loads(raw.decode())                                   # strict JSON: no bare NaN / Infinity
    jsonschema.validate(d, MODEL)
    L = d["learner"]
    obj = g["params"].get("objective", "reg:squarederror")
    assert L["objective"]["name"] == obj
    if obj.startswith("multi:"):
        assert L["objective"]["softmax_multiclass_param"]["num_class"] == str(g["params"]["num_class"])
        assert L["learner_model_param"]["num_class"] == str(g["params"]["num_class"])
    else:
        assert float(L["objective"]["reg_loss_param"]["scale_pos_weight"]) == 1.0
        assert L["learner_model_param"]["num_class"] == "0"
    assert float(L["learner_model_param"]["base_score"]) == float(g["params"].get("base_score", 0.5))
    assert L["gradient_booster"]["model"]["gbtree_model_param"]["num_trees"] == str(len(g["trees"]))
    assert all(k.startswith("b2.") for k in L["attributes"])           # private data lives under attributes only
    # round trip: same trees, same parameters, byte-identical file
    from xgboost_ray_b200 import engine as E
    f = str(tmp_path / "m.json")
    b.save_model(f)
    b2 = E.Booster(model_file=f)
    assert bytes(b2.save_raw()) == raw
    for t, u in zip(b._trees, b2._trees):
        leaf = t["split_feature"] < 0
        for k in ("left", "right", "parent", "split_feature", "split_bin", "default_left", "split_type", "cat_bits"):
            assert np.array_equal(t[k], u[k]), k
        assert np.array_equal(t["value"][leaf], u["value"][leaf])
        num = ~leaf & (t["split_type"] == 0)
        assert np.array_equal(t["split_cond"][num], u["split_cond"][num])
    assert b2.get_dump(dump_format="json") == b.get_dump(dump_format="json")


XGB_STYLE = {   # written the way XGBoost 2.0 writes a 2-tree binary:logistic model with 3 features (no engine-private keys)
    "learner": {
        "attributes": {"best_iteration": "1"},
        "feature_names": ["a", "b", "c"], "feature_types": ["float", "float", "float"],
        "gradient_booster": {"model": {
            "gbtree_model_param": {"num_parallel_tree": "1", "num_trees": "2"},
            "iteration_indptr": [0, 1, 2], "tree_info": [0, 0],
            "trees": [
                {"base_weights": [0.1, -0.4, 0.6], "categories": [], "categories_nodes": [], "categories_segments": [],
                 "categories_sizes": [], "default_left": [1, 0, 0], "id": 0, "left_children": [1, -1, -1],
                 "loss_changes": [12.5, 0.0, 0.0], "parents": [2147483647, 0, 0], "right_children": [2, -1, -1],
                 "split_conditions": [2.5, -0.12, 0.18], "split_indices": [1, 0, 0], "split_type": [0, 0, 0],
                 "sum_hessian": [25.0, 10.0, 15.0],
                 "tree_param": {"num_deleted": "0", "num_feature": "3", "num_nodes": "3", "size_leaf_vector": "1"}},
                {"base_weights": [0.0, 0.2, -0.3], "categories": [], "categories_nodes": [], "categories_segments": [],
                 "categories_sizes": [], "default_left": [0, 0, 0], "id": 1, "left_children": [1, -1, -1],
                 "loss_changes": [3.25, 0.0, 0.0], "parents": [2147483647, 0, 0], "right_children": [2, -1, -1],
                 "split_conditions": [-1.0, 0.06, -0.09], "split_indices": [2, 0, 0], "split_type": [0, 0, 0],
                 "sum_hessian": [24.0, 12.0, 12.0],
                 "tree_param": {"num_deleted": "0", "num_feature": "3", "num_nodes": "3", "size_leaf_vector": "1"}}]},
            "name": "gbtree"},
        "learner_model_param": {"base_score": "2.5E-1", "boost_from_average": "1", "num_class": "0", "num_feature": "3",
                                "num_target": "1"},
        "objective": {"name": "binary:logistic", "reg_loss_param": {"scale_pos_weight": "2"}}},
    "version": [2, 0, 3]}


def test_loads_a_file_written_the_way_xgboost_writes_it(tmp_path):
    from xgboost_ray_b200 import engine as E
    jsonschema.validate(XGB_STYLE, MODEL)                              # the hand-written file is itself schema-valid
    f = str(tmp_path / "xgb.json")
    json.dump(XGB_STYLE, open(f, "w"))
    b = E.Booster(model_file=f)
    assert b.params["objective"] == "binary:logistic" and float(b.params["base_score"]) == 0.25
    assert float(b.params["scale_pos_weight"]) == 2.0
    assert b.num_features() == 3 and b.num_trees() == 2 and b.num_boosted_rounds() == 2
    assert b.feature_names == ["a", "b", "c"] and b.attr("best_iteration") == "1"
    t0 = b._trees[0]
    assert list(t0["split_feature"]) == [1, -1, -1] and t0["split_cond"][0] == np.float32(2.5)
    assert list(t0["value"][1:]) == [np.float32(-0.12), np.float32(0.18)] and list(t0["parent"]) == [-1, 0, 0]
    dump = b.get_dump()
    assert dump[0].startswith("0:[b<2.5] yes=1,no=2,missing=1") and "leaf=-0.119999997" in dump[0]
    d = _strict_loads(bytes(b.save_raw()).decode())
    jsonschema.validate(d, MODEL)
    assert d["learner"]["objective"]["reg_loss_param"]["scale_pos_weight"] == "2"
    assert float(d["learner"]["learner_model_param"]["base_score"]) == 0.25
    assert d["learner"]["gradient_booster"]["model"]["trees"][1]["split_conditions"][0] == -1.0
