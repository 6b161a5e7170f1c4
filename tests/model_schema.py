"""XGBoost's JSON model schema (dmlc/xgboost doc/model.schema, 2.x), restated for the gbtree / gbtree-with-categories
part that this engine writes.  Upstream source is not in /root/reference (the reference only round-trips models through
`bst.save_model` / `xgb.Booster(model_file=...)`, xgboost_ray/tests/utils.py:107-108, test_fault_tolerance.py:356-361);
this restatement is what tests/test_model_json.py validates Booster.save_raw() against."""

_NUM_ARRAY = {"type": "array", "items": {"type": "number"}}
_INT_ARRAY = {"type": "array", "items": {"type": "integer"}}
_STR = {"type": "string"}

TREE = {
    "type": "object",
    "properties": {
        "tree_param": {
            "type": "object",
            "properties": {"num_nodes": _STR, "size_leaf_vector": _STR, "num_feature": _STR, "num_deleted": _STR},
            "required": ["num_nodes", "num_feature", "size_leaf_vector"],
        },
        "id": {"type": "integer"},
        "loss_changes": _NUM_ARRAY, "sum_hessian": _NUM_ARRAY, "base_weights": _NUM_ARRAY,
        "left_children": _INT_ARRAY, "right_children": _INT_ARRAY, "parents": _INT_ARRAY,
        "split_indices": _INT_ARRAY, "split_conditions": _NUM_ARRAY, "split_type": _INT_ARRAY,
        "default_left": _INT_ARRAY,
        "categories": _INT_ARRAY, "categories_nodes": _INT_ARRAY, "categories_segments": _INT_ARRAY,
        "categories_sizes": _INT_ARRAY,
    },
    "required": ["tree_param", "loss_changes", "sum_hessian", "base_weights", "left_children", "right_children", "parents",
                 "split_indices", "split_conditions", "default_left"],
    "additionalProperties": False,
}

GBTREE = {
    "type": "object",
    "properties": {
        "name": {"const": "gbtree"},
        "model": {
            "type": "object",
            "properties": {
                "gbtree_model_param": {
                    "type": "object",
                    "properties": {"num_trees": _STR, "num_parallel_tree": _STR, "size_leaf_vector": _STR},
                    "required": ["num_trees", "num_parallel_tree"],
                    "additionalProperties": False,
                },
                "trees": {"type": "array", "items": TREE},
                "tree_info": _INT_ARRAY,
                "iteration_indptr": _INT_ARRAY,
            },
            "required": ["gbtree_model_param", "trees", "tree_info"],
            "additionalProperties": False,
        },
    },
    "required": ["name", "model"],
    "additionalProperties": False,
}

_REG_LOSS = {"type": "object", "properties": {"scale_pos_weight": _STR}, "required": ["scale_pos_weight"],
             "additionalProperties": False}
_SOFTMAX = {"type": "object", "properties": {"num_class": _STR}, "required": ["num_class"], "additionalProperties": False}

OBJECTIVE = {
    "oneOf": [
        {"type": "object",
         "properties": {"name": {"enum": ["reg:squarederror", "reg:linear", "reg:logistic", "binary:logistic", "binary:logitraw"]},
                        "reg_loss_param": _REG_LOSS},
         "required": ["name", "reg_loss_param"], "additionalProperties": False},
        {"type": "object",
         "properties": {"name": {"enum": ["multi:softprob", "multi:softmax"]}, "softmax_multiclass_param": _SOFTMAX},
         "required": ["name", "softmax_multiclass_param"], "additionalProperties": False},
    ]
}

MODEL = {
    "type": "object",
    "properties": {
        "version": {"type": "array", "items": {"type": "integer"}, "minItems": 3, "maxItems": 3},
        "learner": {
            "type": "object",
            "properties": {
                "attributes": {"type": "object", "additionalProperties": _STR},
                "feature_names": {"type": "array", "items": _STR},
                "feature_types": {"type": "array", "items": _STR},
                "gradient_booster": GBTREE,
                "objective": OBJECTIVE,
                "learner_model_param": {
                    "type": "object",
                    "properties": {"base_score": _STR, "boost_from_average": _STR, "num_class": _STR, "num_feature": _STR,
                                   "num_target": _STR},
                    "required": ["base_score", "num_class", "num_feature"],
                    "additionalProperties": False,
                },
            },
            "required": ["gradient_booster", "objective", "learner_model_param"],
            "additionalProperties": False,
        },
    },
    "required": ["version", "learner"],
    "additionalProperties": False,
}
