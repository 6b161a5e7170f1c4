"""TEST-ONLY stand-in for xgboost_ray_b200.engine.  The CPU suite installs it with a monkeypatch of the import seam in
the driver process and with the `UseCpuEngine` distributed callback (below) in every spawned actor process; the product
package has no switch for it.

It lets the -m "not gpu" suite drive the host layer (actors, sharding, communicator bootstrap,
callbacks, checkpoints, restart, predict recombination) on a CPU-only box.  Arithmetic is done by
the oracle; the N>1 exchange is a torch.distributed `gloo` all-gather of the row shards (every rank
then grows the same trees on the union -- fixed-point histograms make the model independent of row
order, which is exactly the property the NCCL histogram allreduce provides on the GPU path).
Never imported by the product package.
"""
import json
import os
import pickle
import socket

import numpy as np

from oracle import oracle as O


class XGBoostError(RuntimeError):
    pass


class TrainingCallback:
    def before_training(self, model):
        return model

    def after_training(self, model):
        return model

    def before_iteration(self, model, epoch, evals_log):
        return False

    def after_iteration(self, model, epoch, evals_log):
        return False


class callback:
    TrainingCallback = TrainingCallback


_state = {"rank": 0, "world": 1, "inited": False}


def install():
    """Route xgboost_ray_b200.xgb.xgboost (the import seam of xgboost_ray/xgb.py) to this module in THIS process."""
    import sys
    import xgboost_ray_b200.xgb as seam
    seam.xgboost = sys.modules[__name__]


try:
    from xgboost_ray_b200.callback import DistributedCallback as _DistributedCallback
except Exception:  # pragma: no cover
    _DistributedCallback = object


class UseCpuEngine(_DistributedCallback):
    """Distributed callback whose on_init hook (first thing an actor process runs) installs the stand-in there."""

    def on_init(self, actor, *args, **kwargs):
        install()


def device_count():
    return 0


def get_unique_id():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return ("gloo:%d" % port).encode()


class CommunicatorContext:
    def __init__(self, **args):
        self.args = args

    def __enter__(self):
        world, rank = int(self.args.get("b2_world", 1)), int(self.args.get("b2_rank", 0))
        _state.update(rank=rank, world=world)
        if world > 1:
            import torch.distributed as dist
            port = int(self.args["b2_uid"].decode().split(":")[1])
            dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
            _state["inited"] = True
        return self

    def activate(self):
        return self

    def abort(self):
        pass

    def __exit__(self, *exc):
        if _state["inited"]:
            import torch.distributed as dist
            dist.destroy_process_group()
            _state["inited"] = False
        _state.update(rank=0, world=1)
        return False


class collective:
    CommunicatorContext = CommunicatorContext

    @staticmethod
    def get_rank():
        return _state["rank"]

    @staticmethod
    def get_world_size():
        return _state["world"]


def _allgather(obj):
    if _state["world"] == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * _state["world"]
    dist.all_gather_object(out, obj)
    return out


class DMatrix:
    def __init__(self, data, label=None, weight=None, base_margin=None, missing=None, feature_names=None,
                 feature_types=None, **kw):
        if isinstance(data, (list, tuple)):     # several row blocks of one shard
            data = np.concatenate([np.asarray(b) for b in data], axis=0)
        self.data = np.ascontiguousarray(np.asarray(data), np.float32)
        self.label = None if label is None else np.asarray(label, np.float32).reshape(-1)
        self.weight = None if weight is None else np.asarray(weight, np.float32).reshape(-1)
        self.base_margin = None if base_margin is None else np.asarray(base_margin, np.float32)
        self.missing = np.nan if missing is None else missing
        self.feature_names = feature_names
        self.feature_weights = None

    def set_info(self, feature_weights=None, **kw):
        if feature_weights is not None:
            self.feature_weights = np.asarray(feature_weights, np.float32)

    def get_label(self):
        return self.label if self.label is not None else np.zeros(0, np.float32)

    def num_row(self):
        return self.data.shape[0]

    def num_col(self):
        return self.data.shape[1]


QuantileDMatrix = DeviceQuantileDMatrix = DMatrix


class Booster:
    def __init__(self, params=None):
        self.params = dict(params or {})
        self.ob = None
        self.cuts_arrays = None
        self._trees = []
        self.n_features = None

    def trees(self):
        live = []
        if self.ob is not None:
            for t in self.ob.trees():
                live.append({k: getattr(t, k) for k in O.Tree.FIELDS})
        return list(self._trees) + live

    def __getstate__(self):
        return {"params": self.params, "trees": self.trees(), "n_features": self.n_features}

    def __setstate__(self, st):
        self.params, self.n_features = st["params"], st["n_features"]
        self.ob, self.cuts_arrays = None, None
        self._trees = st["trees"]

    def num_boosted_rounds(self):
        k = int(self.params.get("num_class", 1)) if str(self.params.get("objective", "")).startswith("multi") else 1
        return len(self.trees()) // k

    def _margin(self, X, missing, base_margin=None):
        K = int(self.params.get("num_class", 1)) if str(self.params.get("objective", "")).startswith("multi") else 1
        p = O.make_params(self.params)
        base = float(O.lib().or_base_margin(p))
        out = np.full((X.shape[0], K), base, np.float32) if base_margin is None else \
            np.asarray(base_margin, np.float32).reshape(X.shape[0], -1).copy()
        for ti, t in enumerate(self.trees()):
            nid = np.zeros(X.shape[0], np.int64)
            active = t["split_feature"][nid] >= 0
            while active.any():
                f = t["split_feature"][nid]
                x = X[np.arange(X.shape[0]), np.maximum(f, 0)]
                miss = np.isnan(x)
                go_left = np.where(miss, t["default_left"][nid] == 1, x < t["split_cond"][nid])
                nxt = np.where(go_left, t["left"][nid], t["right"][nid])
                nid = np.where(active, nxt, nid)
                active = t["split_feature"][nid] >= 0
            out[:, ti % K] += t["value"][nid]
        return out

    def predict(self, data, output_margin=False, **kw):
        m = self._margin(data.data, data.missing, data.base_margin)
        obj = self.params.get("objective", "reg:squarederror")
        K = m.shape[1]
        if not output_margin:
            if obj == "binary:logistic":
                m = 1.0 / (1.0 + np.exp(-m))
            elif obj.startswith("multi"):
                e = np.exp(m - m.max(axis=1, keepdims=True))
                m = e / e.sum(axis=1, keepdims=True)
                if obj == "multi:softmax":
                    return np.argmax(m, axis=1).astype(np.float32)
        return m[:, 0].astype(np.float32) if K == 1 else m.astype(np.float32)

    def get_score(self, fmap="", importance_type="weight"):
        cnt, tot = {}, {}
        for t in self.trees():
            for f, g in zip(t["split_feature"], t["loss_chg"]):
                if f >= 0:
                    k = "f%d" % f
                    cnt[k] = cnt.get(k, 0) + 1
                    tot[k] = tot.get(k, 0.0) + float(g)
        if importance_type == "weight":
            return cnt
        return tot if importance_type.startswith("total") else {k: tot[k] / cnt[k] for k in tot}

    get_fscore = get_score

    def get_dump(self, dump_format="json", **kw):
        return [json.dumps({k: np.asarray(v).tolist() for k, v in t.items()}) for t in self.trees()]

    def save_raw(self, *a, **k):
        return bytearray(pickle.dumps(self.__getstate__()))


def train(params, dtrain, num_boost_round=10, evals=(), obj=None, feval=None, maximize=None,
          early_stopping_rounds=None, evals_result=None, verbose_eval=True, xgb_model=None, callbacks=None,
          custom_metric=None):
    callbacks = list(callbacks or [])
    params = dict(params)
    if params.get("objective") not in O.OBJECTIVES and "objective" in params:
        raise XGBoostError("unsupported objective %r" % params.get("objective"))
    # exchange step: union of all shards (order-independent model, see module docstring)
    shards = _allgather((dtrain.data, dtrain.label, dtrain.weight))
    X = np.concatenate([s[0] for s in shards])
    y = np.concatenate([s[1] for s in shards])
    w = None if shards[0][2] is None else np.concatenate([s[2] for s in shards])
    cuts = O.Cuts.from_data(X, int(params.get("max_bin", 256)))
    bins = cuts.bin(X)
    bst = Booster(params)
    bst.n_features = X.shape[1]
    bst.ob = O.Booster(params, cuts)
    if dtrain.feature_weights is not None:
        bst.ob.set_feature_weights(dtrain.feature_weights)
    bst.ob.init_margin(X.shape[0])
    if xgb_model is not None:
        prev = xgb_model.trees() if isinstance(xgb_model, Booster) else pickle.loads(bytes(xgb_model))["trees"]
        bst._trees = list(prev)
        tmp = Booster(params)
        tmp._trees = list(prev)
        bst.ob.margin[:] = tmp._margin(X, np.nan)
    evals_log = {}
    for cb in callbacks:
        bst = cb.before_training(bst) or bst
    for epoch in range(num_boost_round):
        if any(cb.before_iteration(bst, epoch, evals_log) for cb in callbacks):
            break
        if obj is not None:
            g, h = obj(bst.ob.margin.reshape(-1) if bst.ob.K == 1 else bst.ob.margin, DMatrix(X, y))
            bst.ob.boost(bins, y, w, custom_g=np.asarray(g, np.float32), custom_h=np.asarray(h, np.float32))
        else:
            bst.ob.boost(bins, y, w)
        for dm, name in evals:
            metric = {"reg:squarederror": "rmse", "binary:logistic": "logloss"}.get(params.get("objective", "reg:squarederror"), "mlogloss")
            if dm is dtrain:
                v = bst.ob.metric(metric, bst.ob.margin, y, w)
            else:
                parts = _allgather((dm.data, dm.label))
                Xe, ye = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
                v = bst.ob.metric(metric, bst._margin(Xe, np.nan), ye)
            evals_log.setdefault(name, {}).setdefault(metric, []).append(v)
            fm = custom_metric or feval
            if fm is not None:   # custom metric on the (union) matrix: every rank reports the same value
                Xm, ym = (X, y) if dm is dtrain else (Xe, ye)
                pm = bst.ob.margin.reshape(-1) if dm is dtrain and bst.ob.K == 1 else bst._margin(Xm, np.nan).reshape(-1)
                mname, mval = fm(pm, DMatrix(Xm, ym))
                evals_log[name].setdefault(mname, []).append(float(mval))
        if any([cb.after_iteration(bst, epoch, evals_log) for cb in callbacks]):
            break
    for cb in callbacks:
        bst = cb.after_training(bst) or bst
    if evals_result is not None:
        evals_result.update(evals_log)
    # freeze trees so the object pickles without the live oracle handle
    out = Booster(params)
    out.n_features, out._trees = bst.n_features, bst.trees()
    return out
