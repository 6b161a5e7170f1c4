"""Distributed (per-actor) callbacks -- mirror of xgboost_ray/callback.py:14-110."""
import os
from typing import Any, Dict, Sequence


class DistributedCallback:
    """Hooks run on every actor around init / data loading / train / predict."""

    def on_init(self, actor, *args, **kwargs):
        pass

    def before_data_loading(self, actor, data, *args, **kwargs):
        pass

    def after_data_loading(self, actor, data, *args, **kwargs):
        pass

    def before_train(self, actor, *args, **kwargs):
        pass

    def after_train(self, actor, result_dict: Dict, *args, **kwargs):
        pass

    def before_predict(self, actor, *args, **kwargs):
        pass

    def after_predict(self, actor, predictions, *args, **kwargs):
        pass


class DistributedCallbackContainer:
    def __init__(self, callbacks: Sequence[DistributedCallback]):
        self.callbacks = callbacks or []

    def _each(self, name, *args, **kwargs):
        for cb in self.callbacks:
            getattr(cb, name)(*args, **kwargs)

    def on_init(self, actor, *a, **k):
        self._each("on_init", actor, *a, **k)

    def before_data_loading(self, actor, data, *a, **k):
        self._each("before_data_loading", actor, data, *a, **k)

    def after_data_loading(self, actor, data, *a, **k):
        self._each("after_data_loading", actor, data, *a, **k)

    def before_train(self, actor, *a, **k):
        self._each("before_train", actor, *a, **k)

    def after_train(self, actor, result_dict, *a, **k):
        self._each("after_train", actor, result_dict, *a, **k)

    def before_predict(self, actor, *a, **k):
        self._each("before_predict", actor, *a, **k)

    def after_predict(self, actor, predictions, *a, **k):
        self._each("after_predict", actor, predictions, *a, **k)


class EnvironmentCallback(DistributedCallback):
    def __init__(self, env_dict: Dict[str, Any]):
        self.env_dict = env_dict

    def on_init(self, actor, *args, **kwargs):
        os.environ.update(self.env_dict)
