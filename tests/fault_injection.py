"""Test helpers importable in spawned actor processes: fault injection callbacks
(port of xgboost_ray/tests/utils.py:111-142 `_kill_callback`)."""
import os

from tests.cpu_engine import TrainingCallback


class DieOnceCallback(TrainingCallback):
    """kill -9 the actor of `rank` at global epoch `at`, once (guarded by a lock file)."""

    def __init__(self, lockfile, rank=1, at=6):
        self.lockfile, self.rank, self.at = lockfile, rank, at

    def after_iteration(self, model, epoch, evals_log):
        from xgboost_ray_b200.session import get_actor_rank
        if get_actor_rank() == self.rank and epoch == self.at and not os.path.exists(self.lockfile):
            with open(self.lockfile, "w") as f:
                f.write("died")
            os.kill(os.getpid(), 9)
        return False


class RankRecorder(TrainingCallback):
    def after_iteration(self, model, epoch, evals_log):
        from xgboost_ray_b200.session import get_actor_rank, put_queue
        if epoch == 0:
            put_queue(("rank", get_actor_rank()))
        return False


# ---- custom objective / metric of xgboost_ray/tests/test_xgboost_api.py:20-44 (top level: picklable for the actors)
def squared_log(predt, dtrain):
    import numpy as np
    y = dtrain.get_label()
    predt = np.asarray(predt, np.float64).copy()
    predt[predt < -1] = -1 + 1e-6
    grad = (np.log1p(predt) - np.log1p(y)) / (predt + 1)
    hess = (-np.log1p(predt) + np.log1p(y) + 1) / np.power(predt + 1, 2)
    return grad, hess


def rmsle(predt, dtrain):
    import numpy as np
    y = dtrain.get_label()
    predt = np.asarray(predt, np.float64).copy()
    predt[predt < -1] = -1 + 1e-6
    elements = np.power(np.log1p(y) - np.log1p(predt), 2)
    return "PyRMSLE", float(np.sqrt(np.sum(elements) / len(y)))


class PidRecorder(TrainingCallback):
    def after_iteration(self, model, epoch, evals_log):
        from xgboost_ray_b200.session import get_actor_rank, put_queue
        if epoch == 0:
            put_queue(("pid", os.getpid()))
        return False
