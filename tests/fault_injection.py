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
