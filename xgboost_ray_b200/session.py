"""Per-actor session: rank + queue back to the driver (mirror of xgboost_ray/session.py:8-81)."""
from typing import Optional


class RayXGBoostSession:
    def __init__(self, rank: int, queue=None):
        self._rank = rank
        self._queue = queue

    def get_actor_rank(self):
        return self._rank

    def set_queue(self, queue):
        self._queue = queue

    def put_queue(self, item):
        if self._queue is None:
            raise ValueError("Trying to put something into session queue, but queue was not initialized.")
        self._queue.put((self._rank, item))


_session: Optional[RayXGBoostSession] = None


def init_session(*args, **kwargs):
    global _session
    if _session:
        raise ValueError("Trying to initialize RayXGBoostSession twice."
                         "\nFIX THIS by not calling `init_session()` manually.")
    _session = RayXGBoostSession(*args, **kwargs)


def _reset_session():
    global _session
    _session = None


def get_session() -> RayXGBoostSession:
    if not _session or not isinstance(_session, RayXGBoostSession):
        raise ValueError("Trying to access RayXGBoostSession from outside an XGBoost run."
                         "\nFIX THIS by calling function in `session.py` like `get_actor_rank()` only from within "
                         "an XGBoost actor session.")
    return _session


def set_session_queue(queue):
    get_session().set_queue(queue)


def get_actor_rank() -> int:
    return get_session().get_actor_rank()


def get_rabit_rank() -> int:
    """Rank inside the (NCCL) communicator -- session.py:68-75 returned xgb.collective.get_rank()."""
    from xgboost_ray_b200.xgb import xgboost as xgb
    return xgb.collective.get_rank()


def put_queue(*args, **kwargs):
    get_session().put_queue(*args, **kwargs)
