"""Import seam for the training engine -- the role of xgboost_ray/xgb.py:1-11.

`xgboost` here is the module that provides DMatrix / Booster / train / collective / callback: the sm_100a engine
(xgboost_ray_b200.engine).  There is no switch and no CPU fallback in the package.
"""
from xgboost_ray_b200 import engine as xgboost  # noqa: F401
