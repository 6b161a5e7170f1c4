"""Import seam for the training engine -- the role of xgboost_ray/xgb.py:1-11.

`xgboost` here is the module that provides DMatrix / Booster / train / collective / callback.
It is ALWAYS the sm_100a engine (xgboost_ray_b200.engine) in the product; the environment variable
XGBOOST_RAY_B200_ENGINE exists so the CPU test-suite can substitute a stand-in module that lives
under tests/ (there is no CPU fallback inside the package).
"""
import importlib
import os

xgboost = importlib.import_module(os.environ.get("XGBOOST_RAY_B200_ENGINE", "xgboost_ray_b200.engine"))
