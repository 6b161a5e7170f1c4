"""B200-native histogram-tree training path behind the xgboost_ray public surface
(xgboost_ray/__init__.py:1-41): train, predict, RayParams, RayDMatrix, ..."""
from xgboost_ray_b200.main import RayParams, RayXGBoostActor, predict, train  # noqa: F401
from xgboost_ray_b200.matrix import (RayDeviceQuantileDMatrix, RayDMatrix, RayFileType,  # noqa: F401
                                     RayQuantileDMatrix, RayShardingMode, combine_data)

__version__ = "0.1.0"
__all__ = ["__version__", "RayParams", "RayDMatrix", "RayDeviceQuantileDMatrix", "RayQuantileDMatrix", "RayFileType",
           "RayShardingMode", "train", "predict", "RayXGBoostActor", "combine_data"]
