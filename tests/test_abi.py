"""The C-ABI library loads on a CPU-only box and exports every symbol include/b2hist.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from xgboost_ray_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2hist.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(B2_\w+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s


def test_python_binding_covers_header(built):
    from xgboost_ray_b200 import engine
    assert sorted(engine.ABI) == declared_symbols()
    assert engine.lib().B2_GetVersion() >= 100


def test_no_cpu_fallback(built):
    from xgboost_ray_b200 import engine
    if engine.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(engine.XGBoostError, match="no CUDA device|CUDA"):
        engine.DMatrix(np.zeros((4, 2), np.float32), label=np.zeros(4, np.float32))
    with pytest.raises(engine.XGBoostError):
        engine.hist_build_raw(np.zeros((4, 2), np.uint8), np.zeros(4, np.int32), np.zeros(4, np.int32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "xgboost_ray_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, fn
