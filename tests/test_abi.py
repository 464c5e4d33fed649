"""CPU: the C-ABI library loads and exports every symbol include/pnm_b200.h declares; the
entry points fail loudly (no CPU fallback) when no device is usable."""
import ctypes
import os
import re

import numpy as np
import pytest

from pynmap_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pnm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), "libpnm_b200.so does not export %s" % name
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)


def test_version_and_constants_match_header():
    lib = _lib.load()
    text = open(os.path.join(ROOT, "include", "pnm_b200.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+PNM_([A-Z0-9_]+)\s+\(?(-?\d+)\)?", text)}
    assert lib.pnm_version() == defs["VERSION"]
    for py, c in (("F32", "F32"), ("F64", "F64"), ("ACC_F64", "ACC_F64"), ("ACC_FIXED", "ACC_FIXED"),
                  ("SUM_W", "SUM_W"), ("SUM_WD", "SUM_WD"), ("SUM_WD2", "SUM_WD2"), ("SUM_CUBE", "SUM_CUBE"),
                  ("W_FROM_CUBE", "W_FROM_CUBE"), ("MASK_AND_NONFINITE", "MASK_AND_NONFINITE"),
                  ("MASK_PLAIN", "MASK_PLAIN"), ("MAX_EDGES", "MAX_EDGES"), ("MAX_CHAIN", "MAX_CHAIN"),
                  ("MAX_VIEWS", "MAX_VIEWS"), ("MAX_TAPS", "MAX_TAPS"), ("MAP_SLOTS", "MAP_SLOTS")):
        assert getattr(_lib, py) == defs[c], py


def test_argument_validation_needs_no_device():
    """Bad arguments are rejected before any CUDA call, with a message."""
    lib = _lib.load()
    h = ctypes.c_void_p()
    e = _lib.doubles([0.0, 1.0, 0.5])          # not increasing
    assert lib.pnm_grid_create(2, e, 2, e, 0, None, ctypes.byref(h)) == -1
    assert b"increase" in lib.pnm_last_error()
    assert lib.pnm_rotate(None, None, None, None, None, None, -1, 0, None, None) == -1
    assert lib.pnm_bin_points(None, 0, 0, None, None, None, None, None, 0, 7, 0, None, None, None, None) == -1
    with pytest.raises(_lib.PnmError):
        _lib.call("pnm_smooth_cube", None, None, None, 1, 1, 1, None, 1, None, 1, None)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pynmap_b200
    x = np.zeros(8, dtype=np.float32)
    with pytest.raises(_lib.PnmError):
        pynmap_b200.grid.points_2vmaps(x, x, x, nXY=4)
    with pytest.raises(_lib.PnmError):
        pynmap_b200.rotation.rotate_mat(np.zeros((4, 3), dtype=np.float32), None, 30., 60.)
    with pytest.raises(_lib.PnmError):
        pynmap_b200.snapshot.from_arrays(np.zeros((4, 3)), np.zeros((4, 3)))
    v = np.linspace(-100.0, 100.0, 32)
    with pytest.raises(_lib.PnmError):
        pynmap_b200.fit_losvd.fitgh(v, np.exp(-v * v / 800.0), deg_gh=4, verbose=False)
    with pytest.raises(_lib.PnmError):
        pynmap_b200.LOSVD(v[:4], v[:4], v, np.zeros((4, 4, 32))).fitlosvd()
    with pytest.raises(_lib.PnmError):
        pynmap_b200.amr.spread_amr([v], np.full(32, 5), nsample=2)
    with pytest.raises(_lib.PnmError):
        pynmap_b200.grid.point_2vprofiles(v, v, v * 0, v, v, v, nbins=5)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pynmap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), "%s mentions the oracle" % f
