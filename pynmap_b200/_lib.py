"""ctypes binding of libpnm_b200.so -- the C ABI declared in include/pnm_b200.h.

There is deliberately no fallback: if the shared library is missing or no sm_100
device is usable, importing succeeds (so that CPU-only unit tests can exercise the
host logic) but the first call that needs the device raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNM_B200_LIB", os.path.join(_HERE, "libpnm_b200.so"))

# constants mirrored from include/pnm_b200.h
F32, F64 = 0, 1
ACC_F64, ACC_FIXED = 0, 1
SUM_W, SUM_WD, SUM_WD2, SUM_CUBE, W_FROM_CUBE, MAPS_FOLDED, VAL_D32, VAL_W32, CUBE_MOMENTS = 1, 2, 4, 8, 16, 32, 64, 128, 256
MASK_NONE, MASK_AND_NONFINITE, MASK_PLAIN = 0, 1, 2
MAX_EDGES, MAX_CHAIN, MAX_VIEWS, MAX_TAPS, MAP_SLOTS = 4097, 8, 64, 255, 4
MAX_MATS = 64

_P = c_void_p
_SIGNATURES = {
    "pnm_version": (c_int, []),
    "pnm_last_error": (c_char_p, []),
    "pnm_device_count": (c_int, []),
    "pnm_rotate": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, POINTER(c_float), _P]),
    "pnm_grid_create": (c_int, [c_int32, POINTER(c_double), c_int32, POINTER(c_double), c_int32,
                                POINTER(c_double), POINTER(_P)]),
    "pnm_grid_destroy": (c_int, [_P]),
    "pnm_maps_acc_elems": (c_int64, [_P]),
    "pnm_cube_acc_elems": (c_int64, [_P]),
    "pnm_grid_map_replicas": (c_int, [_P]),
    "pnm_fold_maps": (c_int, [_P, _P, c_int, c_int, _P]),
    "pnm_project_bin": (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int,
                                POINTER(c_float), c_int, c_int, c_int, c_int, POINTER(c_double), _P, _P, _P]),
    "pnm_bin_points": (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int,
                               POINTER(c_double), _P, _P, _P]),
    "pnm_finalize_vmaps": (c_int, [_P, _P, _P, c_int, c_int, POINTER(c_double), _P, _P, _P, _P, _P, _P]),
    "pnm_finalize_map": (c_int, [_P, _P, c_int, POINTER(c_double), _P, _P, _P, _P]),
    "pnm_finalize_cube": (c_int, [_P, _P, c_int, c_int, c_double, _P, _P]),
    "pnm_cube_moments_acc_elems": (c_int64, [_P]),
    "pnm_smooth_cube": (c_int, [_P, _P, _P, c_int32, c_int32, c_int32, POINTER(c_double), c_int32,
                                POINTER(c_double), c_int32, _P]),
    "pnm_cube_moments": (c_int, [_P, c_int32, c_int32, c_int32, _P, c_double, _P, _P, _P, _P]),
    "pnm_profile_acc_elems": (c_int64, [c_int32]),
    "pnm_profile_rmax": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_double, c_double, _P, _P]),
    "pnm_radial_profiles": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_double, c_double, _P, c_int32,
                                    _P, _P, _P, _P, _P]),
    "pnm_amr_jitter": (c_int, [_P, c_int, _P, _P, c_int64, c_int32, c_double, _P, _P]),
    "pnm_amr_repeat": (c_int, [_P, c_int, c_int64, c_int32, c_int, _P, _P]),
    "pnm_fit_gh": (c_int, [_P, _P, c_int64, c_int32, _P, c_int32, POINTER(c_double), POINTER(c_double),
                           POINTER(c_double), POINTER(c_int32), c_double, c_double, c_double, c_double, c_double,
                           c_int32, _P, _P, _P, _P, _P, _P]),
    "pnm_rebin_cube": (c_int, [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int, _P]),
    "pnm_ct_interp": (c_int, [_P, _P, c_int64, c_int, _P, _P, c_int32, _P, _P, c_int32, c_double, c_double, c_double,
                              c_double, _P, c_int32, c_int, _P, _P]),
    "pnm_absmax": (c_int, [_P, c_int64, c_int, _P, _P]),
    "pnm_accumulate_host": (c_int, [_P, c_int, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, POINTER(c_float), c_int,
                                    c_int, c_int, POINTER(c_double), _P, _P, _P]),
    "pnm_project_host": (c_int, [c_int, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, POINTER(c_float), c_int,
                                 c_int32, POINTER(c_double), c_int32, POINTER(c_double), c_int32,
                                 POINTER(c_double), c_int, POINTER(c_double),
                                 POINTER(c_double), c_int32, POINTER(c_double), c_int32, _P, _P, _P, _P]),
    "pnm_host_release": (c_int, []),
    "pnm_copy_async": (c_int, [_P, _P, c_int64, _P]),
    "pnm_zero_async": (c_int, [_P, c_int64, _P]),
    "pnm_pull_chunks": (c_int, [_P, POINTER(_P), c_int32, c_int32, c_int64, _P]),
    "pnm_sum_peers": (c_int, [POINTER(_P), c_int32, c_int64, c_int, _P, _P]),
    "pnm_sum_chunks": (c_int, [_P, c_int32, c_int64, c_int, _P, _P]),
    "pnm_slab_sums": (c_int, [_P, _P, c_int32, c_int32, c_int, _P, _P]),
    "pnm_finalize_sums": (c_int, [_P, _P, c_int, POINTER(c_double), _P, _P, _P, _P]),
    "pnm_finalize_cube_slabs": (c_int, [_P, _P, c_int32, c_int32, c_int, c_double, _P, _P]),
    "pnm_order_probe": (c_int, [_P, _P, _P, c_int64, c_int, c_int64, c_int64, c_double, _P, _P]),
    "pnm_cube_slots_max": (c_int32, []),
    "pnm_cube_plan_create": (c_int, [_P, POINTER(_P)]),
    "pnm_cube_plan_destroy": (c_int, [_P]),
    "pnm_cube_plan_build": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, POINTER(c_float), _P]),
    "pnm_cube_plan_info": (c_int, [_P, POINTER(c_int32), POINTER(c_double), _P]),
    "pnm_cube_plan_set_reserved_sms": (c_int, [_P, c_int32]),
    "pnm_project_cube": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, POINTER(c_float), c_int,
                                 POINTER(c_double), _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_device_ok = False


class PnmError(RuntimeError):
    """A libpnm_b200 entry point returned a negative status."""


def load():
    """dlopen the library once; raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                "libpnm_b200.so not found at %s: build it with `python -m pynmap_b200.csrc.build` "
                "(needs nvcc; pynmap_b200 has no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error():
    return load().pnm_last_error().decode("utf-8", "replace")


def check(status, what=""):
    if status != 0:
        raise PnmError("%s failed (%d): %s" % (what or "libpnm_b200 call", status, last_error()))


def call(name, *args):
    """Invoke an int-returning entry point and raise on a negative status."""
    check(getattr(load(), name)(*args), name)


def require_device():
    """Fail loudly when the CUDA path cannot run (no silent CPU path exists)."""
    global _device_ok
    lib = load()
    if not _device_ok:
        if lib.pnm_device_count() < 1:
            raise PnmError("no usable sm_100 (B200) device: pynmap_b200 has no CPU fallback")
        _device_ok = True        # checked once per process: the query costs microseconds per call
    return lib


def doubles(seq):
    arr = (c_double * len(seq))(*[float(v) for v in seq])
    return arr


def floats(seq):
    arr = (c_float * len(seq))(*[float(v) for v in seq])
    return arr
