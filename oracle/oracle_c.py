"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/pnm_oracle.c (the plain-C restatement).

Mirrors the calling convention of libpnm_b200's entry points with HOST numpy arrays, so that a
parity test reads:  gpu = pnm_project_bin(...); cpu = oracle_c.project_bin(...); compare.
Never imported by the product package.
"""
import ctypes
from ctypes import c_int, c_int64, c_void_p

import numpy as np

from . import build_oracle

F32, F64 = 0, 1
ACC_F64, ACC_FIXED = 0, 1
SUM_W, SUM_WD, SUM_WD2, SUM_CUBE, W_FROM_CUBE = 1, 2, 4, 8, 16
MASK_NONE, MASK_AND_NONFINITE, MASK_PLAIN = 0, 1, 2

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_oracle.build())
        P = c_void_p
        L.pnmo_rotate.argtypes = [c_int, P, P, P, P, P, P, c_int64, P]
        L.pnmo_bin_index.argtypes = [c_int, P, c_int64, P, c_int, P]
        L.pnmo_project_bin.argtypes = [c_int, c_int64, c_int, P, P, P, P, P, P, P, P, c_int, P, c_int, c_int,
                                       c_int, P, c_int, P, c_int, P, c_int, c_int, P, P, P]
        L.pnmo_finalize_vmaps.argtypes = [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P]
        L.pnmo_smooth_cube.argtypes = [P, P, c_int, c_int, c_int, P, c_int, c_int]
        L.pnmo_cube_moments.argtypes = [P, c_int, c_int, c_int, P, P, P, P]
        for f in (L.pnmo_rotate, L.pnmo_bin_index, L.pnmo_project_bin, L.pnmo_finalize_vmaps,
                  L.pnmo_smooth_cube, L.pnmo_cube_moments):
            f.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(c_void_p) if a is not None else None


def _dt(a):
    return F32 if a.dtype == np.float32 else F64


def _c(a, dtype=None):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def rotate(mat, x, y, z):
    x, y, z = _c(x), _c(y), _c(z)
    ox, oy, oz = np.empty_like(x), np.empty_like(y), np.empty_like(z)
    m = _c(mat, np.float32).reshape(9)
    lib().pnmo_rotate(_dt(x), _p(x), _p(y), _p(z), _p(ox), _p(oy), _p(oz), x.size, _p(m))
    return ox, oy, oz


def bin_index(values, edges):
    v = _c(values)
    e = _c(edges, np.float64)
    out = np.empty(v.size, dtype=np.int32)
    lib().pnmo_bin_index(_dt(v), _p(v), v.size, _p(e), e.size - 1, _p(out))
    return out


def project_bin(soa, w, mats, nchain, nviews, ex, ey, ev, sums, acc_mode=ACC_F64, scales=None,
                mask=None, mask_mode=MASK_NONE, fused=True):
    """soa = (x, y, z, vx, vy, vz) when fused else (x, y, d).  Returns (maps, cube) accumulators
    shaped [view][ny][nx][4] and [view][nx][ny][nv] (float64 or int64)."""
    arrs = [_c(a) for a in soa]
    x = arrs[0]
    if fused:
        X, Y, Z, VX, VY, VZ = arrs
    else:
        X, Y, VX = arrs
        Z = VY = VZ = None
    w = _c(w, x.dtype) if w is not None else None
    mask = _c(mask, np.uint8) if mask is not None else None
    ex, ey = _c(ex, np.float64), _c(ey, np.float64)
    ev = _c(ev, np.float64) if ev is not None else None
    nx, ny, nv = ex.size - 1, ey.size - 1, (ev.size - 1 if ev is not None else 0)
    dt = np.float64 if acc_mode == ACC_F64 else np.int64
    maps = np.zeros((nviews, ny, nx, 4), dtype=dt)
    cube = np.zeros((nviews, nx, ny, max(nv, 1) if sums & SUM_CUBE else 0), dtype=dt)
    m = _c(mats, np.float32).reshape(-1) if mats is not None else np.eye(3, dtype=np.float32).reshape(-1)
    sc = _c(scales, np.float64) if scales is not None else None
    lib().pnmo_project_bin(1 if fused else 0, x.size, _dt(x), _p(X), _p(Y), _p(Z), _p(VX), _p(VY), _p(VZ), _p(w),
                           _p(mask), mask_mode, _p(m), nchain, nviews, nx, _p(ex), ny, _p(ey), nv, _p(ev),
                           sums, acc_mode, _p(sc), _p(maps), _p(cube))
    return maps, cube


def finalize_vmaps(maps, cube, sums, acc_mode=ACC_F64, scales=None):
    """maps [ny][nx][4], cube [nx][ny][nv] (one view) -> M, V, S (ny, nx)."""
    ny, nx = maps.shape[0], maps.shape[1]
    nv = cube.shape[2] if cube is not None and cube.size else 0
    M, V, S = (np.empty((ny, nx), dtype=np.float64) for _ in range(3))
    sc = _c(scales, np.float64) if scales is not None else None
    lib().pnmo_finalize_vmaps(_p(_c(maps)), _p(_c(cube)) if nv else None, nx, ny, nv, sums, acc_mode, _p(sc),
                              _p(M), _p(V), _p(S))
    return M, V, S


def smooth_cube(cube, ker2d):
    """Direct 2-D convolution of every velocity slice; ker2d[a, b]: a along cube axis 0."""
    cube = _c(cube, np.float64)
    ker = _c(ker2d, np.float64)
    out = np.empty_like(cube)
    lib().pnmo_smooth_cube(_p(cube), _p(out), cube.shape[0], cube.shape[1], cube.shape[2], _p(ker),
                           ker.shape[0], ker.shape[1])
    return out


def rebin_cube(cube, factorx, factory, mean=False):
    cube = _c(cube, np.float64)
    nx, ny, nv = cube.shape
    out = np.empty((nx // factorx, ny // factory, nv), dtype=np.float64)
    L = lib()
    L.pnmo_rebin_cube.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]
    L.pnmo_rebin_cube.restype = None
    L.pnmo_rebin_cube(_p(cube), _p(out), nx, ny, nv, int(factorx), int(factory), 1 if mean else 0)
    return out


def cube_moments(cube, binv):
    cube = _c(cube, np.float64)
    binv = _c(binv, np.float64)
    nx, ny, nv = cube.shape
    m0, m1, m2 = (np.empty((nx, ny), dtype=np.float64) for _ in range(3))
    lib().pnmo_cube_moments(_p(cube), nx, ny, nv, _p(binv), _p(m0), _p(m1), _p(m2))
    return m0, m1, m2
