"""End-to-end path for particles that live in HOST memory (numpy arrays or pinned torch CPU
tensors): what a pynmap user has right after ``snapshot.read`` (pynmap/pynmap.py:182-232).

``project_host`` is rotate -> velocity_moments -> create_losvds (-> convolve_spatial) with host
buffers in and out.  Single GPU: one call of the C ABI's ``pnm_project_host``.  Sharded: every
rank streams its slice with ``pnm_accumulate_host``, the partial grids are summed by the reducer
and the holder rank finalises, smooths and copies back.
"""
import ctypes

import numpy as np
import torch

from . import _lib, engine
from .rotation import rotation_matrix


def _host_ptr(a):
    if isinstance(a, torch.Tensor):
        if a.is_cuda or not a.is_contiguous():
            raise ValueError("host path wants contiguous CPU tensors")
        return ctypes.c_void_p(a.data_ptr())
    if not a.flags.c_contiguous:
        raise ValueError("host path wants contiguous numpy arrays")
    return a.ctypes.data_as(ctypes.c_void_p)


def _dtype_code(a):
    dt = a.dtype
    if dt in (torch.float32, np.float32):
        return _lib.F32
    if dt in (torch.float64, np.float64):
        return _lib.F64
    raise TypeError("host path supports float32 / float64 arrays, got %s" % (dt,))


def smoothing_taps(fwhmx, fwhmy, stepx, stepy):
    """Taps of LOSVD.convolve_spatial starting from FWHM 0 (losvd.py:103-128, axis quirk included)."""
    x_std = fwhmx * engine.FWHM_TO_SIGMA / stepx
    y_std = fwhmy * engine.FWHM_TO_SIGMA / stepy
    support = engine.gaussian_support(x_std, y_std)
    return engine.gaussian_taps(y_std, support), engine.gaussian_taps(x_std, support)


def project_host(arrays, PA, inclination, limXY, nXY, limV=None, nV=0, accumulate="f64", scales=None,
                 smooth_fwhm=None, reducer=None, out=None, device=None):
    """arrays = (x, y, z, vx, vy, vz, w-or-None) on the host, one dtype.  Returns a dict with
    float64 host arrays M, V, S (nY, nX) and, when nV > 0, cube (nX, nY, nV) -- or None on ranks
    that do not hold the reduced result.  ``out`` may provide preallocated (pinned) outputs."""
    # argument checks first (they need no device): the C ABI only sees pointers, so a shorter or differently typed
    # array would be read out of bounds
    if len(arrays) != 7 or any(a is None for a in arrays[:6]):
        raise ValueError("host path wants (x, y, z, vx, vy, vz, w-or-None)")
    x = arrays[0]
    n = int(x.shape[0])
    code = _dtype_code(x)
    for a in arrays:
        if a is not None and (a.dtype != x.dtype or tuple(a.shape) != (n,)):
            raise ValueError("host path: every array must be 1-D with %d entries of dtype %s, got shape %s dtype %s"
                             % (n, x.dtype, tuple(a.shape), a.dtype))
    acc = engine.AccMode(accumulate)
    if acc.code == _lib.ACC_FIXED and scales is None:
        raise ValueError("fixed-point host path needs explicit scales (engine.fixed_scales)")
    n2 = engine.expand_nxy(nXY)
    nx, ny, nv = n2[0], n2[1], int(nV or 0)
    _lib.require_device()
    device = torch.cuda.current_device() if device is None else int(device)
    mat = np.ascontiguousarray(rotation_matrix(np.float32(PA), np.float32(inclination)).reshape(9))
    matp = mat.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    ex = engine.bin_edges(limXY[0], limXY[1], nx)
    ey = engine.bin_edges(limXY[2], limXY[3], ny)
    ev = engine.bin_edges(limV[0], limV[1], nv) if nv else None
    taps0 = taps1 = None
    if smooth_fwhm is not None and nv:
        stepx = np.average(np.diff(np.linspace(limXY[0], limXY[1], nx)))
        stepy = np.average(np.diff(np.linspace(limXY[2], limXY[3], ny)))
        taps0, taps1 = smoothing_taps(smooth_fwhm[0], smooth_fwhm[1], stepx, stepy)
    ptrs = [_host_ptr(a) if a is not None else None for a in arrays]
    sc = _lib.doubles(scales) if scales is not None else None

    out = out or {}
    res = {k: out.get(k) for k in ("M", "V", "S", "cube")}
    for k in ("M", "V", "S"):
        if res[k] is None:
            res[k] = np.empty((ny, nx), dtype=np.float64)
    if nv and res["cube"] is None:
        res["cube"] = np.empty((nx, ny, nv), dtype=np.float64)

    if reducer is None:
        _lib.call("pnm_project_host", device, n, code, *ptrs, matp, 1,
                  nx, _lib.doubles(ex), ny, _lib.doubles(ey), nv, _lib.doubles(ev) if nv else None,
                  acc.code, sc,
                  _lib.doubles(taps0) if taps0 is not None else None, len(taps0) if taps0 is not None else 0,
                  _lib.doubles(taps1) if taps1 is not None else None, len(taps1) if taps1 is not None else 0,
                  _host_ptr(res["M"]), _host_ptr(res["V"]), _host_ptr(res["S"]),
                  _host_ptr(res["cube"]) if nv else None)
        return res

    grid = engine.get_grid(limXY, nx, ny, limV, nv)
    sums = _lib.SUM_W | _lib.SUM_WD | _lib.SUM_WD2 | ((_lib.SUM_CUBE | _lib.W_FROM_CUBE | _lib.CUBE_MOMENTS) if nv else 0)
    maps, cube = engine.new_accumulators(grid, sums, acc)
    _lib.call("pnm_accumulate_host", grid.handle, device, n, code, *ptrs, matp, 1, sums, acc.code, sc,
              engine.ptr(maps), engine.ptr(cube), engine.stream_ptr())
    if not sums & _lib.CUBE_MOMENTS:
        engine.fold_maps(grid, maps, acc, clear=False)
    if not reducer(*engine.reduce_span(grid, maps, cube)):
        return None
    M, V, S = engine.finalize_vmaps(grid, maps, cube, sums | _lib.MAPS_FOLDED, acc, scales)
    for k, t in (("M", M), ("V", V), ("S", S)):
        _copy_out(res[k], t)
    if nv:
        c = engine.finalize_cube(grid, cube, acc, scales, sums)
        if taps0 is not None:
            c = engine.smooth_cube(c, taps0, taps1)
        _copy_out(res["cube"], c)
    torch.cuda.current_stream().synchronize()
    return res


def _copy_out(dst, src):
    if isinstance(dst, torch.Tensor):
        dst.copy_(src.view(dst.shape), non_blocking=True)
    else:
        dst[...] = src.cpu().numpy().reshape(dst.shape)
