"""Gridding layer with the reference's signatures (pynmap/grid.py:70-262):
``points_2vmaps``, ``points_2map``, ``points_2losvds``, ``point_2vprofiles``.  Bodies run on the GPU through
libpnm_b200 (pnm_bin_points + pnm_finalize_*); numpy in -> numpy out, CUDA tensors in ->
CUDA tensors out.  The extra keyword ``accumulate`` ('f64' default, 'fixed' = int64
fixed point) is the only addition to the reference's call surface.
"""
import numpy as np
import torch

from . import _lib, engine
from .losvd import LOSVD


def _prepare(arrays, mask):
    """Common dtype (np.result_type like the reference's `wm * vm`, grid.py:128) and device copies."""
    want_torch = any(engine.is_cuda_tensor(a) for a in arrays if a is not None)
    dtype = engine.result_dtype(*arrays)
    n = None
    dev = []
    for a in arrays:
        t = engine.to_device(a, dtype) if a is not None else None
        if t is not None:
            t = t.reshape(-1)
            if n is None:
                n = t.numel()
            elif t.numel() != n:
                raise ValueError("operands could not be broadcast together: lengths %d and %d" % (n, t.numel()))
        dev.append(t)
    m = None
    if mask is not None:
        m = engine.to_device(mask, torch.uint8).reshape(-1)
        if m.numel() != n:
            raise IndexError("boolean index did not match indexed array: mask has %d entries for %d points"
                             % (m.numel(), n))
    return dev, m, want_torch, dtype


def _value_flags(dtype, d, w):
    """PNM_VAL_D32 / PNM_VAL_W32: numpy's result type for `wm * vm` and `vm**2` (grid.py:128-131) when the
    coordinates forced float64 transport but the values are float32."""
    if dtype != torch.float64:
        return 0
    flags = 0
    if d is not None and engine.result_dtype(d) == torch.float32:
        flags |= _lib.VAL_D32
        if w is not None and engine.result_dtype(w) == torch.float32:
            flags |= _lib.VAL_W32
    return flags


def _scales_for(acc, n, w, d):
    if acc.code != _lib.ACC_FIXED:
        return None
    return engine.fixed_scales(n, engine.absmax(w) if w is not None else 1.0, engine.absmax(d))


def _mesh(grid, want_torch):
    return grid.mesh(want_torch)


def point_2vprofiles(x, y, z, Vx, Vy, Vz, nbins=100, dz=np.array([-0.2, 0.2]), Rmin=1.e-2, Rmax=None, Vmax=None,
                     plot=False, saveplot=False, **kwargs):
    """Mean and dispersion of (V_R, V_theta, V_z) in log-spaced radial bins for particles close to the
    equatorial plane (pynmap/grid.py:16-68).  Returns Rbin (nbins,) float64, v and s (3, nbins) float32;
    empty bins are NaN.  One pass on the GPU (pnm_radial_profiles) instead of a Python loop over bins."""
    if plot:
        raise NameError("name 'plot_sigma' is not defined")        # grid.py:65: the reference's own failure
    mask = kwargs.pop("mask", None)
    (xd, yd, zd, vxd, vyd, vzd), m, want_torch, dtype = _prepare([x, y, z, Vx, Vy, Vz], mask)
    n = xd.numel()
    code = _lib.F32 if dtype == torch.float32 else _lib.F64
    z_lo, z_hi = float(dz[0]), float(dz[1])
    st = engine.stream_ptr()
    if Rmax is None:                                               # R.max() of the selected particles (grid.py:51)
        rmax_d = torch.empty(1, dtype=torch.float64, device=xd.device)
        _lib.call("pnm_profile_rmax", engine.ptr(xd), engine.ptr(yd), engine.ptr(zd), engine.ptr(m), n, code,
                  z_lo, z_hi, engine.ptr(rmax_d), st)
        Rmax = np.float32(rmax_d.item()) if dtype == torch.float32 else np.float64(rmax_d.item())
    Rbin = np.logspace(np.log10(Rmin), np.log10(Rmax), nbins)      # grid.py:52, numpy's dtype rules
    rbin_d = engine.to_device(np.asarray(Rbin, dtype=np.float64), torch.float64, xd.device)
    acc = torch.empty(int(_lib.load().pnm_profile_acc_elems(int(nbins))), dtype=torch.float64, device=xd.device)
    v = torch.empty((3, nbins), dtype=torch.float32, device=xd.device)
    s = torch.empty((3, nbins), dtype=torch.float32, device=xd.device)
    _lib.call("pnm_radial_profiles", engine.ptr(xd), engine.ptr(yd), engine.ptr(zd), engine.ptr(vxd), engine.ptr(vyd),
              engine.ptr(vzd), engine.ptr(m), n, code, z_lo, z_hi, engine.ptr(rbin_d), int(nbins), engine.ptr(acc),
              engine.ptr(v), engine.ptr(s), None, st)
    return Rbin, engine.to_output(v, want_torch), engine.to_output(s, want_torch)


def points_2vmaps(x, y, v, weights=None, limXY=[-1, 1, -1, 1], nXY=11, mask=None, accumulate="f64"):
    """First two velocity moments on a grid (pynmap/grid.py:70-143).

    Returns gridX, gridY, M, V, S, all (nY, nX) float64.  Quirk kept from the reference: a
    mask entry only removes a particle whose ``v`` is also non-finite (grid.py:113-115)."""
    n2 = engine.expand_nxy(nXY)
    if n2 is None:
        print("ERROR: dimension of n should be 1 or 2")
        return 0, 0, 0, 0, 0
    (xd, yd, vd, wd), m, want_torch, dtype = _prepare([x, y, v, weights], mask)
    grid = engine.get_grid(limXY, n2[0], n2[1])
    acc = engine.AccMode(accumulate)
    sums = _lib.SUM_W | _lib.SUM_WD | _lib.SUM_WD2
    scales = _scales_for(acc, xd.numel(), wd, vd)
    maps, _ = engine.new_accumulators(grid, sums, acc)
    engine.bin_points(grid, xd, yd, vd, wd, sums | _value_flags(dtype, v, weights), acc, scales, maps, None, m,
                      _lib.MASK_AND_NONFINITE if m is not None else _lib.MASK_NONE)
    M, V, S = engine.finalize_vmaps(grid, maps, None, sums, acc, scales)
    gx, gy = _mesh(grid, want_torch)
    return gx, gy, engine.to_output(M, want_torch), engine.to_output(V, want_torch), engine.to_output(S, want_torch)


def points_2map(x, y, data=None, weights=None, limXY=[-1, 1, -1, 1], nXY=10, mask=None, accumulate="f64"):
    """Weighted mean of ``data`` on a grid (pynmap/grid.py:145-210) -> gridX, gridY, W, D, DW."""
    n2 = engine.expand_nxy(nXY)
    if n2 is None:
        print("ERROR: dimension of n should be 1 or 2")
        return 0, 0, 0, 0, 0
    if data is None:
        data = torch.ones_like(x) if isinstance(x, torch.Tensor) else np.ones_like(x)
    (xd, yd, dd, wd), m, want_torch, dtype = _prepare([x, y, data, weights], mask)
    grid = engine.get_grid(limXY, n2[0], n2[1])
    acc = engine.AccMode(accumulate)
    sums = _lib.SUM_W | _lib.SUM_WD
    scales = _scales_for(acc, xd.numel(), wd, dd)
    maps, _ = engine.new_accumulators(grid, sums, acc)
    engine.bin_points(grid, xd, yd, dd, wd, sums | _value_flags(dtype, data, weights), acc, scales, maps, None, m,
                      _lib.MASK_AND_NONFINITE if m is not None else _lib.MASK_NONE)
    W, D, DW = engine.finalize_map(grid, maps, acc, scales)
    gx, gy = _mesh(grid, want_torch)
    return gx, gy, engine.to_output(W, want_torch), engine.to_output(D, want_torch), engine.to_output(DW, want_torch)


def points_2losvds(x, y, data, weights=None, limXY=[-1, 1, -1, 1], nXY=10, limV=[-1000, 1000], nV=10,
                   mask=None, accumulate="f64"):
    """(x, y, data) weighted histogram as a LOSVD object (pynmap/grid.py:212-262).
    The cube is indexed [ix, iy, iv] (not transposed, unlike the maps)."""
    n2 = engine.expand_nxy(nXY)
    if n2 is None:
        print("ERROR: dimension of n should be 1 or 2")
        return 0, 0, 0, 0, 0
    (xd, yd, dd, wd), m, want_torch, _ = _prepare([x, y, data, weights], mask)
    if m is not None and wd is not None and bool(m.any()):
        # grid.py:251-259 masks x/y/data but not the weights: numpy then refuses the lengths
        raise ValueError("The weights and list don't have the same length.")
    nV = int(nV)
    grid = engine.get_grid(limXY, n2[0], n2[1], limV, nV)
    acc = engine.AccMode(accumulate)
    scales = _scales_for(acc, xd.numel(), wd, None)
    _, cube = engine.new_accumulators(grid, _lib.SUM_CUBE, acc)
    engine.bin_points(grid, xd, yd, dd, wd, _lib.SUM_CUBE, acc, scales, None, cube, m,
                      _lib.MASK_PLAIN if m is not None else _lib.MASK_NONE)
    out = engine.finalize_cube(grid, cube, acc, scales)
    px, py, pv = grid.centres()
    return LOSVD(px, py, pv, engine.to_output(out, want_torch))
