"""AMR gas cells: replicate every cell inside its footprint before gridding
(pynmap/misc_io.py:184-224 ``spread_amr``; caller pynmap/pynmap.py:342-419).

The uniform deviates come from numpy's global generator with exactly the reference's draws
(``np.random.random_sample((nsample, n)).ravel()``, once per coordinate, in order), so a run after
``np.random.seed(s)`` reproduces the reference's jittered coordinates bit for bit; the arithmetic
(pnm_amr_jitter / pnm_amr_repeat) runs on the GPU and the replicated arrays never visit the host.
"""
import ctypes

import numpy as np
import torch

from . import _lib, engine

_PNM_DTYPE = {torch.float32: _lib.F32, torch.float64: _lib.F64}


def _as_device(a):
    """Device tensor that keeps the input's dtype (float16/bfloat16 are not reference dtypes)."""
    if isinstance(a, torch.Tensor):
        return a.detach().to(engine.current_device()).contiguous().reshape(-1)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a).reshape(-1))).to(engine.current_device())


def _repeat(values, nsample, divide):
    t = _as_device(values)
    if t.dtype not in _PNM_DTYPE:
        if not divide:
            return torch.repeat_interleave(t, nsample)            # exact copy of an integer / bool array
        t = t.to(torch.float64)                                    # numpy: integer array / int -> float64
    out = torch.empty(t.numel() * nsample, dtype=t.dtype, device=t.device)
    _lib.call("pnm_amr_repeat", engine.ptr(t), _PNM_DTYPE[t.dtype], t.numel(), int(nsample), 1 if divide else 0,
              engine.ptr(out), engine.stream_ptr())
    return out


def spread_amr(list_coord, level, list_val=[], list_const=[], nsample=10, box=10, stretch=[]):
    """Same signature and return value as misc_io.spread_amr: ([coords], [values / nsample],
    [constants]), every array nsample times longer, as CUDA tensors."""
    level_h = level.detach().cpu().numpy() if isinstance(level, torch.Tensor) else np.asarray(level)
    scales = box / 2.0**level_h                                    # misc_io.py:200, numpy's dtype rules
    sizex = len(list_coord[0])
    if stretch == []:
        stretch = [1.] * len(list_coord)
    scale_d = engine.to_device(np.asarray(scales, dtype=np.float64), torch.float64)
    list_cr = []
    for cx, s in zip(list_coord, stretch):
        u = np.random.random_sample((nsample, sizex)).ravel()      # misc_io.py:209
        u_d = engine.to_device(u, torch.float64)
        c = _as_device(cx)
        if c.dtype not in _PNM_DTYPE:
            c = c.to(torch.float64)
        out = torch.empty(sizex * nsample, dtype=torch.float64, device=c.device)
        _lib.call("pnm_amr_jitter", engine.ptr(c), _PNM_DTYPE[c.dtype], engine.ptr(scale_d), engine.ptr(u_d), sizex,
                  int(nsample), ctypes.c_double(float(s)), engine.ptr(out), engine.stream_ptr())
        list_cr.append(out)
    list_valr = [_repeat(val, nsample, True) for val in list_val]
    list_constr = [_repeat(const, nsample, False) for const in list_const]
    return list_cr, list_valr, list_constr
