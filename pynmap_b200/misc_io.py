"""``pynmap.misc_io`` surface (pynmap/misc_io.py) for what sits on or next to the projection path.

  convert_to_list, read_ascii_files     pynmap/misc_io.py:359-368, :143-182 (used by snapshot.read)
  spread_amr                            pynmap/misc_io.py:184-224 -> amr.spread_amr (GPU kernels)
  rebin_1darray, rebin_2darray          pynmap/misc_io.py:370-394 (host helpers on small arrays; whole cubes are
                                        rebinned on the GPU by LOSVD.rebin_spatial)
  return_dummy                          pynmap/misc_io.py:356-357

The RAMSES readers and gas_to_cube (misc_io.py:13-141, :226-354) are I/O for another code path and outside the scope of
this package: calling them raises NotImplementedError naming the reference function.
"""
import numpy as np

from .amr import spread_amr  # noqa: F401
from .pynmap import convert_to_list, read_ascii_files  # noqa: F401


def return_dummy(n=1, default=0):
    return [default] * n


def _fold(a, sh, axes, function):
    if function not in ('sum', 'mean'):
        print("WARNING: doing the sum as input function {}  not recognised".format(function))
        function = 'sum'
    out = np.asarray(a).reshape(sh)
    for ax in axes:                       # the reference chains .sum(-1).sum(1) / .mean(-1).mean(1)
        out = out.mean(ax) if function == 'mean' else out.sum(ax)
    return out


def rebin_1darray(a, shape, function='sum'):
    """``shape`` groups of len(a) // shape consecutive entries, summed or averaged (misc_io.py:370-381)."""
    a = np.asarray(a)
    return _fold(a, (shape, a.shape[0] // shape), (-1,), function)


def rebin_2darray(a, shape, function='sum'):
    """(shape[0], shape[1]) blocks of the 2-D array, summed or averaged: last axis first, then axis 1 of the
    4-D view, like the reference's chained reductions (misc_io.py:383-394)."""
    a = np.asarray(a)
    return _fold(a, (shape[0], a.shape[0] // shape[0], shape[1], a.shape[1] // shape[1]), (-1, 1), function)


def _out_of_scope(name, where):
    def stub(*args, **kwargs):
        raise NotImplementedError("pynmap_b200 does not implement %s (%s): it is outside the particle-projection path "
                                  "this package accelerates; use the reference's pynmap for it" % (name, where))
    stub.__name__ = name
    return stub


read_rdramses_part = _out_of_scope("read_rdramses_part", "pynmap/misc_io.py:13-81")
read_rdramses_hydro = _out_of_scope("read_rdramses_hydro", "pynmap/misc_io.py:83-141")
gas_to_cube = _out_of_scope("gas_to_cube", "pynmap/misc_io.py:226-354")
