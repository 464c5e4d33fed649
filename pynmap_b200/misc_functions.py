"""``pynmap.misc_functions`` surface (pynmap/misc_functions.py) for what sits on or next to the projection path.

  gausshermite                pynmap/misc_functions.py:15-68  -> fit_losvd.gausshermite
  moment012                   pynmap/misc_functions.py:70-90  (one profile on the host; whole cubes: LOSVD.vmoments on the GPU)
  compute_ml, compute_ml_ssps pynmap/misc_functions.py:92-207 -> ssp (Clough-Tocher interpolation on the GPU)

The morphology helpers (find_center*, characterise_ima2d, morph_ima2d, ...; misc_functions.py:251-440) are outside the
scope of this package and raise NotImplementedError.
"""
import numpy as np
import torch

from .fit_losvd import gausshermite  # noqa: F401
from . import ssp


def moment012(x, data):
    """m0 = sum, m1 = weighted mean, m2 = weighted standard deviation of one profile; an all-zero profile gives
    (0, 0, min(diff(x)) / 3) (misc_functions.py:70-90)."""
    x, data = np.asarray(x), np.asarray(data)
    if not np.any(data != 0):
        return 0., 0., np.min(np.diff(x)) / 3.0
    m0 = np.sum(data)
    m1 = np.average(x, weights=data)
    return m0, m1, np.sqrt(np.average(x ** 2, weights=data) - m1 ** 2)


def compute_ml_ssps(mass, age=None, MH=None, recipe='EMILES', IMF='KB', slope=1.30, iso='BaSTI', band='V', fill_nan=True,
                    ssp_dir=None):
    """M/L per particle from the E-MILES SSP tables (misc_functions.py:115-207) with the reference's signature;
    ``ssp_dir`` (an addition) says where the reference's SSPLibraries directory is (default: $PNM_SSP_DIR or an
    installed reference).  Any recipe without 'EMILES', or missing age / MH, gives ones, like the reference."""
    if 'EMILES' not in recipe or age is None or MH is None:
        return torch.ones_like(mass) if isinstance(mass, torch.Tensor) else np.ones_like(mass)
    return ssp.compute_ml_ssps(mass, age, MH, ssp_dir=ssp_dir, IMF=IMF, slope=slope, iso=iso, band=band, fill_nan=fill_nan)


def compute_ml(mass, age=None, ZH=None, method="SSPs"):
    """M/L per particle (misc_functions.py:92-113): the SSP interpolation, or ones for any other method."""
    if method == "SSPs":
        return compute_ml_ssps(mass, age, ZH)
    return torch.ones_like(mass) if isinstance(mass, torch.Tensor) else np.ones_like(mass)


def _out_of_scope(name, where):
    def stub(*args, **kwargs):
        raise NotImplementedError("pynmap_b200 does not implement %s (%s): it is outside the particle-projection path "
                                  "this package accelerates; use the reference's pynmap for it" % (name, where))
    stub.__name__ = name
    return stub


for _name, _where in (("SSPMass", "misc_functions.py:209-226"), ("SSPPhot", "misc_functions.py:228-249"),
                      ("find_centerodd", "misc_functions.py:251-272"), ("find_centereven", "misc_functions.py:274-290"),
                      ("xy_cov_matrix2d", "misc_functions.py:292-305"), ("characterise_ima2d", "misc_functions.py:307-354"),
                      ("guess_stepx", "misc_functions.py:356-358"), ("guess_stepxy", "misc_functions.py:360-383"),
                      ("morph_ima2d", "misc_functions.py:385-416"), ("comp_pa_eps", "misc_functions.py:418-440")):
    globals()[_name] = _out_of_scope(_name, "pynmap/" + _where)
