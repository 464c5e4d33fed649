"""Gauss-Hermite fitting of LOSVDs with the reference's call surface (pynmap/fit_losvd.py:57-235),
batched on the GPU: one warp per profile (csrc/pnm_fit.cuh) instead of one mpfit run per spaxel.

What is kept from the reference: the objective (amplitude re-solved at each evaluation,
fit_losvd.py:25-54 and :178-185; ``gausshermite`` of misc_functions.py:15-68), the start values and
box limits (fit_losvd.py:157-165), the tolerances (ftol = xtol = gtol = 1e-10, :119-121), mpfit's
status codes, and the outputs.  The optimiser is a bounded Levenberg-Marquardt with an analytic
Jacobian, so fitted parameters agree with mpfit's to its own convergence tolerance, not bitwise.
"""
import numpy as np
import torch

from . import engine

OUTSIDE_LIMITS_MSG = 'ERROR: parameters are not within PARINFO limits'     # utils/mpfit.py:957


class FitResult(object):
    """The fields of an mpfit result that the reference reads (fit_losvd.py:204-233)."""

    def __init__(self, params, status, niter, fnorm):
        self.params = params
        self.status = int(status)
        self.niter = int(niter)
        self.fnorm = float(fnorm)
        self.errmsg = OUTSIDE_LIMITS_MSG if self.status == 0 else ''

    def __repr__(self):
        return "FitResult(status={}, niter={}, fnorm={:.6g}, params={})".format(
            self.status, self.niter, self.fnorm, np.array2string(np.asarray(self.params), precision=6))


class FitResults(object):
    """``LOSVD.resfit``: one FitResult per spaxel in the reference's loop order (ix outer, iy
    inner; losvd.py:60-67), materialised lazily from the (nx, ny) arrays."""

    def __init__(self, gh, status, niter, fnorm):
        self.gh, self.status, self.niter, self.fnorm = gh, status, niter, fnorm

    def __len__(self):
        return int(self.status.size)

    def __getitem__(self, i):
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        ix, iy = divmod(i, self.status.shape[1])
        return FitResult(self.gh[1:, ix, iy], self.status[ix, iy], self.niter[ix, iy], self.fnorm[ix, iy])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _set_ghparameters(parlist, default, deg_gh):
    """fit_losvd.py:414-439: a short list is completed with the defaults."""
    if parlist is None:
        parlist = default
    parlist = list(parlist)[:deg_gh]
    return parlist + list(default[len(parlist):])


def _parinfo(xax, deg_gh, params=None, fixed=None, limitedmin=None, limitedmax=None, minpars=None, maxpars=None):
    """The per-parameter start / limits / fixed arrays of fit_losvd.py:157-176 and :203-205."""
    d_start, d_lo, d_hi = engine.fit_gh_defaults(xax, deg_gh)
    start = np.asarray(_set_ghparameters(params, d_start, deg_gh), dtype=np.float64)
    fixed = np.asarray(_set_ghparameters(fixed, [False] * deg_gh, deg_gh), dtype=bool)
    limitedmin = np.asarray(_set_ghparameters(limitedmin, [True] * deg_gh, deg_gh), dtype=bool)
    limitedmax = np.asarray(_set_ghparameters(limitedmax, [True] * deg_gh, deg_gh), dtype=bool)
    lo = np.where(limitedmin, np.asarray(_set_ghparameters(minpars, d_lo, deg_gh), dtype=np.float64), -np.inf)
    hi = np.where(limitedmax, np.asarray(_set_ghparameters(maxpars, d_hi, deg_gh), dtype=np.float64), np.inf)
    return dict(start=start, lo=lo, hi=hi, fixed=fixed)


def _fit_options(kwargs):
    opts = dict(ftol=kwargs.pop("ftol", 1.e-10), xtol=kwargs.pop("xtol", 1.e-10),
                gtol=kwargs.pop("gtol", 1.e-10), maxiter=kwargs.pop("maxiter", 200))
    for name in ("quiet", "verbose", "veryverbose"):      # printing switches of fitgh / mpfit
        kwargs.pop(name, None)
    if kwargs:
        raise TypeError("fitgh: unsupported mpfit keyword(s) {}".format(sorted(kwargs)))
    return opts


def fit_profiles(xax, profiles, deg_gh=4, err=None, want_bestfit=True, params=None, fixed=None,
                 limitedmin=None, limitedmax=None, minpars=None, maxpars=None, **kwargs):
    """Fit all profiles (..., nv) at once with one set of start values / limits (fitgh's keywords);
    tensors stay on the device.  Returns (gh (deg_gh + 1, ...), bestfit, status, niter, chi2)."""
    opts = _fit_options(kwargs)
    opts.update(_parinfo(xax, deg_gh, params, fixed, limitedmin, limitedmax, minpars, maxpars))
    dev = engine.current_device()
    if not isinstance(profiles, torch.Tensor):
        profiles = torch.from_numpy(np.ascontiguousarray(profiles, dtype=np.float64))
    profiles = profiles.to(device=dev, dtype=torch.float64)
    if err is not None:
        if not isinstance(err, torch.Tensor):
            err = torch.from_numpy(np.ascontiguousarray(err, dtype=np.float64))
        err = err.to(device=dev, dtype=torch.float64)
    return engine.fit_gh(profiles, xax, deg_gh=deg_gh, err=err, want_bestfit=want_bestfit, **opts)


def fitgh(xax, data, deg_gh=2, err=None, params=None, fixed=None, limitedmin=None, limitedmax=None,
          minpars=None, maxpars=None, verbose=True, veryverbose=False, **kwargs):
    """One profile (fit_losvd.py:57-82 -> fitgh_mpfit, :84-235).  Returns
    (Ibestpar_array [I, V, S, h3 ..], result, bestfit profile)."""
    optimiser = kwargs.pop("optimiser", "mpfit")
    if optimiser != "mpfit":
        print("ERROR: optimiser must be lmfit or mpfit")      # lmfit is not available on this path
        return None, None, None
    if isinstance(params, np.ndarray):
        params = params.tolist()
    if params is not None and len(params) > deg_gh:           # fit_losvd.py:140-146
        print("ERROR: input parameter array (params) is "
              "larger than expected")
        print("ERROR: It is %d while it should be smaller or "
              "equal to %s (deg_gh)", len(params), deg_gh)
        return 0., 0., 0., 0.
    data = np.asarray(data, dtype=np.float64)
    if xax is None:
        xax = np.arange(len(data))
    xax = np.asarray(xax, dtype=np.float64).ravel()
    gh, best, status, niter, chi2 = fit_profiles(xax, data.reshape(1, -1), deg_gh=deg_gh,
                                                 err=None if err is None else np.asarray(err).reshape(1, -1),
                                                 params=params, fixed=fixed, limitedmin=limitedmin,
                                                 limitedmax=limitedmax, minpars=minpars, maxpars=maxpars, **kwargs)
    gh = gh.cpu().numpy()[:, 0]
    result = FitResult(gh[1:].copy(), status.item(), niter.item(), chi2.item())
    if result.status == 0:
        raise Exception(result.errmsg)                        # fit_losvd.py:210-211
    if verbose:
        print("=====")
        print("FIT: ")
        print("=================================")
        print("        I         V         Sig   ")
        print("   %8.3f  %8.3f   %8.3f " % (gh[0], gh[1], gh[2]))
        print("=================================")
        for i in range(2, deg_gh):
            print("GH %02d: %8.4f " % (i + 1, gh[i + 1]))
        print("=================================")
        print("Chi2: ", result.fnorm, " Reduced Chi2: ", result.fnorm / len(data))
    return gh, result, best.cpu().numpy()[0]


def gausshermite(vbin=None, gh=None, default_nv=101):
    """misc_functions.py:15-68 for callers that plot a fitted profile: evaluated by the same
    device code as the fit (bestfit of a zero-iteration call is not needed; this is cheap numpy-free
    torch arithmetic on the current device)."""
    gh = [float(g) for g in gh]
    if vbin is None:
        vbin = np.linspace(-gh[2] * 5. + gh[1], gh[2] * 5. + gh[1], default_nv)
    degree = len(gh) - 1
    x = torch.as_tensor(np.asarray(vbin, dtype=np.float64), device=engine.current_device())
    if degree < 2:
        print("Error: no enough parameters here")
        return (x * 0.).cpu().numpy()
    if gh[2] == 0.:
        print("Error: Sigma is 0!")
        return (x * 0.).cpu().numpy()
    w = (x - gh[1]) / gh[2]
    w2 = w * w
    h_prev = (2. * w2 - 1.0) / np.sqrt(2.)
    h_cur = (2. * w2 - 3.) * w / np.sqrt(3.)
    h_use = h_cur
    var = torch.ones_like(w)
    for i in range(3, degree + 1):
        var = var + gh[i] * h_use
        h_use = (np.sqrt(2.) * h_cur * w - h_prev) / np.sqrt(i + 1.0)
        h_prev = h_cur
        h_cur = h_use
    return (gh[0] * var * torch.exp(-w2 / 2.)).cpu().numpy()
