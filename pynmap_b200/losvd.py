"""LOSVD container with the reference's interface (pynmap/losvd.py:16-141).  The cube may be
a numpy array or a CUDA tensor; ``convolve_spatial`` (Gaussian smoothing of every velocity
slice) and ``vmoments`` run on the GPU.
"""
import copy

import numpy as np
import torch

from . import engine


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


class LOSVD(object):
    """Line-of-sight velocity distributions on an (X, Y, V) grid: ``losvd[ix, iy, iv]`` with
    pixel-centre coordinates ``binx``, ``biny``, ``binv`` (pynmap/losvd.py:20-39)."""

    def __init__(self, binx=None, biny=None, binv=None, losvd=None, err_losvd=None,
                 fwhmx=0., fwhmy=0., name='losvd'):
        self.name = name
        self.binx = binx
        self.biny = biny
        self.binv = binv
        self.stepx = np.average(np.diff(_np(self.binx)))
        self.stepy = np.average(np.diff(_np(self.biny)))
        self.losvd = losvd
        self._err_losvd = err_losvd          # None: the reference's (nX, nY) array of None, built on first access
        self.fwhmx = fwhmx
        self.fwhmy = fwhmy

    @property
    def err_losvd(self):
        """Uncertainties of the LOSVDs; an (nX, nY) object array of None when there are none (pynmap/losvd.py:33-35)."""
        if self._err_losvd is None:
            self._err_losvd = np.full(tuple(self.losvd.shape[:2]), None)
        return self._err_losvd

    @err_losvd.setter
    def err_losvd(self, value):
        self._err_losvd = value

    @property
    def extent(self):
        bx, by = _np(self.binx), _np(self.biny)
        return [np.min(bx), np.max(bx), np.min(by), np.max(by)]

    # ------------------------------------------------------------------ device plumbing
    def _cube_on_device(self, cube):
        want_torch = engine.is_cuda_tensor(cube)
        if isinstance(cube, torch.Tensor):
            t = cube.detach().to(device=engine.current_device(), dtype=torch.float64)
        else:
            t = torch.from_numpy(np.ascontiguousarray(cube, dtype=np.float64)).to(engine.current_device())
        return t.contiguous(), want_torch

    def _has_errors(self):
        e = self._err_losvd
        if e is None:
            return False
        if isinstance(e, torch.Tensor):
            return True
        return e[0, 0] is not None

    # ------------------------------------------------------------------ f1: moments
    def vmoments(self, **kwargs):
        """Per-spaxel mom0, mom1, mom2 of the LOSVDs (pynmap/losvd.py:41-51 with
        misc_functions.moment012, :70-90): one warp per spaxel instead of a Python double loop."""
        cube, want_torch = self._cube_on_device(self.losvd)
        m0, m1, m2 = engine.cube_moments(cube, _np(self.binv))
        self.mom0, self.mom1, self.mom2 = (engine.to_output(m, want_torch) for m in (m0, m1, m2))

    # ------------------------------------------------------------------ f4: Gauss-Hermite fit
    def fitlosvd(self, degree=4, on_error="raise", **kwargs):
        """Fit every LOSVD with a Gauss-Hermite series (pynmap/losvd.py:53-67 looping
        fit_losvd.fitgh): all spaxels in one kernel.  Sets ``GH`` (degree + 1, nx, ny) = I, V, S,
        h3.., ``bestfit`` (nx, ny, nv) and ``resfit`` (per-spaxel status / niter / fnorm).  The
        reference allocates ``GH`` with 5 rows, i.e. only degree=4 works there.

        A spaxel whose start values fall outside the limits (e.g. a single occupied channel:
        sigma = 0) makes the reference raise from mpfit; ``on_error="flag"`` marks such spaxels
        with status 0 and NaN parameters instead."""
        from . import fit_losvd
        cube, want_torch = self._cube_on_device(self.losvd)
        err = None
        if self._has_errors():
            err, _ = self._cube_on_device(self.err_losvd)
        gh, best, status, niter, chi2 = fit_losvd.fit_profiles(_np(self.binv), cube, deg_gh=degree, err=err, **kwargs)
        status_h = status.cpu().numpy()
        if on_error == "raise" and np.any(status_h == 0):
            raise Exception(fit_losvd.OUTSIDE_LIMITS_MSG)
        self.GH = engine.to_output(gh, want_torch)
        self.bestfit = engine.to_output(best, want_torch)
        self.resfit = fit_losvd.FitResults(gh.cpu().numpy(), status_h, niter.cpu().numpy(), chi2.cpu().numpy())

    # ------------------------------------------------------------------ f3: spatial rebin
    def rebin_spatial(self, factorx=1, factory=1, function='sum'):
        """Rebin by integer factors in X and Y (pynmap/losvd.py:69-92 with misc_io.rebin_1darray /
        rebin_2darray, misc_io.py:370-394).  Quirk kept: ``binx`` / ``biny`` go through the same
        ``function``, so with the default 'sum' the new coordinates are SUMS of the old centres."""
        nx, ny = int(self.losvd.shape[0]), int(self.losvd.shape[1])
        newshape = (nx // factorx, ny // factory)
        print("Newshape will be {}".format(newshape))
        if function not in ('sum', 'mean'):
            print("WARNING: doing the sum as input function {}  not recognised".format(function))
        mean = function == 'mean'
        if newshape[0] * factorx != nx or newshape[1] * factory != ny:
            # a.reshape(sh) of the reference fails for sizes that are not multiples of the factors
            raise ValueError("cannot reshape array of size {} into shape {}".format(
                nx * ny, (newshape[0], factorx, newshape[1], factory)))

        def rebin1d(a, n, f):
            a = _np(a).reshape((n, f))
            return a.mean(-1) if mean else a.sum(-1)

        binx = rebin1d(self.binx, newshape[0], factorx)
        biny = rebin1d(self.biny, newshape[1], factory)
        cube, want_torch = self._cube_on_device(self.losvd)
        losvd = engine.to_output(engine.rebin_cube(cube, factorx, factory, mean), want_torch)
        err_losvd = None
        if self._has_errors():
            ecube, e_torch = self._cube_on_device(self.err_losvd)
            err_losvd = engine.to_output(engine.rebin_cube(ecube, factorx, factory, mean), e_torch)
        return LOSVD(binx, biny, self.binv, losvd, err_losvd, self.fwhmx, self.fwhmy)

    # ------------------------------------------------------------------ a5: smoothing
    def convolve_spatial(self, fwhmx=None, fwhmy=None):
        """Smooth every velocity slice to the target FWHM (pynmap/losvd.py:94-141).

        The convolution width is the quadrature difference with the current FWHM; the kernel
        is astropy's Gaussian2DKernel (support ceil(8*max sigma) made odd, normalised),
        applied here as two 1-D passes.  Axis quirk kept: the kernel array is indexed [y, x]
        while slice axis 0 is X, so ``fwhmx`` acts along the Y axis of the cube and vice versa."""
        if fwhmx is None and fwhmy is None:
            print("No convolution to be done with both FWHM as None")
            return
        elif fwhmx <= 0. and fwhmy <= 0.:
            print("No convolution to be done with FWHM <= 0")
            return

        if fwhmx > self.fwhmx:
            conv_fwhmx = np.sqrt(fwhmx**2 - self.fwhmx**2)
            reached_fwhmx = fwhmx
        else:
            print("Warning: objective FWHM_X smaller than present value")
            conv_fwhmx = 0.
            reached_fwhmx = self.fwhmx
        if fwhmy > self.fwhmy:
            conv_fwhmy = np.sqrt(fwhmy**2 - self.fwhmy**2)
            reached_fwhmy = fwhmy
        else:
            print("Warning: objective FWHM_Y smaller than present value")
            conv_fwhmy = 0.
            reached_fwhmy = self.fwhmy

        # the reference deep-copies (losvd.py:118): the new object shares nothing with this one -- the coordinate arrays
        # are copied here, the cube and its uncertainties are replaced below
        closvd = copy.copy(self)
        closvd.binx, closvd.biny, closvd.binv = (a.copy() if isinstance(a, np.ndarray) else copy.deepcopy(a)
                                                 for a in (self.binx, self.biny, self.binv))
        print("LOSVD FWHM: XLOS[{0}] YLOS[{1}]".format(self.fwhmx, self.fwhmy))
        print("Starting the Convolution to reach FWHM: XCLOS[{0}] YCLOS[{1}]".format(
              reached_fwhmx, reached_fwhmy))
        print("Convolving with quadratic residuals FWHM: DX[{0}] DY[{1}]".format(conv_fwhmx, conv_fwhmy))

        x_stddev = conv_fwhmx * engine.FWHM_TO_SIGMA / self.stepx
        y_stddev = conv_fwhmy * engine.FWHM_TO_SIGMA / self.stepy
        support = engine.gaussian_support(x_stddev, y_stddev)
        with np.errstate(divide="ignore", invalid="ignore"):
            taps_axis0 = engine.gaussian_taps(y_stddev, support)   # kernel rows (y) meet cube axis 0 (X)
            taps_axis1 = engine.gaussian_taps(x_stddev, support)

        cube, want_torch = self._cube_on_device(self.losvd)
        closvd.losvd = engine.to_output(engine.smooth_cube(cube, taps_axis0, taps_axis1), want_torch)
        if self._has_errors():
            ecube, e_torch = self._cube_on_device(self.err_losvd)
            closvd.err_losvd = engine.to_output(engine.smooth_cube(ecube, taps_axis0, taps_axis1), e_torch)
        else:
            closvd._err_losvd = None                     # a fresh (nX, nY) array of None on first access

        closvd.fwhmx = reached_fwhmx
        closvd.fwhmy = reached_fwhmy
        return closvd
