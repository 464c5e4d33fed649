"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference under an import shim.

This module exists so that ``oracle/make_golden.py`` (and the container-only
cross-checks in ``tests/test_oracle_vs_reference.py``) can execute the reference's
own ``pynmap/rotation.py``, ``pynmap/grid.py`` and ``pynmap/losvd.py`` in THIS
container.  ``/root/reference`` does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this file.

Why a shim is needed:
  * the reference uses ``np.int`` (``pynmap/grid.py:104,180,234``), removed in numpy 1.24;
  * the reference imports astropy (``pynmap/pynmap.py:16-17``, ``pynmap/losvd.py:13-14``,
    ``pynmap/misc_functions.py:12``), which is not installed and cannot be fetched.

The stub ``astropy`` registered here provides only the four names the hot path
touches.  ``Gaussian2DKernel`` / ``convolve`` are a restatement of astropy's
*documented* behaviour (kernel support ceil(8*max sigma) bumped to odd, sampled at
pixel centres, normalised; ``convolve`` = direct correlation-free 'same' convolution,
boundary='fill', fill_value=0) -- astropy itself is absent, so the smoothing parity
is "unpinned" (see DESIGN.md).
"""
import importlib
import math
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PNM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pynmap", "grid.py"))


class _Gaussian2DKernel(object):
    """astropy.convolution.Gaussian2DKernel as documented: array[y, x], square
    support ceil(8*max(sx, sy)) made odd, evaluated at integer offsets from the
    centre (mode='center'), then normalised to unit sum."""

    def __init__(self, x_stddev, y_stddev=None, theta=0.0, x_size=None, y_size=None):
        if y_stddev is None:
            y_stddev = x_stddev
        default = int(math.ceil(8 * max(x_stddev, y_stddev)))
        default += 1 - default % 2
        nx = default if x_size is None else int(x_size) + 1 - int(x_size) % 2
        ny = default if y_size is None else int(y_size) + 1 - int(y_size) % 2
        xs = np.arange(nx, dtype=np.float64) - (nx - 1) // 2
        ys = np.arange(ny, dtype=np.float64) - (ny - 1) // 2
        gx = np.exp(-0.5 * (xs / x_stddev) ** 2)
        gy = np.exp(-0.5 * (ys / y_stddev) ** 2)
        arr = np.outer(gy, gx) / (2.0 * np.pi * x_stddev * y_stddev)
        self.array = arr / arr.sum()


def _convolve(array, kernel, boundary="fill", fill_value=0.0, **_):
    from scipy.signal import convolve2d
    karr = kernel.array if hasattr(kernel, "array") else np.asarray(kernel)
    karr = karr / karr.sum()
    return convolve2d(np.asarray(array, dtype=np.float64), karr, mode="same",
                      boundary="fill", fillvalue=fill_value)


class _Table(dict):
    @classmethod
    def read(cls, filename, format=None, names=None, **_):
        """Stand-in for astropy's Table.read(format='ascii', names=...).  Format guessing validates the first line as
        a header BEFORE ``names`` is applied (astropy.io.ascii.core.BaseHeader.check_column_names with strict_names:
        a token that looks like a number is not a column name), so a header-less numeric table such as the SSP files
        (misc_functions.py:223-226, :242-249; first line 'Kb 1.30 -2.2653 ...') falls through to the no_header reader
        and all 636 rows are data; a first line of words only is a header.  (astropy is absent here; this restates
        its documented guessing rule.)"""
        def is_number(tok):
            try:
                float(tok)
                return True
            except ValueError:
                return False
        with open(filename) as fh:
            first = fh.readline().lstrip("#").split()
        has_header = bool(first) and not any(is_number(t) for t in first)
        header = list(names) if names is not None else (first if has_header else ["col%d" % (i + 1) for i in range(len(first))])
        data = np.loadtxt(filename, skiprows=1 if has_header else 0, dtype=str)
        out = cls()
        for i, name in enumerate(header):
            col = data[:, i]
            try:
                out[name] = col.astype(np.float64)
            except ValueError:
                out[name] = col
        return out


def _install_stub_astropy():
    if "astropy" in sys.modules and not getattr(sys.modules["astropy"], "_pnm_stub", False):
        return  # a real astropy is present: use it
    astropy = types.ModuleType("astropy")
    astropy._pnm_stub = True
    stats = types.ModuleType("astropy.stats")
    stats.gaussian_fwhm_to_sigma = 1.0 / (2.0 * math.sqrt(2.0 * math.log(2.0)))
    conv = types.ModuleType("astropy.convolution")
    conv.Gaussian2DKernel = _Gaussian2DKernel
    conv.convolve = _convolve
    conv.convolve_fft = _convolve
    table = types.ModuleType("astropy.table")
    table.Table = _Table
    astropy.stats, astropy.convolution, astropy.table = stats, conv, table
    sys.modules.update({"astropy": astropy, "astropy.stats": stats,
                        "astropy.convolution": conv, "astropy.table": table})


def load_reference():
    """Import and return the reference ``pynmap`` package (unmodified sources)."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference mount is read-only
    if not hasattr(np, "int"):
        np.int = int
    try:
        import astropy  # noqa: F401
    except ImportError:
        _install_stub_astropy()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return importlib.import_module("pynmap")
