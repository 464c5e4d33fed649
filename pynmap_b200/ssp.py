"""SSP mass-to-light weights on the GPU.

The reference turns per-particle (age, [M/H]) into M/L with
``scipy.interpolate.griddata(..., method='cubic')`` -- a Clough-Tocher C1 cubic on the Delaunay
triangulation of the 636-node E-MILES table -- and fills points outside the convex hull with the
nearest node (``pynmap/misc_functions.py:115-207``, interpolation at ``:195-202``); then
``flux = mass / ML`` (``pynmap/pynmap.py:234-236``).

Split used here: everything that depends only on the TABLE stays on the host and runs once
(scipy's own Delaunay triangulation and gradient estimation, then the 19 Bezier coefficients of
every triangle's Clough-Tocher macro-element, restated from scipy's ``_clough_tocher_2d_single``);
everything per PARTICLE runs in the ``pnm_ct_interp`` kernel: point location through a uniform
cell -> triangle index, barycentric coordinates, the cubic, and the nearest-node fill.
"""
import os

import numpy as np
import torch

from . import _lib, engine


def clough_tocher_tables(points, values):
    """Host precomputation.  points (n, 2), values (n,) -> dict of float64 / int32 numpy tables:
       xform (ntri, 6)  rows [T00, T01, T10, T11, r0, r1]: b[:2] = T (p - r), b2 = 1 - b0 - b1
       coef  (ntri, 19) Bezier ordinates c3000 c2100 c2010 c2001 c1200 c1101 c1020 c1011 c1002
                        c0300 c0210 c0201 c0120 c0111 c0102 c0030 c0021 c0012 c0003
    """
    from scipy.interpolate import CloughTocher2DInterpolator
    from scipy.spatial import Delaunay
    points = np.ascontiguousarray(points, dtype=np.float64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    tri = Delaunay(points)
    grad = CloughTocher2DInterpolator(tri, values).grad.reshape(-1, 2)      # scipy's global gradient estimate
    S, P, T = tri.simplices, tri.points, tri.transform
    v1, v2, v3 = P[S[:, 0]], P[S[:, 1]], P[S[:, 2]]
    e12, e23, e31 = v2 - v1, v3 - v2, v1 - v3
    f1, f2, f3 = values[S[:, 0]], values[S[:, 1]], values[S[:, 2]]
    d1, d2, d3 = grad[S[:, 0]], grad[S[:, 1]], grad[S[:, 2]]
    dot = lambda a, b: np.einsum("ij,ij->i", a, b)  # noqa: E731
    df12, df21 = dot(d1, e12), -dot(d2, e12)
    df23, df32 = dot(d2, e23), -dot(d3, e23)
    df31, df13 = dot(d3, e31), -dot(d1, e31)
    c3000 = f1; c2100 = (df12 + 3 * c3000) / 3; c2010 = (df13 + 3 * c3000) / 3
    c0300 = f2; c1200 = (df21 + 3 * c0300) / 3; c0210 = (df23 + 3 * c0300) / 3
    c0030 = f3; c1020 = (df31 + 3 * c0030) / 3; c0120 = (df32 + 3 * c0030) / 3
    c2001 = (c2100 + c2010 + c3000) / 3
    c0201 = (c1200 + c0300 + c0210) / 3
    c0021 = (c1020 + c0120 + c0030) / 3
    # g_k: position of the neighbour's centroid across edge k, in this triangle's barycentric frame
    g = np.full((len(S), 3), -0.5)
    for k in range(3):
        nb = tri.neighbors[:, k]
        has = nb >= 0
        y = P[S[nb[has]]].sum(axis=1) / 3.0
        Tk = T[has]
        c01 = np.einsum("ijk,ik->ij", Tk[:, :2, :], y - Tk[:, 2, :])
        c = np.concatenate([c01, 1.0 - c01.sum(axis=1, keepdims=True)], axis=1)
        a, b = ((2, 1), (0, 2), (1, 0))[k]
        g[has, k] = (2 * c[:, a] + c[:, b] - 1) / (2 - 3 * c[:, a] - 3 * c[:, b])
    c0111 = (g[:, 0] * (-c0300 + 3 * c0210 - 3 * c0120 + c0030) + (-c0300 + 2 * c0210 - c0120 + c0021 + c0201)) / 2
    c1011 = (g[:, 1] * (-c0030 + 3 * c1020 - 3 * c2010 + c3000) + (-c0030 + 2 * c1020 - c2010 + c2001 + c0021)) / 2
    c1101 = (g[:, 2] * (-c3000 + 3 * c2100 - 3 * c1200 + c0300) + (-c3000 + 2 * c2100 - c1200 + c2001 + c0201)) / 2
    c1002 = (c1101 + c1011 + c2001) / 3
    c0102 = (c1101 + c0111 + c0201) / 3
    c0012 = (c1011 + c0111 + c0021) / 3
    c0003 = (c1002 + c0102 + c0012) / 3
    coef = np.stack([c3000, c2100, c2010, c2001, c1200, c1101, c1020, c1011, c1002, c0300, c0210, c0201, c0120,
                     c0111, c0102, c0030, c0021, c0012, c0003], axis=1)
    xform = np.concatenate([T[:, :2, :].reshape(-1, 4), T[:, 2, :]], axis=1)
    return {"xform": np.ascontiguousarray(xform), "coef": np.ascontiguousarray(coef),
            "simplices": S.astype(np.int32), "points": P, "values": values}


def cell_index(points, simplices, ncell=64):
    """Uniform ncell x ncell grid over the bounding box; CSR list of the triangles whose bounding
    box touches each cell (a superset of the triangles that can contain a point of the cell)."""
    lo, hi = points.min(axis=0), points.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    inv = ncell / span
    tri_lo = points[simplices].min(axis=1)
    tri_hi = points[simplices].max(axis=1)
    c_lo = np.clip(np.floor((tri_lo - lo) * inv).astype(int) - 1, 0, ncell - 1)     # one cell of slack
    c_hi = np.clip(np.floor((tri_hi - lo) * inv).astype(int) + 1, 0, ncell - 1)
    lists = [[] for _ in range(ncell * ncell)]
    for t in range(len(simplices)):
        for cx in range(c_lo[t, 0], c_hi[t, 0] + 1):
            for cy in range(c_lo[t, 1], c_hi[t, 1] + 1):
                lists[cx * ncell + cy].append(t)
    start = np.zeros(ncell * ncell + 1, dtype=np.int32)
    start[1:] = np.cumsum([len(l) for l in lists])
    tris = np.array([t for l in lists for t in l], dtype=np.int32)
    return {"start": start, "tris": tris, "ncell": ncell, "lo": lo, "inv": inv}


def eval_tables_numpy(tabs, cells, x, y, fill_nearest=True, eps=100 * np.finfo(np.float64).eps):
    """Host statement of what the pnm_ct_interp kernel computes (used by the CPU tests to check the
    TABLES against scipy; the kernel itself is checked against scipy on the GPU)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    out = np.full(x.shape, np.nan)
    n = cells["ncell"]
    for i in range(x.size):
        cx = int(np.floor((x[i] - cells["lo"][0]) * cells["inv"][0]))
        cy = int(np.floor((y[i] - cells["lo"][1]) * cells["inv"][1]))
        if 0 <= cx <= n and 0 <= cy <= n:
            cx, cy = min(cx, n - 1), min(cy, n - 1)
            for t in cells["tris"][cells["start"][cx * n + cy]:cells["start"][cx * n + cy + 1]]:
                T = tabs["xform"][t]
                dx, dy = x[i] - T[4], y[i] - T[5]
                b0, b1 = T[0] * dx + T[1] * dy, T[2] * dx + T[3] * dy
                b2 = 1.0 - b0 - b1
                if min(b0, b1, b2) >= -eps:
                    out[i] = bezier(tabs["coef"][t], b0, b1, b2)
                    break
    if fill_nearest:
        bad = np.isnan(out)
        if bad.any():
            P = tabs["points"]
            d2 = (x[bad, None] - P[None, :, 0]) ** 2 + (y[bad, None] - P[None, :, 1]) ** 2
            out[bad] = tabs["values"][np.argmin(d2, axis=1)]
    return out


def bezier(c, b0, b1, b2):
    """The cubic of scipy's _clough_tocher_2d_single in extended barycentric coordinates."""
    mn = min(b0, b1, b2)
    b1_, b2_, b3_, b4_ = b0 - mn, b1 - mn, b2 - mn, 3 * mn
    (c3000, c2100, c2010, c2001, c1200, c1101, c1020, c1011, c1002, c0300, c0210, c0201, c0120, c0111, c0102,
     c0030, c0021, c0012, c0003) = c
    return (b1_ ** 3 * c3000 + 3 * b1_ ** 2 * b2_ * c2100 + 3 * b1_ ** 2 * b3_ * c2010 + 3 * b1_ ** 2 * b4_ * c2001
            + 3 * b1_ * b2_ ** 2 * c1200 + 6 * b1_ * b2_ * b4_ * c1101 + 3 * b1_ * b3_ ** 2 * c1020
            + 6 * b1_ * b3_ * b4_ * c1011 + 3 * b1_ * b4_ ** 2 * c1002 + b2_ ** 3 * c0300 + 3 * b2_ ** 2 * b3_ * c0210
            + 3 * b2_ ** 2 * b4_ * c0201 + 3 * b2_ * b3_ ** 2 * c0120 + 6 * b2_ * b3_ * b4_ * c0111
            + 3 * b2_ * b4_ ** 2 * c0102 + b3_ ** 3 * c0030 + 3 * b3_ ** 2 * b4_ * c0021 + 3 * b3_ * b4_ ** 2 * c0012
            + b4_ ** 3 * c0003)


class MLInterpolator(object):
    """griddata(method='cubic') + nearest fill of a scattered (age, [M/H]) -> M/L table, evaluated
    per particle on the GPU.  Build once per table, call per snapshot."""

    def __init__(self, ages, metals, ml, ncell=256, device=None):
        # ncell: 64 -> 13.6 ms, 128 -> 10.3, 256 -> 8.9, 512 -> 8.2 ms per 1e8 particles on a B200 (9.5 ->
        # 2.7 candidate triangles per cell); the reference's scipy path takes ~1 s per 1e6 particles
        pts = np.stack([np.asarray(ages, dtype=np.float64), np.asarray(metals, dtype=np.float64)], axis=1)
        self.tabs = clough_tocher_tables(pts, ml)
        self.cells = cell_index(self.tabs["points"], self.tabs["simplices"], ncell)
        self.device = device or engine.current_device()
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=self.device, dtype=dt)  # noqa: E731
        self.d_xform = up(self.tabs["xform"], torch.float64)
        self.d_coef = up(self.tabs["coef"], torch.float64)
        self.d_start = up(self.cells["start"], torch.int32)
        self.d_tris = up(self.cells["tris"], torch.int32)
        self.d_nodes = up(np.concatenate([self.tabs["points"], self.tabs["values"][:, None]], axis=1), torch.float64)

    def __call__(self, age, MH, fill_nan=True):
        """age, MH: numpy arrays or tensors (float32 / float64) -> float64 M/L, same kind as the input."""
        want_torch = engine.is_cuda_tensor(age)
        dtype = engine.result_dtype(age, MH)
        a = engine.to_device(age, dtype, self.device).reshape(-1)
        m = engine.to_device(MH, dtype, self.device).reshape(-1)
        if a.numel() != m.numel():
            raise ValueError("age and MH must have the same length")
        out = torch.empty(a.numel(), dtype=torch.float64, device=self.device)
        c = self.cells
        _lib.call("pnm_ct_interp", engine.ptr(a), engine.ptr(m), a.numel(), engine._TORCH_TO_PNM[dtype],
                  engine.ptr(self.d_xform), engine.ptr(self.d_coef), int(self.d_xform.shape[0]),
                  engine.ptr(self.d_start), engine.ptr(self.d_tris), int(c["ncell"]),
                  float(c["lo"][0]), float(c["lo"][1]), float(c["inv"][0]), float(c["inv"][1]),
                  engine.ptr(self.d_nodes), int(self.d_nodes.shape[0]), 1 if fill_nan else 0,
                  engine.ptr(out), engine.stream_ptr())
        return engine.to_output(out, want_torch)


# ----------------------------------------------------------------------------- E-MILES tables
VEGA_BANDS = ['U', 'B', 'V', 'R', 'I', 'J', 'H', 'K']
VEGA_SUN_MAG = [5.600, 5.441, 4.820, 4.459, 4.148, 3.711, 3.392, 3.334]       # misc_functions.py:170-171
MASS_COLS = ['IMF', 'slope', 'MH', 'age', 'mTot', 'mSRem', 'mS', 'mRem', 'mGas', 'mlSRemV', 'mlSV', 'mV', 'unk']
PHOT_COLS = ['IMF', 'slope', 'MH', 'age', 'U', 'B', 'V', 'R', 'I', 'J', 'H', 'K']   # first 12 of 34 columns


def _looks_like_number(tok):
    try:
        float(tok)
        return True
    except ValueError:
        return False


def _read_table(path, names):
    """Whitespace ascii table as astropy's Table.read(format='ascii', names=...) reads it (misc_functions.py:223-226,
    :242-249).  Format guessing checks the FIRST line as a header before ``names`` is applied and rejects it when a
    token looks like a number (astropy.io.ascii.core.BaseHeader.check_column_names with strict_names), falling
    through to the no_header reader: the E-MILES tables have no header line ('Kb 1.30 -2.2653 ...'), so all 636 rows
    are data.  A first line made of words only is taken as a header and skipped."""
    with open(path) as fh:
        first = fh.readline().split()
    header = bool(first) and not any(_looks_like_number(t) for t in first)
    raw = np.loadtxt(path, dtype=str, comments=None, skiprows=1 if header else 0)
    return {name: (raw[:, i].astype(np.float64) if i > 0 else raw[:, i]) for i, name in enumerate(names) if i < raw.shape[1]}


def ssp_ml_table(ssp_dir, IMF='KB', slope=1.30, iso='BaSTI', band='V'):
    """(ages, metals, M/L) nodes exactly as misc_functions.compute_ml_ssps builds them
    (misc_functions.py:153-193) from the E-MILES files in ``ssp_dir`` (the reference's
    pynmap/SSPLibraries directory; the tables are data of the reference and are not shipped here)."""
    import os
    iso = {'padova': 'PADOVA00', 'basti': 'BASTI'}[iso.lower()]
    if band not in VEGA_BANDS:
        raise AssertionError("Photometric band '{}' is not in the EMILES SSP predictions".format(band))
    sun = VEGA_SUN_MAG[VEGA_BANDS.index(band)]
    m = _read_table(os.path.join(ssp_dir, "out_mass_{0}_{1}.txt".format(IMF, iso)), MASS_COLS)
    l = _read_table(os.path.join(ssp_dir, "out_phot_{0}_{1}.txt".format(IMF, iso)), PHOT_COLS)
    mw = np.isclose(slope, m['slope'], atol=1e-2)
    lw = np.isclose(slope, l['slope'], atol=1e-2)
    flux = 10 ** (-0.4 * (l[band][lw] - sun))
    return m['age'][mw], m['MH'][mw], m['mSRem'][mw] / flux


def default_ssp_dir():
    """Where the E-MILES tables are looked for when ``ssp_dir`` is not given: $PNM_SSP_DIR, else the
    ``SSPLibraries`` directory of an installed reference package (located, not imported)."""
    env = os.environ.get("PNM_SSP_DIR")
    if env:
        return env
    try:
        import importlib.util
        spec = importlib.util.find_spec("pynmap")
    except (ImportError, ValueError):
        spec = None
    if spec is not None and spec.submodule_search_locations:
        for loc in spec.submodule_search_locations:
            cand = os.path.join(loc, "SSPLibraries")
            if os.path.isdir(cand):
                return cand
    return None


def compute_ml_ssps(mass, age=None, MH=None, ssp_dir=None, IMF='KB', slope=1.30, iso='BaSTI', band='V',
                    fill_nan=True):
    """Drop-in for misc_functions.compute_ml_ssps (:115-207) with the interpolation on the GPU.
    Usable as ``snapshot(..., ml_function=functools.partial(compute_ml_ssps, ssp_dir=...))``."""
    if age is None or MH is None:
        return torch.ones_like(mass) if isinstance(mass, torch.Tensor) else np.ones_like(mass)
    if ssp_dir is None:
        ssp_dir = default_ssp_dir()
    if ssp_dir is None:
        raise ValueError("ssp_dir: path of the reference's pynmap/SSPLibraries directory is required "
                         "(or set PNM_SSP_DIR / install the reference package)")
    ages, metals, ml = ssp_ml_table(ssp_dir, IMF, slope, iso, band)
    return MLInterpolator(ages, metals, ml)(age, MH, fill_nan=fill_nan)
