"""TEST INFRASTRUCTURE ONLY -- numpy restatement of pynmap's particle-projection path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package
``pynmap_b200`` never does (it fails loudly when its CUDA library is missing).

What is restated, and from where (paths relative to /root/reference):

  rotation matrix + application   pynmap/rotation.py:73-119 (rotate_mat), :14-71 (rotateXYZ_mat)
  angle wrapping in float32       pynmap/misc_io.py:359-368 (convert_to_list) via pynmap/pynmap.py:274-282
  bin edges / pixel centres       pynmap/grid.py:119-124, :194-197, :239-249
  moment maps M, V, S             pynmap/grid.py:70-143  (points_2vmaps)
  weighted data map W, D, DW      pynmap/grid.py:145-210 (points_2map)
  LOSVD cube                      pynmap/grid.py:212-262 (points_2losvds)
  spatial Gaussian smoothing      pynmap/losvd.py:94-141 (LOSVD.convolve_spatial)
  per-spaxel velocity moments     pynmap/losvd.py:41-51 + pynmap/misc_functions.py:70-90

The reference delegates its arithmetic to numpy (``np.dot``, ``np.histogram2d``,
``np.histogramdd``) and astropy (``Gaussian2DKernel``, ``convolve``).  numpy is a
third-party dependency the reference does not pin (no install_requires in
setup.py:35-42); this container has numpy 2.3.5.  The histogram semantics below
are restated explicitly (``searchsorted(edges, x, 'right')``, the closed last
edge, outlier stripping and a sequential float64 ``bincount``) following
numpy/lib/_histograms_impl.py:1040-1071 of that version, instead of calling
``np.histogramdd``, so that bin *indices* are observable on their own.

PINNING: ``oracle/make_golden.py`` executes the reference's own code under
``oracle/refshim.py`` in the build container and stores its outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function in
this file against those vectors (bit-exact for rotation, bin indices, counts,
sums).  astropy is absent, so the smoothing is pinned only against the shim's
restatement of astropy's documented behaviour: smoothing parity is "unpinned".
"""
import math

import numpy as np

FWHM_TO_SIGMA = 1.0 / (2.0 * math.sqrt(2.0 * math.log(2.0)))  # astropy.stats.gaussian_fwhm_to_sigma


# --------------------------------------------------------------------------- rotation
def as_f32_angle(angle):
    """snapshot.rotate wraps each angle as np.float32 (misc_io.py:363), which makes the
    trigonometry below run in float32 (pynmap.py:274-278)."""
    return np.float32(angle)


def rotation_matrix(PA=0.0, inclination=0.0, direct=True):
    """float32 3x3 matrix of rotate_mat (rotation.py:99-113).  The dtype of PA /
    inclination decides whether deg2rad/cos/sin run in float32 or float64 before the
    final cast -- the two differ in the last ulp (SURVEY.md 8c KAT 4)."""
    PA = np.deg2rad(PA)
    inclination = np.deg2rad(inclination)
    cPA, sPA = np.cos(PA), np.sin(PA)
    ci, si = np.cos(inclination), np.sin(inclination)
    mat = np.array([[cPA, -sPA, 0], [ci * sPA, ci * cPA, -si], [si * sPA, cPA * si, ci]],
                   dtype=np.float32)
    return mat if direct else np.transpose(mat)


def rotation_matrix_xyz(alpha=0.0, beta=0.0, gamma=0.0, rotorder=(0, 1, 2), direct=True):
    """float32 matrix of rotateXYZ_mat (rotation.py:32-65): product of the three axis
    matrices in ``rotorder``; the products are float32 np.dot."""
    ca, sa = np.cos(np.deg2rad(alpha)), np.sin(np.deg2rad(alpha))
    cb, sb = np.cos(np.deg2rad(beta)), np.sin(np.deg2rad(beta))
    cg, sg = np.cos(np.deg2rad(gamma)), np.sin(np.deg2rad(gamma))
    mx = np.array([[1, 0, 0], [0, ca, sa], [0, -sa, ca]], dtype=np.float32)
    my = np.array([[cb, 0, -sb], [0, 1, 0], [sb, 0, cb]], dtype=np.float32)
    mz = np.array([[cg, sg, 0], [-sg, cg, 0], [0, 0, 1]], dtype=np.float32)
    mats = [mx, my, mz] if direct else [mx.T, my.T, mz.T]
    return np.dot(np.dot(mats[rotorder[0]], mats[rotorder[1]]), mats[rotorder[2]])


def _round_f64_sum_to_f32(p, t):
    """fl32(p + t) for float64 ``p`` (an exact product) and float32 ``t`` with a single
    rounding: TwoSum recovers the float64 rounding error, the float64 sum is then
    forced to an odd significand when inexact (round-to-odd), after which the final
    float64 -> float32 rounding is the correctly rounded result (53 >= 24 + 2 bits)."""
    t = t.astype(np.float64)
    s = p + t
    bb = s - p
    err = (p - (s - bb)) + (t - bb)
    even = (s.view(np.int64) & 1) == 0
    toward = np.where(err > 0, np.inf, -np.inf)
    s_odd = np.where((err != 0) & even, np.nextafter(s, toward), s)
    return s_odd.astype(np.float32)


def _fma_chain_row(m0, m1, m2, x, y, z):
    """fl(fma(m2, z, fma(m1, y, fl(m0 * x)))).

    np.dot(mat(3x3), pos.T(3xN)) rounds like this forward FMA chain for both float32 and
    float64 data (SURVEY.md 8c KAT 7; re-checked against the reference's own output in
    tests/golden/rotate_*.npz).  numpy has no fma: float32 goes through exact float64
    products and a round-to-odd sum; float64 uses exact rational arithmetic per element (small arrays
    only -- oracle/pnm_oracle.c is the fast path)."""
    if x.dtype == np.float32:
        t = (np.float64(m0) * x.astype(np.float64)).astype(np.float32)   # product exact in f64
        t = _round_f64_sum_to_f32(np.float64(m1) * y.astype(np.float64), t)
        return _round_f64_sum_to_f32(np.float64(m2) * z.astype(np.float64), t)
    # float64: exact rational arithmetic per element, one correctly rounded conversion per fma
    from fractions import Fraction
    f0, f1, f2 = Fraction(float(m0)), Fraction(float(m1)), Fraction(float(m2))

    def one(a, b, c):
        if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
            return float(m2) * c + (float(m1) * b + float(m0) * a)
        t = float(m0) * a
        t = float(f1 * Fraction(b) + Fraction(t))
        return float(f2 * Fraction(c) + Fraction(t))

    flat = [one(a, b, c) for a, b, c in zip(x.ravel().tolist(), y.ravel().tolist(), z.ravel().tolist())]
    return np.array(flat, dtype=np.float64).reshape(x.shape)


def rotate_soa(mat, x, y, z):
    """Apply ``mat`` (float32 3x3) to SoA coordinates; result dtype follows
    np.result_type(float32, input) like np.dot (rotation.py:115-117)."""
    mat = np.asarray(mat, dtype=np.float32)
    dt = np.result_type(np.float32, x.dtype)
    x, y, z = (np.ascontiguousarray(a, dtype=dt) for a in (x, y, z))
    return tuple(_fma_chain_row(mat[r, 0], mat[r, 1], mat[r, 2], x, y, z) for r in range(3))


def rotate_mat(pos, vel=None, PA=0.0, inclination=0.0, direct=True):
    """(N,3) in / (N,3) out restatement of rotation.py:73-119."""
    mat = rotation_matrix(PA, inclination, direct)
    pos = np.asarray(pos)
    out = np.stack(rotate_soa(mat, pos[:, 0], pos[:, 1], pos[:, 2]), axis=0).T
    if vel is None:
        return out, None
    vel = np.asarray(vel)
    return out, np.stack(rotate_soa(mat, vel[:, 0], vel[:, 1], vel[:, 2]), axis=0).T


# --------------------------------------------------------------------------- grid
def expand_nxy(nXY):
    """grid.py:103-107: scalar -> (n, n); anything but 1 or 2 entries is the error path."""
    if np.size(nXY) == 1:
        n = int(np.asarray(nXY).reshape(-1)[0])
        return n, n
    if np.size(nXY) != 2:
        return None
    return int(nXY[0]), int(nXY[1])


def bin_edges(lo, hi, n):
    """Half-step padded edges (grid.py:119-124): pixel *centres* are linspace(lo, hi, n)."""
    half = (hi - lo) / (n - 1) / 2.0
    return np.linspace(lo - half, hi + half, n + 1)


def pixel_centres(lo, hi, n):
    return np.linspace(lo, hi, n)


def bin_index(values, edges):
    """numpy histogramdd's per-axis bin rule (numpy/lib/_histograms_impl.py:1040-1053):
    searchsorted(..., 'right'), values equal to the last edge go to the last bin,
    under/overflow and NaN are dropped.  Returns int32 indices in [0, n) or -1."""
    edges = np.asarray(edges, dtype=np.float64)
    n = edges.size - 1
    v = np.asarray(values)
    i = np.searchsorted(edges, v, side="right")
    i[v == edges[-1]] -= 1
    out = (i - 1).astype(np.int32)
    out[(i == 0) | (i == n + 1)] = -1
    return out


def _weighted_hist(flat, keep, weights, size):
    """np.bincount: sequential float64 accumulation in particle order."""
    if weights is None:
        return np.bincount(flat[keep], minlength=size).astype(np.float64)
    # histogramdd returns float64 even for empty input (bincount alone would give intp there)
    return np.bincount(flat[keep], weights=np.asarray(weights)[keep], minlength=size).astype(np.float64)


def hist2d(x, y, ex, ey, weights=None):
    """np.histogram2d(x, y, [ex, ey], weights)[0]  -> shape (nX, nY)."""
    nx, ny = len(ex) - 1, len(ey) - 1
    ix, iy = bin_index(x, ex), bin_index(y, ey)
    keep = (ix >= 0) & (iy >= 0)
    flat = ix.astype(np.int64) * ny + iy
    return _weighted_hist(flat, keep, weights, nx * ny).reshape(nx, ny)


def hist3d(x, y, v, ex, ey, ev, weights=None):
    """np.histogramdd((x,y,v), (ex,ey,ev), weights)[0] -> shape (nX, nY, nV)."""
    nx, ny, nv = len(ex) - 1, len(ey) - 1, len(ev) - 1
    ix, iy, iv = bin_index(x, ex), bin_index(y, ey), bin_index(v, ev)
    keep = (ix >= 0) & (iy >= 0) & (iv >= 0)
    flat = (ix.astype(np.int64) * ny + iy) * nv + iv
    return _weighted_hist(flat, keep, weights, nx * ny * nv).reshape(nx, ny, nv)


def moment_sums(x, y, v, weights=None, limXY=(-1, 1, -1, 1), nXY=11, mask=None):
    """The three raw sums of points_2vmaps before finalisation, each (nY, nX)
    (grid.py:109-131), including the mask quirk: a particle is removed only if it is
    both masked and non-finite in v (grid.py:114-115)."""
    nx, ny = expand_nxy(nXY)
    x, y, v = np.asarray(x), np.asarray(y), np.asarray(v)
    w = np.ones_like(x) if weights is None else np.asarray(weights)
    if mask is None:
        mask = np.zeros_like(w, dtype=bool)
    mask = mask & ~np.isfinite(v)
    xm, ym, vm, wm = x[~mask], y[~mask], v[~mask], w[~mask]
    ex, ey = bin_edges(limXY[0], limXY[1], nx), bin_edges(limXY[2], limXY[3], ny)
    s0 = hist2d(xm, ym, ex, ey, wm).T
    s1 = hist2d(xm, ym, ex, ey, wm * vm).T          # product rounded in the input dtype
    s2 = hist2d(xm, ym, ex, ey, wm * vm ** 2).T     # vm**2 == vm*vm, then times wm
    return s0, s1, s2


def finalise_moments(s0, s1, s2):
    """grid.py:133-138."""
    good = s0 != 0
    V = np.zeros_like(s0)
    S = np.zeros_like(s0)
    V[good] = s1[good] / s0[good]
    with np.errstate(invalid="ignore"):
        S[good] = np.sqrt(s2[good] / s0[good] - V[good] * V[good])
    return V, S


def points_2vmaps(x, y, v, weights=None, limXY=(-1, 1, -1, 1), nXY=11, mask=None):
    """grid.py:70-143 -> gridX, gridY, M, V, S, all (nY, nX) float64."""
    if expand_nxy(nXY) is None:
        return 0, 0, 0, 0, 0
    nx, ny = expand_nxy(nXY)
    s0, s1, s2 = moment_sums(x, y, v, weights, limXY, nXY, mask)
    V, S = finalise_moments(s0, s1, s2)
    gx, gy = np.meshgrid(pixel_centres(limXY[0], limXY[1], nx), pixel_centres(limXY[2], limXY[3], ny))
    return gx, gy, s0, V, S


def points_2map(x, y, data=None, weights=None, limXY=(-1, 1, -1, 1), nXY=10, mask=None):
    """grid.py:145-210 -> gridX, gridY, W, D, DW."""
    if expand_nxy(nXY) is None:
        return 0, 0, 0, 0, 0
    nx, ny = expand_nxy(nXY)
    x, y = np.asarray(x), np.asarray(y)
    if mask is None:
        mask = np.zeros_like(x, dtype=bool)
    w = np.ones_like(x) if weights is None else np.asarray(weights)
    d = np.ones_like(x) if data is None else np.asarray(data)
    mask = mask & ~np.isfinite(d)
    xm, ym, wm, dm = x[~mask], y[~mask], w[~mask], d[~mask]
    ex, ey = bin_edges(limXY[0], limXY[1], nx), bin_edges(limXY[2], limXY[3], ny)
    W = hist2d(xm, ym, ex, ey, wm).T
    DW = hist2d(xm, ym, ex, ey, wm * dm).T
    D = np.zeros_like(W)
    good = W != 0
    D[good] = DW[good] / W[good]
    gx, gy = np.meshgrid(pixel_centres(limXY[0], limXY[1], nx), pixel_centres(limXY[2], limXY[3], ny))
    return gx, gy, W, D, DW


def points_2losvds(x, y, data, weights=None, limXY=(-1, 1, -1, 1), nXY=10,
                   limV=(-1000, 1000), nV=10, mask=None):
    """grid.py:212-262 -> (pixX, pixY, pixV, cube[nX,nY,nV]).  The mask is applied
    to x/y/data only (grid.py:251-254): masked particles together with weights is a
    length mismatch, i.e. ValueError, exactly like the reference."""
    if expand_nxy(nXY) is None:
        return 0, 0, 0, 0, 0
    nx, ny = expand_nxy(nXY)
    x, y, d = np.asarray(x), np.asarray(y), np.asarray(data)
    ex, ey = bin_edges(limXY[0], limXY[1], nx), bin_edges(limXY[2], limXY[3], ny)
    ev = bin_edges(limV[0], limV[1], nV)
    if mask is None:
        mask = np.zeros_like(x, dtype=bool)
    xs, ys, ds = x[~mask], y[~mask], d[~mask]
    # np.hstack of the three columns promotes them to one common dtype (grid.py:253-254)
    dt = np.result_type(xs.dtype, ys.dtype, ds.dtype)
    xs, ys, ds = xs.astype(dt), ys.astype(dt), ds.astype(dt)
    if weights is not None and np.size(weights) != xs.size:
        raise ValueError("The weights and list don't have the same length.")
    cube = hist3d(xs, ys, ds, ex, ey, ev, weights)
    return (pixel_centres(limXY[0], limXY[1], nx), pixel_centres(limXY[2], limXY[3], ny),
            pixel_centres(limV[0], limV[1], nV), cube)


# --------------------------------------------------------------------------- fixed point (O3)
def fixed_quantise(values, scale):
    """int64 fixed-point image of float values: rint(value * scale), value widened to
    float64 first; scale is a power of two so the product is exact."""
    return np.rint(np.asarray(values, dtype=np.float64) * scale).astype(np.int64)


def fixed_hist(flat, keep, values, scale, size):
    """Exact integer accumulation (order independent) of the quantised values."""
    q = fixed_quantise(values, scale)[keep]
    out = np.zeros(size, dtype=np.int64)
    np.add.at(out, flat[keep], q)
    return out


# --------------------------------------------------------------------------- smoothing (O2)
def gaussian_taps(sigma_pix, support):
    """1-D factor of astropy's Gaussian2DKernel sampled at integer offsets."""
    half = (support - 1) // 2
    i = np.arange(-half, half + 1, dtype=np.float64)
    g = np.exp(-0.5 * (i / sigma_pix) ** 2)
    return g / g.sum()


def kernel_support(sig_a, sig_b):
    s = int(math.ceil(8 * max(sig_a, sig_b)))
    return s + 1 - s % 2


def gaussian_kernel2d(x_stddev, y_stddev):
    """Normalised 2-D kernel, array[y, x] (astropy Gaussian2DKernel)."""
    s = kernel_support(x_stddev, y_stddev)
    half = (s - 1) // 2
    i = np.arange(-half, half + 1, dtype=np.float64)
    arr = np.outer(np.exp(-0.5 * (i / y_stddev) ** 2), np.exp(-0.5 * (i / x_stddev) ** 2))
    return arr / arr.sum()


def convolve2d_fill0(img, ker):
    """Direct 'same' convolution with zero fill in float64 -- astropy.convolution.convolve with its defaults
    boundary='fill', fill_value=0, nan_treatment='interpolate', normalize_kernel=True (restated from astropy's
    documentation; astropy is absent -> parity unpinned): NaN samples are skipped and each output is divided by the
    sum of the kernel values it used (the zero padding counts); an output with only NaN samples keeps the input."""
    img = np.asarray(img, dtype=np.float64)
    ker = np.asarray(ker, dtype=np.float64)
    ker = ker / ker.sum()
    ha, hb = ker.shape[0] // 2, ker.shape[1] // 2
    pad = np.zeros((img.shape[0] + 2 * ha, img.shape[1] + 2 * hb))
    pad[ha:ha + img.shape[0], hb:hb + img.shape[1]] = img
    good = ~np.isnan(pad)
    vals = np.where(good, pad, 0.0)
    top, bot = np.zeros_like(img), np.zeros_like(img)
    for a in range(ker.shape[0]):
        for b in range(ker.shape[1]):
            k = ker[ker.shape[0] - 1 - a, ker.shape[1] - 1 - b]
            top += k * vals[a:a + img.shape[0], b:b + img.shape[1]]
            bot += k * good[a:a + img.shape[0], b:b + img.shape[1]]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(bot == 0.0, img, top / bot)


def residual_fwhm(target, current):
    """losvd.py:103-116: quadrature difference, 0 (with a warning) if not larger."""
    if target > current:
        return math.sqrt(target ** 2 - current ** 2), target
    return 0.0, current


def convolve_spatial(cube, stepx, stepy, fwhmx, fwhmy, cur_fwhmx=0.0, cur_fwhmy=0.0):
    """losvd.py:94-141 on a bare cube[nX,nY,nV].  Axis quirk kept: the kernel array is
    [y, x] but slice axis 0 is X, so x_stddev acts along array axis 1 (= Y)."""
    cx, rx = residual_fwhm(fwhmx, cur_fwhmx)
    cy, ry = residual_fwhm(fwhmy, cur_fwhmy)
    ker = gaussian_kernel2d(cx * FWHM_TO_SIGMA / stepx, cy * FWHM_TO_SIGMA / stepy)
    out = np.empty(cube.shape, dtype=np.float64)
    for k in range(cube.shape[2]):
        out[:, :, k] = convolve2d_fill0(cube[:, :, k], ker)
    return out, rx, ry


# --------------------------------------------------------------------------- rebin (f3)
def rebin_spatial(cube, binx, biny, factorx=1, factory=1, function="sum"):
    """losvd.py:69-92 with misc_io.py:370-394: per velocity slice
    a.reshape(n0, fx, n1, fy).sum(-1).sum(1) (or .mean(-1).mean(1)); binx / biny through the same
    function (so 'sum' SUMS the centres).  Raises ValueError like the reference's reshape when the
    shape is not a multiple of the factors."""
    nx, ny, nv = cube.shape
    n0, n1 = nx // factorx, ny // factory
    red = (lambda a, ax: a.mean(ax)) if function == "mean" else (lambda a, ax: a.sum(ax))
    bx = red(np.asarray(binx).reshape((n0, nx // n0)), -1)
    by = red(np.asarray(biny).reshape((n1, ny // n1)), -1)
    out = np.zeros((n0, n1, nv))
    for k in range(nv):
        a = cube[:, :, k].reshape((n0, nx // n0, n1, ny // n1))
        out[:, :, k] = red(red(a, -1), 1)
    return bx, by, out


# --------------------------------------------------------------------------- SSP M/L (f2)
def ml_from_nodes(node_age, node_MH, node_ML, age, MH, fill_nan=True):
    """misc_functions.py:195-202: scipy griddata 'cubic' (Clough-Tocher on the Delaunay triangulation
    of the table nodes), NaN outside the hull replaced by the nearest node.  scipy is the
    reference's own dependency for this step (unpinned; 1.18.1 here), so the oracle calls it the
    way the reference does."""
    from scipy.interpolate import griddata
    age, MH = np.asarray(age, dtype=np.float64), np.asarray(MH, dtype=np.float64)
    ML = griddata((node_age, node_MH), node_ML, (age, MH), method="cubic")
    if fill_nan:
        bad = np.isnan(ML)
        ML[bad] = griddata((node_age, node_MH), node_ML, (age[bad], MH[bad]), method="nearest")
    return ML


# --------------------------------------------------------------------------- LOSVD moments (f1)
def moment012(x, data):
    """misc_functions.py:70-90."""
    if np.all(data == 0):
        return 0.0, 0.0, np.min(np.diff(x)) / 3.0
    m0 = np.sum(data)
    m1 = np.average(x, weights=data)
    m2 = np.sqrt(np.average(x ** 2, weights=data) - m1 ** 2)
    return m0, m1, m2


# --------------------------------------------------------------------------- whole path
def project_path(pos, vel, mass, PA, inclination, limXY, nXY, limV=None, nV=None,
                 f32_angles=True):
    """rotate -> velocity moments (-> LOSVD cube): the timed CPU path of bench.py.
    Uses np.dot / np.histogram2d / np.histogramdd exactly as the reference does
    (rotation.py:115-117, grid.py:126-131, grid.py:259) so that its speed is the
    reference's speed."""
    if f32_angles:
        PA, inclination = as_f32_angle(PA), as_f32_angle(inclination)
    mat = rotation_matrix(PA, inclination)
    p = np.transpose(np.dot(mat, np.transpose(pos)))
    v = np.transpose(np.dot(mat, np.transpose(vel)))
    nx, ny = expand_nxy(nXY)
    ex, ey = bin_edges(limXY[0], limXY[1], nx), bin_edges(limXY[2], limXY[3], ny)
    x, y, vz = p[:, 0], p[:, 1], v[:, 2]
    s0 = np.histogram2d(x, y, [ex, ey], weights=mass)[0].T
    s1 = np.histogram2d(x, y, [ex, ey], weights=mass * vz)[0].T
    s2 = np.histogram2d(x, y, [ex, ey], weights=mass * vz ** 2)[0].T
    V, S = finalise_moments(s0, s1, s2)
    cube = None
    if nV is not None:
        ev = bin_edges(limV[0], limV[1], nV)
        n = x.size
        sample = np.hstack((x.reshape(n, 1), y.reshape(n, 1), vz.reshape(n, 1)))
        cube = np.histogramdd(sample, bins=(ex, ey, ev), weights=mass)[0]
    return s0, V, S, cube


# --------------------------------------------------------------------------- Gauss-Hermite fit
# Row f4.  The *objective* below restates the reference exactly (misc_functions.py:15-68,
# fit_losvd.py:25-54, :178-185, :456-459); the optimiser does not: the reference drives it with a
# vendored MINPACK port (utils/mpfit.py:599-2326, finite-difference Jacobian), which is
# tolerance-defined, so the checkers here are (a) the reference's own fitted parameters, stored as
# golden vectors by oracle/make_golden.py, (b) scipy's bounded trust-region solver on the restated
# objective, and (c) first-order optimality of the restated objective.
def gausshermite(vbin, gh):
    """misc_functions.py:15-68 (note its recursion drops the sqrt(i) factor of the textbook one)."""
    gh = np.asarray(gh, dtype=np.float64)
    degree = len(gh) - 1
    if degree < 2 or gh[2] == 0.0:
        return vbin * 0.0
    w = (vbin - gh[1]) / gh[2]
    w2 = w * w
    h_prev = (2.0 * w2 - 1.0) / np.sqrt(2.0)
    h_cur = (2.0 * w2 - 3.0) * w / np.sqrt(3.0)
    h_use = h_cur
    var = 1.0
    for i in range(3, degree + 1):
        var = var + gh[i] * h_use
        h_use = (np.sqrt(2.0) * h_cur * w - h_prev) / np.sqrt(i + 1.0)
        h_prev = h_cur
        h_cur = h_use
    return gh[0] * var * np.exp(-w2 / 2.0)


def gh_amplitude(data, ufit, err=None):
    """fit_losvd.py:25-54: sum(dn*dn) / sum(dn*fn) -- not the least-squares amplitude."""
    if err is None:
        err = np.ones_like(data)
    err = np.where(err == 0.0, 1.0, err)
    dn = data / err
    fn = ufit / err
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.sum(dn * dn) / np.sum(dn * fn)


def gh_residuals(p, x, data, err=None):
    """fit_losvd.py:178-185 (mpfitfun): amplitude re-solved at every evaluation."""
    gh = np.concatenate(([1.0], p))
    gh[0] = gh_amplitude(data, gausshermite(x, gh), err)
    res = data - gausshermite(x, gh)
    return res if err is None else res / err


def gh_guess_and_limits(x, data, deg_gh):
    """fit_losvd.py:157-165: start at the moments, h_n = 0.02, box limits from the axis."""
    m0, m1, m2 = moment012(x, data)
    p0 = np.full(deg_gh, 0.02)
    p0[:2] = (m1, m2)
    lo = np.full(deg_gh, -0.2)
    lo[:2] = (np.min(x), np.min(np.diff(x)) / 3.0)
    hi = np.full(deg_gh, 0.2)
    hi[:2] = (np.max(x), (np.max(x) - np.min(x)) / 3.0)
    return p0, lo, hi


def fit_gh_scipy(x, data, deg_gh=4, err=None):
    """Independent checker: scipy's bounded trust-region least squares on the restated objective.
    Returns ([I, V, S, h3..], chi2)."""
    from scipy.optimize import least_squares
    p0, lo, hi = gh_guess_and_limits(x, data, deg_gh)
    sol = least_squares(gh_residuals, np.clip(p0, lo, hi), bounds=(lo, hi), args=(x, data, err),
                        x_scale="jac", ftol=1e-14, xtol=1e-14, gtol=1e-14, max_nfev=2000)
    gh = np.concatenate(([1.0], sol.x))
    gh[0] = gh_amplitude(data, gausshermite(x, gh), err)
    return gh, float(np.sum(gh_residuals(sol.x, x, data, err) ** 2))


# --------------------------------------------------------------------------- AMR spreading (f3)
def spread_amr(list_coord, level, list_val=(), list_const=(), nsample=10, box=10, stretch=(), uniforms=None):
    """misc_io.py:184-224.  ``uniforms``: one (nsample * n,) array of U[0,1) deviates per coordinate;
    None draws them from numpy's global generator exactly as the reference does (:209)."""
    scales = box / 2.0 ** np.asarray(level)
    scalesr = np.repeat(scales, nsample)
    n = len(list_coord[0])
    stretch = list(stretch) or [1.0] * len(list_coord)
    coords = []
    for k, (cx, s) in enumerate(zip(list_coord, stretch)):
        u = np.random.random_sample((nsample, n)).ravel() if uniforms is None else uniforms[k]
        coords.append(np.repeat(cx, nsample) + (u - 0.5) * s * scalesr)
    vals = [np.repeat(v / nsample, nsample) for v in list_val]
    consts = [np.repeat(c, nsample) for c in list_const]
    return coords, vals, consts


# --------------------------------------------------------------------------- radial profiles
def point_2vprofiles(x, y, z, Vx, Vy, Vz, nbins=100, dz=(-0.2, 0.2), Rmin=1.e-2, Rmax=None):
    """grid.py:16-68 (with rotation.xy_to_polar, rotation.py:121-144, for its default PA): numpy's own
    float32 / float64 arithmetic, mean() and std() per np.digitize bin."""
    s = np.zeros((3, nbins), dtype=np.float32)
    v = np.zeros_like(s)
    sel = np.where((dz[0] < z) & (z < dz[1]))
    x, y = x[sel], y[sel]
    R = np.sqrt(x ** 2 + y ** 2)
    theta = np.arctan2(y, x)
    VR = Vx[sel] * np.cos(theta) + Vy[sel] * np.sin(theta)
    VT = -Vx[sel] * np.sin(theta) + Vy[sel] * np.cos(theta)
    VZ = Vz[sel]
    if Rmax is None:
        Rmax = R.max()
    Rbin = np.logspace(np.log10(Rmin), np.log10(Rmax), nbins)
    digit = np.digitize(R, Rbin)
    counts = np.zeros(nbins, dtype=np.int64)
    with np.errstate(all="ignore"):
        for i in range(nbins):
            idx = np.where(digit == i)[0]
            counts[i] = idx.size
            for c, comp in enumerate((VR, VT, VZ)):
                s[c, i] = comp[idx].std()
                v[c, i] = comp[idx].mean()
    return Rbin, v, s, counts
