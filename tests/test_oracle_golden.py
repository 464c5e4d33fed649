"""CPU: pins the oracle (numpy restatement + C restatement) against the golden vectors the
reference's own code produced (oracle/make_golden.py).  Bit-exact unless stated."""
import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as on

ANGLES = {"f64ang": (30.0, 60.0), "f32ang": (np.float32(30.0), np.float32(60.0)), "odd": (12.3, 77.7)}


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_rotation_bit_exact(golden, dtname):
    g = golden("rotate_" + dtname)
    pos, vel = g["pos"], g["vel"]
    for tag, (pa, inc) in ANGLES.items():
        for direct, dtag in ((True, "dir"), (False, "inv")):
            mat = on.rotation_matrix(pa, inc, direct)
            # C restatement (explicit fmaf / fma chain)
            rp = np.stack(oc.rotate(mat, pos[:, 0], pos[:, 1], pos[:, 2]), 1)
            rv = np.stack(oc.rotate(mat, vel[:, 0], vel[:, 1], vel[:, 2]), 1)
            assert same(rp, g["pos_%s_%s" % (tag, dtag)]), (tag, dtag)
            assert same(rv, g["vel_%s_%s" % (tag, dtag)]), (tag, dtag)
    # numpy restatement on one case per dtype (exact rational arithmetic is slow for float64)
    n = 256
    p, v = on.rotate_mat(pos[:n], vel[:n], np.float32(30.0), np.float32(60.0))
    assert same(p, g["pos_f32ang_dir"][:n]) and same(v, g["vel_f32ang_dir"][:n])
    mat = on.rotation_matrix_xyz(10.0, 20.0, 30.0, rotorder=[2, 0, 1])
    assert same(np.stack(oc.rotate(mat, pos[:, 0], pos[:, 1], pos[:, 2]), 1), g["pos_xyz"])
    assert same(np.stack(oc.rotate(mat, vel[:, 0], vel[:, 1], vel[:, 2]), 1), g["vel_xyz"])
    p, v = on.rotate_mat(pos[:n], None, 45.0, 45.0)
    assert v is None and same(p, g["pos_novel"][:n])


def test_kat_matrices_and_edges(golden):
    g = golden("kat")
    assert same(on.rotation_matrix(30.0, 60.0), g["mat_f64ang"])
    assert same(on.rotation_matrix(np.float32(30.0), np.float32(60.0)), g["mat_f32ang"])
    assert not np.array_equal(g["mat_f64ang"], g["mat_f32ang"])          # SURVEY 8c KAT 4
    assert np.array_equal(on.bin_edges(-1, 1, 3), [-1.5, -0.5, 0.5, 1.5])
    x = g["x_edge"]
    _, _, M, V, S = on.points_2vmaps(x, np.zeros_like(x), np.ones_like(x), limXY=[-1, 1, -1, 1], nXY=3)
    assert same(M, g["M_edge"]) and same(V, g["V_edge"]) and same(S, g["S_edge"])
    assert M[1].tolist() == [1.0, 2.0, 2.0] and M.sum() == 5.0               # SURVEY 8c KAT 1
    assert oc.bin_index(x, on.bin_edges(-1, 1, 3)).tolist() == [0, 1, 2, 2, -1, -1, -1, -1, -1, 1]
    assert np.array_equal(on.bin_index(x, on.bin_edges(-1, 1, 3)), oc.bin_index(x, on.bin_edges(-1, 1, 3)))
    # mask quirk (SURVEY 8c KAT 2)
    vq = g["v_quirk"]
    _, _, M0, V0, _ = on.points_2vmaps(np.zeros(3), np.zeros(3), vq, limXY=[-1, 1, -1, 1], nXY=3)
    _, _, M1, V1, _ = on.points_2vmaps(np.zeros(3), np.zeros(3), vq, limXY=[-1, 1, -1, 1], nXY=3,
                                       mask=np.array([True, True, False]))
    assert same(M0, g["M_nomask"]) and same(V0, g["V_nomask"]) and same(M1, g["M_mask"]) and same(V1, g["V_mask"])
    assert M0[1, 1] == 3 and np.isnan(V0[1, 1]) and M1[1, 1] == 2 and V1[1, 1] == 2


@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_maps_and_cubes_bit_exact(golden, dtname):
    g = golden("grid_" + dtname)
    x, y, d, mass = g["x"], g["y"], g["d"], g["mass"]
    lim, nxy, limv, nv = g["limXY"].tolist(), tuple(int(v) for v in g["nXY"]), g["limV"].tolist(), int(g["nV"])
    # the oracle's rotation reproduces the projected coordinates the reference binned
    mat = on.rotation_matrix(np.float32(30.0), np.float32(60.0))
    rx, ry, _ = oc.rotate(mat, *g["pos"].T)
    _, _, rd = oc.rotate(mat, *g["vel"].T)
    assert same(rx, x) and same(ry, y) and same(rd, d)
    for wtag, w in (("w", mass), ("u", None)):
        gx, gy, M, V, S = on.points_2vmaps(x, y, d, weights=w, limXY=lim, nXY=nxy)
        assert same(M, g["vm_%s_M" % wtag]) and same(V, g["vm_%s_V" % wtag]) and same(S, g["vm_%s_S" % wtag])
        assert same(gx, g["gridX"]) and same(gy, g["gridY"])
        _, _, W, D, DW = on.points_2map(x, y, d, weights=w, limXY=lim, nXY=nxy)
        assert same(W, g["map_%s_W" % wtag]) and same(D, g["map_%s_D" % wtag]) and same(DW, g["map_%s_DW" % wtag])
        px, py, pv, cube = on.points_2losvds(x, y, d, weights=w, limXY=lim, nXY=nxy, limV=limv, nV=nv)
        assert same(cube, g["cube_%s" % wtag])
        assert same(px, g["binx"]) and same(py, g["biny"]) and same(pv, g["binv"])
        # C restatement, fused (rotation + binning) from the original particles
        sums = oc.SUM_W | oc.SUM_WD | oc.SUM_WD2 | oc.SUM_CUBE
        ex, ey, ev = on.bin_edges(lim[0], lim[1], nxy[0]), on.bin_edges(lim[2], lim[3], nxy[1]), on.bin_edges(limv[0], limv[1], nv)
        soa = tuple(g["pos"].T) + tuple(g["vel"].T)
        maps, ccube = oc.project_bin(soa, w, mat, 1, 1, ex, ey, ev, sums)
        assert same(ccube[0], g["cube_%s" % wtag])
        Mc, Vc, Sc = oc.finalize_vmaps(maps[0], None, sums & 7)
        assert same(Mc, g["vm_%s_M" % wtag]) and same(Vc, g["vm_%s_V" % wtag]) and same(Sc, g["vm_%s_S" % wtag])
        if wtag == "w":
            assert same(maps[0, :, :, 1], g["s1_w"]) and same(maps[0, :, :, 2], g["s2_w"])
        # the W-from-cube split is an exact re-partition of the same terms (tolerance: summation order)
        maps2, cube2 = oc.project_bin(soa, w, mat, 1, 1, ex, ey, ev, sums | oc.W_FROM_CUBE)
        M2, _, _ = oc.finalize_vmaps(maps2[0], cube2[0], sums | oc.W_FROM_CUBE)
        assert np.allclose(M2, Mc, rtol=1e-13, atol=0)
    assert same(on.points_2losvds(x, y, d, limXY=lim, nXY=nxy, limV=limv, nV=nv, mask=g["mask"])[3], g["cube_masked_u"])
    _, _, Mq, Vq, Sq = on.points_2vmaps(x, y, g["d_nan"], weights=mass, limXY=lim, nXY=nxy, mask=g["mask"])
    assert same(Mq, g["vm_q_M"]) and same(Vq, g["vm_q_V"]) and same(Sq, g["vm_q_S"])
    with pytest.raises(ValueError):                                           # SURVEY 8c KAT 3
        on.points_2losvds(x, y, d, weights=mass, limXY=lim, nXY=nxy, limV=limv, nV=nv, mask=g["mask"])
    assert on.points_2vmaps(x, y, d, nXY=(1, 2, 3)) == (0, 0, 0, 0, 0)


@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_smoothing_and_moments(golden, dtname):
    """Smoothing parity is UNPINNED (astropy absent): the golden cubes come from the shim's
    restatement of astropy's documented behaviour; tolerance 1e-12 between restatements."""
    g = golden("grid_" + dtname)
    cube = g["cube_w"]
    stepx, stepy = np.average(np.diff(g["binx"])), np.average(np.diff(g["biny"]))
    for key, (fx, fy, cx, cy, src) in {"smooth_eq": (1.5, 1.5, 0, 0, "cube_w"), "smooth_neq": (1.2, 2.4, 0, 0, "cube_w"),
                                        "smooth_twice": (2.5, 2.0, 1.5, 1.5, "smooth_eq")}.items():
        out, rx, ry = on.convolve_spatial(g[src], stepx, stepy, fx, fy, cx, cy)
        assert np.allclose(out, g[key], rtol=1e-12, atol=1e-12 * np.abs(g[key]).max()), key
        assert (rx, ry) == (fx, fy)
        # C restatement: ker[a][b], a along cube axis 0 = the kernel's y axis (axis quirk)
        k2 = on.gaussian_kernel2d(on.residual_fwhm(fx, cx)[0] * on.FWHM_TO_SIGMA / stepx,
                                  on.residual_fwhm(fy, cy)[0] * on.FWHM_TO_SIGMA / stepy)
        outc = oc.smooth_cube(g[src], k2)
        assert np.allclose(outc, g[key], rtol=1e-12, atol=1e-12 * np.abs(g[key]).max()), key
    m0, m1, m2 = oc.cube_moments(cube, g["binv"])
    assert np.allclose(m0, g["mom0"], rtol=1e-13) and np.allclose(m1, g["mom1"], rtol=1e-10, atol=1e-9)
    assert np.allclose(m2, g["mom2"], rtol=1e-9, atol=1e-6, equal_nan=True)
    r = [on.moment012(g["binv"], cube[i, j]) for i in range(3) for j in range(3)]
    assert np.allclose([v[0] for v in r], g["mom0"][:3, :3].ravel())


@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_rebin_spatial(golden, dtname):
    """f3: losvd.py:69-92 -- sum and mean, coordinates rebinned through the same function."""
    g = golden("grid_" + dtname)
    cube = g["cube_w"]
    for key, (fx, fy, fn) in {"rebin_sum": (4, 2, "sum"), "rebin_mean": (2, 4, "mean")}.items():
        bx, by, out = on.rebin_spatial(cube, g["binx"], g["biny"], fx, fy, fn)
        assert same(out, g[key]) and same(bx, g[key + "_binx"]) and same(by, g[key + "_biny"])
        assert np.allclose(oc.rebin_cube(cube, fx, fy, fn == "mean"), g[key], rtol=1e-14, atol=0)
    with pytest.raises(ValueError):
        on.rebin_spatial(cube, g["binx"], g["biny"], 3, 2)


def test_snapshot_level(golden):
    """pynmap.py:246-455 through the oracle: float32 angles, rotate, maps, cube, compounded rotation."""
    g = golden("snapshot")
    pos, vel, mass = g["pos"], g["vel"], g["mass"]
    m1 = on.rotation_matrix(on.as_f32_angle(30.), on.as_f32_angle(60.))
    p = np.stack(oc.rotate(m1, *pos.T), 1)
    v = np.stack(oc.rotate(m1, *vel.T), 1)
    assert same(p, g["pos_rot"]) and same(v, g["vel_rot"])
    lim = [-12, 12, -12, 12]
    _, _, M, V, S = on.points_2vmaps(p[:, 0], p[:, 1], v[:, 2], weights=mass, limXY=lim, nXY=32)
    assert same(M, g["M"]) and same(V, g["V"]) and same(S, g["S"])
    _, _, W, D, DW = on.points_2map(p[:, 0], p[:, 1], mass, weights=mass, limXY=lim, nXY=32)
    assert same(W, g["W"]) and same(D, g["D"]) and same(DW, g["DW"])
    _, _, pv, cube = on.points_2losvds(p[:, 0], p[:, 1], v[:, 2], weights=mass, limXY=lim, nXY=16, limV=[-350, 350], nV=20)
    assert same(cube, g["cube"]) and same(pv, g["binv"])
    m2 = on.rotation_matrix(on.as_f32_angle(15.), on.as_f32_angle(20.))
    assert same(np.stack(oc.rotate(m2, *p.T), 1), g["pos_rot2"])
    # fused chain (two matrices applied per particle) == compounded materialised rotations
    ex = on.bin_edges(-12, 12, 32)
    maps, _ = oc.project_bin(tuple(pos.T) + tuple(vel.T), mass, np.stack([m1, m2]), 2, 1, ex, ex, None, 7)
    M2, V2, S2 = oc.finalize_vmaps(maps[0], None, 7)
    assert same(M2, g["M2"]) and same(V2, g["V2"]) and same(S2, g["S2"])


def test_fixed_point_oracle():
    """O3: the int64 fixed-point image is order independent and within quantisation of O1."""
    rng = np.random.default_rng(3)
    n = 20000
    soa = [rng.normal(0, 4, n).astype(np.float32) for _ in range(3)] + [rng.normal(0, 120, n).astype(np.float32) for _ in range(3)]
    w = rng.uniform(0.5, 2, n).astype(np.float32)
    mat = on.rotation_matrix(np.float32(30), np.float32(60))
    ex, ev = on.bin_edges(-10, 10, 24), on.bin_edges(-300, 300, 16)
    scales = [2.0 ** 30, 2.0 ** 22, 2.0 ** 14]
    a = oc.project_bin(soa, w, mat, 1, 1, ex, ex, ev, 15, oc.ACC_FIXED, scales)
    perm = rng.permutation(n)
    b = oc.project_bin([s[perm] for s in soa], w[perm], mat, 1, 1, ex, ex, ev, 15, oc.ACC_FIXED, scales)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    f = oc.project_bin(soa, w, mat, 1, 1, ex, ex, ev, 15)
    assert np.allclose(a[1] / scales[0], f[1], rtol=0, atol=n * 2.0 ** -31)
    assert np.allclose(a[0][..., 2] / scales[2], f[0][..., 2], rtol=1e-6, atol=n * 2.0 ** -15)
    # numpy statement of the same quantisation
    rx, ry, _ = oc.rotate(mat, *soa[:3])
    ix, iy = on.bin_index(rx, ex), on.bin_index(ry, ex)
    keep = (ix >= 0) & (iy >= 0)
    q = on.fixed_hist(iy.astype(np.int64) * 24 + ix, keep, w, scales[0], 24 * 24)
    assert np.array_equal(q.reshape(24, 24), a[0][0, :, :, 0])


def test_convolve_nan_interpolate_known_answers():
    """astropy.convolution.convolve defaults (boundary='fill', fill_value=0, nan_treatment='interpolate',
    normalize_kernel=True), restated in oracle_np.convolve2d_fill0 and pnmo_smooth_cube.  astropy is absent, so these
    known answers follow its documented algorithm (and its docstring example for the NaN-free case) -- the smoothing
    row stays 'parity unpinned'."""
    # NaN-free, zero fill (astropy docs: convolve([1, 4, 5, 6, 5, 7, 8], [0.2, 0.6, 0.2]) -> [1.4, 3.6, 5, 5.6, 5.6, 6.8, 6.2])
    row = np.array([[1., 4., 5., 6., 5., 7., 8.]])
    k = np.array([[0.2, 0.6, 0.2]])
    want = np.array([[1.4, 3.6, 5.0, 5.6, 5.6, 6.8, 6.2]])
    assert np.allclose(on.convolve2d_fill0(row, k), want, rtol=1e-14)
    assert np.allclose(oc.smooth_cube(row[:, :, None], k)[:, :, 0], want, rtol=1e-14)
    # a NaN sample is left out and the kernel renormalised over the samples used; the zero padding counts
    row = np.array([[1., np.nan, 3.]])
    k = np.array([[1., 1., 1.]]) / 3.0
    want = np.array([[(0. + 1.) / 2., (1. + 3.) / 2., (3. + 0.) / 2.]])
    assert np.allclose(on.convolve2d_fill0(row, k), want, rtol=1e-14)
    assert np.allclose(oc.smooth_cube(row[:, :, None], k)[:, :, 0], want, rtol=1e-14)
    # 2-D: centre NaN, 3 x 3 box -> mean of the 8 neighbours at the centre; corners use 3 real samples + 5 padded zeros
    img = np.arange(1., 10.).reshape(3, 3)
    img[1, 1] = np.nan
    box = np.full((3, 3), 1. / 9.)
    got = on.convolve2d_fill0(img, box)
    assert np.isclose(got[1, 1], (1 + 2 + 3 + 4 + 6 + 7 + 8 + 9) / 8.)
    assert np.isclose(got[0, 0], (1 + 2 + 4) / 8.)                # 3 samples + 5 zeros; the NaN is not counted
    assert np.allclose(oc.smooth_cube(img[:, :, None], box)[:, :, 0], got, rtol=1e-14)
    # every sample NaN -> the output keeps the input value (NaN): a 1 x 1 kernel on a NaN pixel
    one = np.array([[np.nan, 2.]])
    assert np.isnan(on.convolve2d_fill0(one, np.array([[1.]]))[0, 0]) and on.convolve2d_fill0(one, np.array([[1.]]))[0, 1] == 2.
