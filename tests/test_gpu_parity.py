"""GPU (-m gpu): parity of the CUDA path, called through the reference-shaped Python API and the
C ABI underneath it, against (1) the golden vectors the reference produced and (2) the oracle on
seeded inputs.  Bars: bit-exact for rotation, bin indices, counts and everything in fixed-point
mode; <= 1e-12 relative (contract: 1e-6) for float64-atomic sums, whose only freedom is the
order of the additions.  /root/reference is never touched here."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import oracle_c as oc
from oracle import oracle_np as on

pytestmark = pytest.mark.gpu

TOL = 1e-12          # observed ~1e-15; the contract (BASELINE.json north_star) is 1e-6
ANGLES = {"f64ang": (30.0, 60.0), "f32ang": (np.float32(30.0), np.float32(60.0)), "odd": (12.3, 77.7)}


@pytest.fixture(scope="module")
def pnm():
    import pynmap_b200
    pynmap_b200._lib.require_device()
    return pynmap_b200


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def sums_close(a, b, absref=None, tol=TOL):
    """|a-b| <= tol * (sum of |terms|): relative error for positive sums, backward-stable measure
    for the signed sum w*v (absref = the same sum over |terms|)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert (np.isnan(a) == np.isnan(b)).all()
    scale = np.abs(b) if absref is None else np.maximum(np.abs(absref), np.abs(b))
    ok = ~np.isnan(b)
    return bool(np.all(np.abs(a[ok] - b[ok]) <= tol * scale[ok] + 1e-300))


# ------------------------------------------------------------------------------- a1 rotation
@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_rotate_mat_bit_exact_vs_reference(pnm, golden, dtname):
    g = golden("rotate_" + dtname)
    pos, vel = g["pos"], g["vel"]
    for tag, (pa, inc) in ANGLES.items():
        for direct, dtag in ((True, "dir"), (False, "inv")):
            p, v = pnm.rotation.rotate_mat(pos, vel, pa, inc, direct)
            assert same(p, g["pos_%s_%s" % (tag, dtag)]) and same(v, g["vel_%s_%s" % (tag, dtag)]), (tag, dtag)
            assert p.flags.f_contiguous            # column slices contiguous like the reference's result
    p, v = pnm.rotation.rotateXYZ_mat(pos, vel, 10.0, 20.0, 30.0, rotorder=[2, 0, 1])
    assert same(p, g["pos_xyz"]) and same(v, g["vel_xyz"])
    p, v = pnm.rotation.rotate_mat(pos, None, 45.0, 45.0)
    assert v is None and same(p, g["pos_novel"])
    # CUDA tensors in -> CUDA tensors out, same bits; F-ordered input is taken without copy
    tp = torch.from_numpy(np.asfortranarray(pos)).cuda()
    p2, _ = pnm.rotation.rotate_mat(tp, None, 45.0, 45.0)
    assert p2.is_cuda and same(p2.cpu().numpy(), g["pos_novel"])
    with pytest.raises(ValueError):
        pnm.rotation.rotate_mat(pos[:, :2], None, 1.0, 2.0)


def test_rotate_large_random_vs_oracle(pnm):
    rng = np.random.default_rng(7)
    for dt in (np.float32, np.float64):
        pos = (rng.standard_normal((300001, 3)) * 7).astype(dt)
        mat = on.rotation_matrix(np.float32(33.3), np.float32(71.9))
        p, _ = pnm.rotation.rotate_mat(pos, None, np.float32(33.3), np.float32(71.9))
        ref = np.stack(oc.rotate(mat, pos[:, 0], pos[:, 1], pos[:, 2]), 1)
        assert same(p, ref)


# ------------------------------------------------------------------------------- a2-a4 grid
def test_kat_edges_nan_inf_mask_quirk(pnm, golden):
    g = golden("kat")
    x = g["x_edge"]
    for dt in (np.float64, np.float32):
        xs = x.astype(dt)
        gx, gy, M, V, S = pnm.grid.points_2vmaps(xs, np.zeros_like(xs), np.ones_like(xs), limXY=[-1, 1, -1, 1], nXY=3)
        if dt == np.float64:
            assert same(M, g["M_edge"]) and same(V, g["V_edge"]) and same(S, g["S_edge"])
            assert same(gx, g["gridX_edge"]) and same(gy, g["gridY_edge"])
        else:   # float32 cannot hold 1.5000001: compare with the oracle on the rounded inputs
            _, _, Mo, _, _ = on.points_2vmaps(xs, np.zeros_like(xs), np.ones_like(xs), limXY=[-1, 1, -1, 1], nXY=3)
            assert same(M, Mo)
    vq = g["v_quirk"]
    _, _, M0, V0, _ = pnm.grid.points_2vmaps(np.zeros(3), np.zeros(3), vq, limXY=[-1, 1, -1, 1], nXY=3)
    _, _, M1, V1, _ = pnm.grid.points_2vmaps(np.zeros(3), np.zeros(3), vq, limXY=[-1, 1, -1, 1], nXY=3,
                                             mask=np.array([True, True, False]))
    assert same(M0, g["M_nomask"]) and same(V0, g["V_nomask"]) and same(M1, g["M_mask"]) and same(V1, g["V_mask"])
    assert pnm.grid.points_2vmaps(vq, vq, vq, nXY=(1, 2, 3)) == (0, 0, 0, 0, 0)
    assert pnm.grid.points_2map(vq, vq, vq, nXY=(1, 2, 3)) == (0, 0, 0, 0, 0)
    assert pnm.grid.points_2losvds(vq, vq, vq, nXY=(1, 2, 3)) == (0, 0, 0, 0, 0)


@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_points_api_vs_reference_golden(pnm, golden, dtname):
    g = golden("grid_" + dtname)
    x, y, d, mass = g["x"], g["y"], g["d"], g["mass"]
    lim, nxy, limv, nv = g["limXY"].tolist(), tuple(int(v) for v in g["nXY"]), g["limV"].tolist(), int(g["nV"])
    # unweighted: every sum of ones is an exact integer -> bit-exact maps and cube (bin indices pinned)
    gx, gy, M, V, S = pnm.grid.points_2vmaps(x, y, d, limXY=lim, nXY=nxy)
    assert same(M, g["vm_u_M"]) and same(gx, g["gridX"]) and same(gy, g["gridY"])
    assert rel_err(V, g["vm_u_V"], floor=1e-3) < 1e-9
    L = pnm.grid.points_2losvds(x, y, d, limXY=lim, nXY=nxy, limV=limv, nV=nv)
    assert same(L.losvd, g["cube_u"]) and same(L.binx, g["binx"]) and same(L.biny, g["biny"]) and same(L.binv, g["binv"])
    assert L.losvd.shape == (nxy[0], nxy[1], nv) and M.shape == (nxy[1], nxy[0])   # SURVEY 8c KAT 5
    Lm = pnm.grid.points_2losvds(x, y, d, limXY=lim, nXY=nxy, limV=limv, nV=nv, mask=g["mask"])
    assert same(Lm.losvd, g["cube_masked_u"])
    # weighted: float64 atomics, order of additions is the only freedom
    _, _, M, V, S = pnm.grid.points_2vmaps(x, y, d, weights=mass, limXY=lim, nXY=nxy)
    assert sums_close(M, g["vm_w_M"])
    assert rel_err(V, g["vm_w_V"], floor=1e-3) < 1e-9
    big = g["vm_w_S"] > 1.0          # sigma is ill-conditioned in cold single-particle pixels
    assert rel_err(np.asarray(S)[big], g["vm_w_S"][big]) < 1e-8
    _, _, W, D, DW = pnm.grid.points_2map(x, y, d, weights=mass, limXY=lim, nXY=nxy)
    absd = on.hist2d(x, y, g["edges_x"], g["edges_y"], np.abs(mass * d)).T
    assert sums_close(W, g["map_w_W"]) and sums_close(DW, g["map_w_DW"], absd) and same(DW == 0, g["map_w_DW"] == 0)
    assert rel_err(D, g["map_w_D"], floor=1e-3) < 1e-9
    _, _, W, D, DW = pnm.grid.points_2map(x, y, d, limXY=lim, nXY=nxy)
    assert same(W, g["map_u_W"])
    L = pnm.grid.points_2losvds(x, y, d, weights=mass, limXY=lim, nXY=nxy, limV=limv, nV=nv)
    assert sums_close(L.losvd, g["cube_w"]) and same(L.losvd == 0, g["cube_w"] == 0)
    # mask quirk with NaN data (only masked AND non-finite particles are removed)
    _, _, Mq, Vq, Sq = pnm.grid.points_2vmaps(x, y, g["d_nan"], weights=mass, limXY=lim, nXY=nxy, mask=g["mask"])
    assert sums_close(Mq, g["vm_q_M"]) and same(np.isnan(Vq), np.isnan(g["vm_q_V"]))
    with pytest.raises(ValueError):
        pnm.grid.points_2losvds(x, y, d, weights=mass, limXY=lim, nXY=nxy, limV=limv, nV=nv, mask=g["mask"])
    # CUDA tensors in -> CUDA tensors out
    tx, ty, td = (torch.from_numpy(a).cuda() for a in (x, y, d))
    _, _, Mt, _, _ = pnm.grid.points_2vmaps(tx, ty, td, limXY=lim, nXY=nxy)
    assert Mt.is_cuda and same(Mt.cpu().numpy(), g["vm_u_M"])


@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_fixed_point_bit_exact_vs_oracle(pnm, golden, dtname):
    """int64 fixed-point accumulation: identical bits to the CPU restatement of the same
    quantisation (O3), and within quantisation of the float64 reference."""
    from pynmap_b200 import _lib, engine
    g = golden("grid_" + dtname)
    soa_np = [np.ascontiguousarray(c) for c in g["pos"].T] + [np.ascontiguousarray(c) for c in g["vel"].T]
    mass = g["mass"]
    lim, nxy, limv, nv = g["limXY"].tolist(), tuple(int(v) for v in g["nXY"]), g["limV"].tolist(), int(g["nV"])
    mat = on.rotation_matrix(np.float32(30.0), np.float32(60.0))
    grid = engine.get_grid(lim, nxy[0], nxy[1], limv, nv)
    acc = engine.AccMode("fixed")
    scales = engine.fixed_scales(mass.size, float(np.abs(mass).max()), 2.0 * float(np.abs(g["vel"]).max()))
    for sums in (15, 15 | 16, 7, 8):
        dt = torch.float32 if dtname == "float32" else torch.float64
        soa = [torch.from_numpy(a).cuda().to(dt) for a in soa_np]
        w = torch.from_numpy(mass).cuda().to(dt)
        maps, cube = engine.new_accumulators(grid, sums, acc)
        engine.project_bin(grid, soa, w, mat, 1, 1, sums, acc, scales, maps, cube)
        omaps, ocube = oc.project_bin(soa_np, mass, mat, 1, 1, grid.edges_x, grid.edges_y, grid.edges_v, sums,
                                      oc.ACC_FIXED, scales)
        if cube is not None:
            got = engine.finalize_cube(grid, cube, acc, scales).cpu().numpy()
            assert same(got, ocube[0].astype(np.float64) * (1.0 / scales[0]))
        if sums & 7 == 7:
            M, V, S, S1, S2 = (t.cpu().numpy() for t in engine.finalize_vmaps(grid, maps, cube, sums, acc, scales, want_raw=True))
            Mo, Vo, So = oc.finalize_vmaps(omaps[0], ocube[0] if sums & 8 else None, sums, oc.ACC_FIXED, scales)
            assert same(M, Mo) and same(V, Vo) and same(S, So)          # finalisation is bit-exact too
            assert same(S1, omaps[0, :, :, 1].astype(np.float64) * (1.0 / scales[1]))
            assert same(S2, omaps[0, :, :, 2].astype(np.float64) * (1.0 / scales[2]))
            assert rel_err(M, g["vm_w_M"]) < 1e-6
            folded = engine.fold_maps(grid, maps, acc).cpu().numpy().reshape(omaps.shape[1:])
            # raw int64 sums, replicas folded; device slot order is (S1, S2, S0, -), the oracle's (S0, S1, S2, -)
            assert np.array_equal(folded[..., [2, 0, 1]], omaps[0][..., :3])
    # the user-facing keyword
    _, _, Mf, Vf, Sf = pnm.grid.points_2vmaps(g["x"], g["y"], g["d"], weights=mass, limXY=lim, nXY=nxy, accumulate="fixed")
    assert rel_err(Mf, g["vm_w_M"]) < 1e-6
    _, _, Mf2, Vf2, Sf2 = pnm.grid.points_2vmaps(g["x"][::-1].copy(), g["y"][::-1].copy(), g["d"][::-1].copy(),
                                                 weights=mass[::-1].copy(), limXY=lim, nXY=nxy, accumulate="fixed")
    assert same(Mf, Mf2) and same(Vf, Vf2) and same(Sf, Sf2)            # order independent


def test_bin_indices_on_adversarial_points(pnm):
    """Points exactly on, one ulp below and one ulp above every edge, in float32 and float64:
    integer counts must equal the oracle's searchsorted everywhere."""
    for n_bins, lo, hi in ((256, -15.0, 15.0), (100, -1000.0, 1000.0), (7, -0.3, 0.9)):
        e = on.bin_edges(lo, hi, n_bins)
        for dt in (np.float32, np.float64):
            ed = e.astype(dt)
            pts = np.concatenate([ed, np.nextafter(ed, dt(-np.inf)), np.nextafter(ed, dt(np.inf)),
                                  np.array([np.nan, np.inf, -np.inf, 0.0], dtype=dt)])
            other = np.zeros_like(pts)
            for xs, ys in ((pts, other), (other, pts)):
                _, _, M, _, _ = pnm.grid.points_2vmaps(xs, ys, np.ones_like(pts), limXY=[lo, hi, lo, hi], nXY=n_bins)
                _, _, Mo, _, _ = on.points_2vmaps(xs, ys, np.ones_like(pts), limXY=[lo, hi, lo, hi], nXY=n_bins)
                assert same(M, Mo), (n_bins, dt)
            L = pnm.grid.points_2losvds(other, other, pts, limXY=[-1, 1, -1, 1], nXY=3, limV=[lo, hi], nV=n_bins)
            assert same(L.losvd, on.points_2losvds(other, other, pts, limXY=[-1, 1, -1, 1], nXY=3, limV=[lo, hi], nV=n_bins)[3])


def test_ragged_empty_and_unaligned(pnm):
    from pynmap_b200 import engine
    rng = np.random.default_rng(5)
    for n in (0, 1, 3, 5, 1023, 4097):
        x = rng.normal(0, 3, n).astype(np.float32)
        y = rng.normal(0, 3, n).astype(np.float32)
        d = rng.normal(0, 100, n).astype(np.float32)
        _, _, M, _, _ = pnm.grid.points_2vmaps(x, y, d, limXY=[-5, 5, -5, 5], nXY=16)
        assert same(M, on.points_2vmaps(x, y, d, limXY=[-5, 5, -5, 5], nXY=16)[2])
    # views that start 4 bytes into an allocation: the scalar-load kernel variant
    n = 10001
    buf = torch.from_numpy(rng.normal(0, 3, (3, n + 1)).astype(np.float32)).cuda()
    x, y, d = buf[0, 1:], buf[1, 1:], buf[2, 1:]
    assert x.data_ptr() % 16 != 0
    grid = engine.get_grid([-5, 5, -5, 5], 16, 16)
    acc = engine.AccMode("f64")
    maps, _ = engine.new_accumulators(grid, 7, acc)
    engine.bin_points(grid, x, y, d, None, 7, acc, None, maps, None)
    M, _, _ = engine.finalize_vmaps(grid, maps, None, 7, acc, None)
    xs, ys, ds = (t.cpu().numpy() for t in (x, y, d))
    assert same(M.cpu().numpy(), on.points_2vmaps(xs, ys, ds, limXY=[-5, 5, -5, 5], nXY=16)[2])


# ------------------------------------------------------------------------------- snapshot (fused)
def test_snapshot_fused_vs_reference_golden(pnm, golden):
    g = golden("snapshot")
    snap = pnm.snapshot.from_arrays(g["pos"], g["vel"], mass=g["mass"])
    snap.rotate(PA=30., inclination=60.)
    assert same(snap.pos, g["pos_rot"]) and same(snap.vel, g["vel_rot"])
    lim = [-12, 12, -12, 12]
    vm = snap.velocity_moments(weight_name="mass", limXY=lim, nXY=32)
    assert sums_close(vm.M, g["M"]) and rel_err(vm.V, g["V"], floor=1e-3) < 1e-9 and same(vm.X, g["X"]) and same(vm.Y, g["Y"])
    big = g["S"] > 1.0
    assert rel_err(vm.S[big], g["S"][big]) < 1e-8
    assert [float(p) for p in vm.PAs] == [30.0] and [float(i) for i in vm.inclinations] == [60.0]
    mp = snap.create_map(weight_name="mass", limXY=lim, nXY=32)
    assert sums_close(mp.W, g["W"]) and sums_close(mp.DW, g["DW"]) and rel_err(mp.D, g["D"], floor=1e-3) < 1e-9
    lv = snap.create_losvds(weight_name="mass", limXY=lim, nXY=16, limV=[-350, 350], nV=20)
    assert sums_close(lv.losvd, g["cube"]) and same(lv.binv, g["binv"]) and same(lv.losvd == 0, g["cube"] == 0)
    # fixed-point fused path agrees with the float64 reference to quantisation
    vmf = snap.velocity_moments(weight_name="mass", limXY=lim, nXY=32, accumulate="fixed")
    assert rel_err(vmf.M, g["M"]) < 1e-6
    # one read for maps + cube (W recovered from the cube) == separate calls
    m2, l2 = snap.project(weight_name="mass", limXY=lim, nXY=32, limV=[-350, 350], nV=20)
    assert sums_close(m2.M, g["M"]) and rel_err(m2.V, g["V"], floor=1e-3) < 1e-9
    # compounded rotation on rounded coordinates (reset=False), applied as a register chain
    snap.rotate(PA=15., inclination=20., reset=False)
    assert same(snap.pos, g["pos_rot2"])
    vm2 = snap.velocity_moments(weight_name="mass", limXY=lim, nXY=32)
    assert sums_close(vm2.M, g["M2"]) and rel_err(vm2.V, g["V2"], floor=1e-3) < 1e-9
    assert [float(p) for p in vm2.PAs] == g["PAs2"].tolist()
    snap.reset()
    assert same(snap.pos, g["pos"]) and snap.PAs == []
    with pytest.raises(ValueError):
        snap.rotate(PA=[30, 40], inclination=[60, 70])          # SURVEY 8c KAT 6


def test_project_views_equals_loop(pnm):
    import workloads
    soa = workloads.make_snapshot(200000, seed=9, device="cuda")
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="numpy")
    pas, incs = [0.0, 45.0, 90.0, 157.5], [0.0, 30.0, 60.0, 87.5]
    views = snap.project_views(pas, incs, weight_name="mass", limXY=workloads.LIM_XY, nXY=64)
    for k, (pa, inc) in enumerate(zip(pas, incs)):
        snap.rotate(pa, inc)
        one = snap.velocity_moments(weight_name="mass", limXY=workloads.LIM_XY, nXY=64)
        assert same(views[k].M, one.M)                           # unit masses: exact counts
        assert rel_err(views[k].V, one.V, floor=1e-3) < 1e-9


# ------------------------------------------------------------------------------- a5 smoothing, f1
@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_smoothing_and_vmoments(pnm, golden, dtname):
    g = golden("grid_" + dtname)
    L = pnm.LOSVD(g["binx"], g["biny"], g["binv"], g["cube_w"])
    c1 = L.convolve_spatial(1.5, 1.5)
    c2 = L.convolve_spatial(1.2, 2.4)
    c3 = c1.convolve_spatial(2.5, 2.0)
    for out, key in ((c1, "smooth_eq"), (c2, "smooth_neq"), (c3, "smooth_twice")):
        assert np.allclose(out.losvd, g[key], rtol=1e-10, atol=1e-12 * np.abs(g[key]).max()), key
    assert (c3.fwhmx, c3.fwhmy) == (2.5, 2.0) and L.fwhmx == 0.0 and same(L.losvd, g["cube_w"])
    assert L.convolve_spatial() is None and L.convolve_spatial(-1., 0.) is None
    # against the C oracle's direct 2-D convolution on a CUDA-tensor cube
    Lt = pnm.LOSVD(g["binx"], g["biny"], g["binv"], torch.from_numpy(g["cube_u"]).cuda())
    ct = Lt.convolve_spatial(2.0, 1.0)
    k2 = on.gaussian_kernel2d(2.0 * on.FWHM_TO_SIGMA / L.stepx, 1.0 * on.FWHM_TO_SIGMA / L.stepy)
    ref = oc.smooth_cube(g["cube_u"], k2)
    assert ct.losvd.is_cuda and np.allclose(ct.losvd.cpu().numpy(), ref, rtol=1e-10, atol=1e-12 * ref.max())
    L.vmoments()
    assert np.allclose(L.mom0, g["mom0"], rtol=1e-12) and np.allclose(L.mom1, g["mom1"], rtol=1e-9, atol=1e-8)
    # mom2 = sqrt(<v^2> - <v>^2) is ill-conditioned where one velocity channel holds all the light
    # (0, 1e-6 or NaN depending on rounding): compare the variance with an absolute floor
    var, var_ref = np.nan_to_num(L.mom2) ** 2, np.nan_to_num(g["mom2"]) ** 2
    assert np.allclose(var, var_ref, rtol=1e-9, atol=1e-9 * float(np.max(g["binv"] ** 2)))
    wide = g["mom2"] > 1.0
    assert np.allclose(L.mom2[wide], g["mom2"][wide], rtol=1e-9)


@pytest.mark.parametrize("dtname", ["float32", "float64"])
def test_rebin_spatial(pnm, golden, dtname):
    """f3: LOSVD.rebin_spatial (losvd.py:69-92) against the reference's output."""
    g = golden("grid_" + dtname)
    L = pnm.LOSVD(g["binx"], g["biny"], g["binv"], g["cube_w"], fwhmx=0.3, fwhmy=0.4)
    rs, rm = L.rebin_spatial(4, 2), L.rebin_spatial(2, 4, "mean")
    assert rs.losvd.shape == (10, 16, 20) and (rs.fwhmx, rs.fwhmy) == (0.3, 0.4)
    assert np.allclose(rs.losvd, g["rebin_sum"], rtol=1e-14, atol=0) and same(rs.binx, g["rebin_sum_binx"])
    assert np.allclose(rm.losvd, g["rebin_mean"], rtol=1e-14, atol=0) and same(rm.biny, g["rebin_mean_biny"])
    assert same(rs.biny, g["rebin_sum_biny"]) and same(rm.binx, g["rebin_mean_binx"])
    with pytest.raises(ValueError):
        L.rebin_spatial(3, 2)
    Lt = pnm.LOSVD(g["binx"], g["biny"], g["binv"], torch.from_numpy(g["cube_w"]).cuda())
    assert Lt.rebin_spatial(4, 2).losvd.is_cuda


# ------------------------------------------------------------------------------- host-buffer ABI
def test_project_host_matches_device_path(pnm):
    import ctypes
    import workloads
    from pynmap_b200 import _lib, engine
    n = 9_000_077                   # three staging chunks (4 Mi particles each), ragged tail
    soa = workloads.make_snapshot(n, seed=4, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    mat = on.rotation_matrix(np.float32(30.0), np.float32(60.0))
    ex, ey, ev = on.bin_edges(-15, 15, 64), on.bin_edges(-15, 15, 48), on.bin_edges(-500, 500, 32)
    M, V, S = (np.empty((48, 64)) for _ in range(3))
    cube = np.empty((64, 48, 32))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    scales = engine.fixed_scales(n, 2.0, 2.0 * float(np.abs(host[3:6]).max()))
    for mode, sc in ((_lib.ACC_F64, None), (_lib.ACC_FIXED, _lib.doubles(scales))):
        _lib.call("pnm_project_host", 0, n, _lib.F32, *[P(host[k]) for k in range(7)],
                  mat.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 1,
                  64, _lib.doubles(ex), 48, _lib.doubles(ey), 32, _lib.doubles(ev), mode, sc, None, 0, None, 0,
                  P(M), P(V), P(S), P(cube))
        sums = 15 | 16
        omaps, ocube = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, ex, ey, ev, sums, mode, scales if sc else None)
        Mo, Vo, So = oc.finalize_vmaps(omaps[0], ocube[0], sums, mode, scales if sc else None)
        if mode == _lib.ACC_FIXED:
            assert same(M, Mo) and same(V, Vo) and same(S, So)
            assert same(cube, ocube[0].astype(np.float64) * (1.0 / scales[0]))
        else:
            assert sums_close(M, Mo) and sums_close(cube, ocube[0]) and rel_err(V, Vo, floor=1e-3) < 1e-9
    _lib.call("pnm_host_release")


# ------------------------------------------------------------------------------- full size
def test_full_size_properties(pnm):
    """BASELINE config 2 size (10^8 particles, 128x128x256): properties that need no oracle pass --
    conservation (cube + remainder == map), linearity over particle halves, determinism and
    float64-vs-fixed agreement; plus an oracle spot check on a 2x10^6 slice."""
    import workloads
    from pynmap_b200 import _lib, engine
    n = 100_000_000
    soa = workloads.make_snapshot(n, seed=1235, device="cuda")
    mat = on.rotation_matrix(np.float32(30.0), np.float32(60.0))
    grid = engine.get_grid(workloads.LIM_XY, 128, 128, workloads.LIM_V, 256)
    f64, fx = engine.AccMode("f64"), engine.AccMode("fixed")
    scales = engine.fixed_scales(n, 1.0, 2.0 * max(engine.absmax(soa[k]) for k in (3, 4, 5)))
    rows = [soa[k] for k in range(6)]

    def run(acc, sums, lo=0, hi=n, maps=None, cube=None):
        if maps is None:
            maps, cube = engine.new_accumulators(grid, sums, acc)
        engine.project_bin(grid, [r[lo:hi] for r in rows], soa[6][lo:hi], mat, 1, 1, sums, acc,
                           scales if acc is fx else None, maps, cube)
        return maps, cube

    full = 15
    maps_a, cube_a = run(fx, full)
    maps_b, cube_b = run(fx, full)
    assert torch.equal(maps_a, maps_b) and torch.equal(cube_a, cube_b)              # deterministic
    fold_a = engine.fold_maps(grid, maps_a.clone(), fx)
    half = (n // 2) + 3                                                             # ragged split
    maps_c, cube_c = run(fx, full, 0, half)
    run(fx, full, half, n, maps_c, cube_c)
    assert torch.equal(fold_a, engine.fold_maps(grid, maps_c, fx)) and torch.equal(cube_a, cube_c)   # linear in the particle set
    # fold without clearing + PNM_MAPS_FOLDED == finalising all replicas; fold with clearing keeps the buffer summable
    maps_e = maps_a.clone()
    fold_e = engine.fold_maps(grid, maps_e, fx, clear=False)
    assert torch.equal(fold_e, fold_a) and torch.equal(maps_e[grid.maps_fold_elems:], maps_a[grid.maps_fold_elems:])
    Me = engine.finalize_vmaps(grid, maps_e, cube_a, full | _lib.MAPS_FOLDED, fx, scales)
    Mref = engine.finalize_vmaps(grid, maps_a, cube_a, full, fx, scales)
    assert all(torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)) for a, b in zip(Me, Mref))
    maps_f = maps_a.clone()
    engine.fold_maps(grid, maps_f, fx)
    assert grid.map_replicas > 1 and not bool(maps_f[grid.maps_fold_elems:].any())
    Mf = engine.finalize_vmaps(grid, maps_f, cube_a, full, fx, scales)
    assert all(torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)) for a, b in zip(Mf, Mref))
    maps_d, cube_d = run(fx, full | 16)                                             # W recovered from the cube
    Ma = engine.finalize_vmaps(grid, maps_a, cube_a, full, fx, scales)
    Md = engine.finalize_vmaps(grid, maps_d, cube_d, full | 16, fx, scales)
    assert all(torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(d, nan=-1.0)) for a, d in zip(Ma, Md))
    assert torch.equal(cube_a, cube_d)
    # unit masses: the map of sum w is an exact particle count; conservation against the inputs
    M = Ma[0]
    assert float(M.sum()) == float(torch.round(M).sum()) and 0.9 * n < float(M.sum()) <= n
    assert float(cube_a.sum()) / scales[0] <= float(M.sum())
    # float64 atomics vs fixed point
    maps_f, cube_f = run(f64, full)
    Mf = engine.finalize_vmaps(grid, maps_f, cube_f, full, f64, None, want_raw=True)
    Mx = engine.finalize_vmaps(grid, maps_a, cube_a, full, fx, scales, want_raw=True)
    assert torch.equal(Mf[0], M)                                                     # counts: exact in both
    assert torch.equal(cube_f, cube_a.double() / scales[0])
    s1f, s2f, s1x, s2x = Mf[3], Mf[4], Mx[3], Mx[4]
    # fixed point rounds every term to 1/scale: |error| <= count * 0.5 / scale (+ float64 summation noise)
    assert bool(((s2f - s2x).abs() <= M * (0.5 / scales[2]) + 1e-12 * s2f.abs()).all())
    assert bool(((s1f - s1x).abs() <= M * (0.5 / scales[1]) + 1e-9 * s2f.abs().sqrt() * M.sqrt()).all())
    # oracle spot check on a slice
    m = 2_000_000
    host = soa[:, :m].cpu().numpy()
    maps_s, cube_s = run(fx, full, 0, m)
    omaps, ocube = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, grid.edges_x, grid.edges_y, grid.edges_v, full,
                                  oc.ACC_FIXED, scales)
    assert np.array_equal(engine.fold_maps(grid, maps_s, fx).cpu().numpy().reshape(omaps.shape[1:])[..., [2, 0, 1]],
                          omaps[0][..., :3])
    assert np.array_equal(engine.finalize_cube(grid, cube_s, fx, scales).cpu().numpy(),
                          ocube[0].astype(np.float64) * (1.0 / scales[0]))


# ------------------------------------------------------------------------------- compiled-in limits
def test_maximum_sizes_and_limit_errors(pnm):
    """The largest grids the ABI accepts (PNM_MAX_EDGES - 1 = 4096 bins on an axis, PNM_MAX_TAPS taps,
    PNM_MAX_CHAIN compounded rotations) still match the oracle; one more is refused loudly."""
    from pynmap_b200 import _lib, engine
    rng = np.random.default_rng(12)
    n = 300_000
    x = rng.normal(0, 3, n).astype(np.float32)
    y = rng.normal(0, 3, n).astype(np.float32)
    d = rng.normal(0, 100, n).astype(np.float32)
    w = rng.uniform(0.5, 2, n).astype(np.float32)
    lim = [-6, 6, -6, 6]
    nmax = _lib.MAX_EDGES - 1
    for nxy in ([nmax, 8], [8, nmax]):
        _, _, M, V, S = pnm.grid.points_2vmaps(x, y, d, weights=w, limXY=lim, nXY=nxy, accumulate="fixed")
        _, _, Mo, Vo, So = on.points_2vmaps(x, y, d, weights=w, limXY=lim, nXY=nxy)
        assert M.shape == (nxy[1], nxy[0]) and same(M == 0, Mo == 0) and rel_err(M, Mo) < 1e-6
    lv = pnm.grid.points_2losvds(x, y, d, weights=w, limXY=lim, nXY=8, limV=[-400, 400], nV=nmax)
    assert sums_close(lv.losvd, on.points_2losvds(x, y, d, weights=w, limXY=lim, nXY=8, limV=[-400, 400], nV=nmax)[3])
    with pytest.raises(_lib.PnmError, match="edges"):
        pnm.grid.points_2vmaps(x, y, d, limXY=lim, nXY=[nmax + 1, 8])
    # widest smoothing kernel: support = PNM_MAX_TAPS
    cube = torch.from_numpy(rng.random((40, 36, 3))).cuda()
    taps = engine.gaussian_taps(_lib.MAX_TAPS / 8.0, _lib.MAX_TAPS)
    out = engine.smooth_cube(cube, taps, taps).cpu().numpy()
    ker = np.outer(taps, taps)
    ref = np.stack([on.convolve2d_fill0(cube[:, :, k].cpu().numpy(), ker) for k in range(3)], axis=2)
    assert rel_err(out, ref, floor=1e-12) < 1e-11
    with pytest.raises(_lib.PnmError, match="taps"):
        engine.smooth_cube(cube, np.ones(_lib.MAX_TAPS + 2) / (_lib.MAX_TAPS + 2), taps)
    # longest rotation chain: 8 compounded rotate(reset=False) calls, then one too many
    pos = rng.normal(0, 3, (5000, 3)).astype(np.float32)
    vel = rng.normal(0, 100, (5000, 3)).astype(np.float32)
    snap = pnm.snapshot.from_arrays(pos, vel, mass=np.ones(5000, dtype=np.float32))
    p_ref, v_ref = pos, vel
    for k in range(_lib.MAX_CHAIN):
        snap.rotate(PA=10. + k, inclination=20. + 3 * k, reset=False)
        p_ref, v_ref = on.rotate_mat(p_ref, v_ref, np.float32(10. + k), np.float32(20. + 3 * k))
    vm = snap.velocity_moments(limXY=lim, nXY=24, accumulate="fixed")
    _, _, Mo, Vo, So = on.points_2vmaps(p_ref[:, 0], p_ref[:, 1], v_ref[:, 2], limXY=lim, nXY=24)
    assert same(vm.M, Mo)                                   # counts: bit-exact through 8 roundings
    snap.rotate(PA=1., inclination=2., reset=False)         # a ninth: the chain is folded by materialising
    p_ref, v_ref = on.rotate_mat(p_ref, v_ref, np.float32(1.), np.float32(2.))
    vm9 = snap.velocity_moments(limXY=lim, nXY=24, accumulate="fixed")
    assert same(vm9.M, on.points_2vmaps(p_ref[:, 0], p_ref[:, 1], v_ref[:, 2], limXY=lim, nXY=24)[2])


def test_mixed_dtypes_follow_numpy_result_types(pnm):
    """float64 coordinates with float32 values: numpy rounds `wm * vm` and `vm**2` in float32
    (grid.py:128-131); the kernel carries everything as float64 and is told so (PNM_VAL_D32/W32)."""
    rng = np.random.default_rng(31)
    n = 40000
    x, y = rng.normal(0, 3, n), rng.normal(0, 3, n)                       # float64
    v32 = rng.normal(0, 150, n).astype(np.float32)
    w32 = rng.uniform(0.5, 2, n).astype(np.float32)
    lim = [-6, 6, -6, 6]
    for v, w in ((v32, w32), (v32, w32.astype(np.float64)), (v32, None), (v32.astype(np.float64), w32)):
        _, _, M, V, S = pnm.grid.points_2vmaps(x, y, v, weights=w, limXY=lim, nXY=20, accumulate="fixed")
        _, _, Mo, Vo, So = on.points_2vmaps(x, y, v, weights=w, limXY=lim, nXY=20)
        ok = Mo > 0
        assert rel_err(M, Mo) < 1e-12
        assert np.max(np.abs(V - Vo)[ok]) < 1e-10 * 150 and np.max(np.abs(S**2 - So**2)[ok & (So > 0)]) < 1e-10 * 150**2
        _, _, W, D, DW = pnm.grid.points_2map(x, y, data=v, weights=w, limXY=lim, nXY=20, accumulate="fixed")
        _, _, Wo, Do, DWo = on.points_2map(x, y, data=v, weights=w, limXY=lim, nXY=20)
        assert np.max(np.abs(DW - DWo)) < 1e-10 * np.max(np.abs(DWo))
    # the distinction is real: exact float64 products differ from numpy's float32 ones at the 1e-8 level
    _, _, _, V64, _ = on.points_2vmaps(x, y, v32.astype(np.float64), weights=w32.astype(np.float64), limXY=lim, nXY=20)
    _, _, _, V32, _ = on.points_2vmaps(x, y, v32, weights=w32, limXY=lim, nXY=20)
    assert np.nanmax(np.abs(V64 - V32)) > 1e-8


def test_tma_bulk_reduction_path_still_exact(pnm, monkeypatch):
    """The TMA bulk-reduction hand-off (off by default since lane pairs / quads made it redundant) stays selectable
    with PNM_TMA_SLOTS and exact: it serves the lane-pair mode (maps + plain cube, 16-byte records); maps-only runs use
    lane quads and ignore it."""
    import workloads
    from pynmap_b200 import _lib, engine
    n = 1_000_003
    soa = workloads.make_snapshot(n, seed=29, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    mat = on.rotation_matrix(np.float32(30.0), np.float32(60.0))
    grid = engine.get_grid(workloads.LIM_XY, 48, 40, workloads.LIM_V, 24)
    fx = engine.AccMode("fixed")
    scales = engine.fixed_scales(n, engine.absmax(soa[6]), 2.0 * max(engine.absmax(soa[k]) for k in (3, 4, 5)))
    rows = [soa[k] for k in range(6)]
    ref = {}
    for slots in ("0", "2", "4"):
        monkeypatch.setenv("PNM_TMA_SLOTS", slots)
        for sums in (7, 31):
            maps, cube = engine.new_accumulators(grid, sums, fx)
            engine.project_bin(grid, rows, soa[6], mat, 1, 1, sums, fx, scales, maps, cube)
            out = engine.finalize_vmaps(grid, maps, cube if sums & 8 else None, sums, fx, scales)
            got = [t.cpu().numpy() for t in out] + ([engine.finalize_cube(grid, cube, fx, scales, sums).cpu().numpy()] if sums & 8 else [])
            if slots == "0":
                ref[sums] = got
            else:
                assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, ref[sums])), (slots, sums)
    omaps, ocube = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, grid.edges_x, grid.edges_y, grid.edges_v, 31,
                                  oc.ACC_FIXED, scales)
    Mo, Vo, So = oc.finalize_vmaps(omaps[0], ocube[0], 31, oc.ACC_FIXED, scales)
    assert np.array_equal(ref[31][0], Mo) and np.array_equal(ref[31][1], Vo) and np.array_equal(ref[31][2], So, equal_nan=True)


def test_smoothing_nan_interpolate_matches_oracle(pnm):
    """convolve_spatial on a cube with NaN voxels (reachable: float64 accumulation propagates NaN weights into the cube):
    astropy's default nan_treatment='interpolate' -- NaN samples are skipped and the kernel renormalised per voxel; slices
    without NaNs are untouched by the repair path.  Against the C oracle (direct 2-D loop), 1e-10."""
    from pynmap_b200.losvd import LOSVD
    rng = np.random.default_rng(17)
    nx, ny, nv = 37, 29, 6
    cube = rng.uniform(0.0, 5.0, (nx, ny, nv))
    cube[5, 7, 0] = np.nan                       # isolated NaN
    cube[0, 0, 1] = np.nan                       # corner
    cube[10:14, 3:20, 2] = np.nan                # a block wider than the kernel in one direction
    cube[:, :, 3] = np.nan                       # a whole slice: outputs keep NaN (away from the border)
    cube[36, 28, 4] = np.nan; cube[35, 28, 4] = np.nan
    px, py, pv = np.linspace(-9, 9, nx), np.linspace(-7, 7, ny), np.linspace(-200, 200, nv)
    for fx, fy in ((1.5, 1.5), (1.0, 2.2)):
        lv = LOSVD(px, py, pv, torch.from_numpy(cube).cuda())
        sm = lv.convolve_spatial(fx, fy)
        k2 = on.gaussian_kernel2d(fx * on.FWHM_TO_SIGMA / lv.stepx, fy * on.FWHM_TO_SIGMA / lv.stepy)
        ref = oc.smooth_cube(cube, k2)
        got = sm.losvd.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        # an all-NaN slice stays NaN where the whole footprint lies inside the array; near the border the zero padding
        # counts as data (0 / sum of the padding's kernel values = 0)
        assert np.isnan(got[8:-8, 8:-8, 3]).all() and got[0, 0, 3] == 0.0 and not np.isnan(got[:, :, [0, 1, 4, 5]]).any()
        ok = ~np.isnan(ref)
        assert np.allclose(got[ok], ref[ok], rtol=1e-10, atol=1e-12)
        clean = oc.smooth_cube(cube[:, :, 5:6], k2)
        assert np.allclose(got[:, :, 5], clean[:, :, 0], rtol=1e-12, atol=1e-14)
        assert sm.binx is not lv.binx and np.array_equal(sm.binx, lv.binx)        # deep-copy semantics (losvd.py:118)
