"""GPU (-m gpu): the five BASELINE.json configurations at sizes the oracle finishes in seconds,
compared with the C oracle (fixed point: bit-exact; float64 atomics: 1e-12), plus float64 particle
arrays (what the reference's ascii reader produces) and the non-uniform-edge path of the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import oracle_c as oc
from oracle import oracle_np as on

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import pynmap_b200
    import workloads
    pynmap_b200._lib.require_device()
    return pynmap_b200, workloads


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= tol * np.abs(b) + 1e-300))


def test_config1_disk_256_maps(env):
    """C1: 10^6-particle exponential disk, rotate(30, 60), 256x256 M/V/sigma maps -- full, un-chunked."""
    pnm, wl = env
    soa = wl.make_snapshot(1_000_000, bulge_fraction=0.0, seed=1234, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="numpy")
    snap.rotate(PA=30., inclination=60.)
    vm = snap.velocity_moments(weight_name="mass", limXY=wl.LIM_XY, nXY=256)
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    ex = on.bin_edges(-15, 15, 256)
    omaps, _ = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, ex, ex, None, 7)
    Mo, Vo, So = oc.finalize_vmaps(omaps[0], None, 7)
    assert vm.M.shape == (256, 256)
    assert close(vm.M, Mo) and same(vm.M == 0, Mo == 0)
    assert rel_err(vm.V, Vo, floor=1e-3) < 1e-9
    ok = So > 1.0
    assert rel_err(vm.S[ok], So[ok]) < 1e-8
    # unweighted run: exact integer counts == bin indices bit-exact
    cnt = snap.velocity_moments(limXY=wl.LIM_XY, nXY=256)
    cmaps, _ = oc.project_bin(list(host[:6]), None, mat, 1, 1, ex, ex, None, 7)
    assert same(cnt.M, cmaps[0, :, :, 0])
    # fixed point: bit-exact against the oracle's quantisation
    fx = snap.velocity_moments(weight_name="mass", limXY=wl.LIM_XY, nXY=256, accumulate="fixed")
    from pynmap_b200 import engine
    scales = snap._fixed_scales(engine.AccMode("fixed"), snap.mass)
    fmaps, _ = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, ex, ex, None, 7, oc.ACC_FIXED, scales)
    Mf, Vf, Sf = oc.finalize_vmaps(fmaps[0], None, 7, oc.ACC_FIXED, scales)
    assert same(fx.M, Mf) and same(fx.V, Vf) and same(fx.S, Sf)


def test_config3_512_maps_and_shard_sum(env):
    """C3 at reduced N: 512x512 maps; two particle shards binned separately and summed give the
    single-pass bits in fixed point (what the NCCL reduce does across GPUs)."""
    pnm, wl = env
    from pynmap_b200 import engine
    from pynmap_b200.sharded import shard_bounds
    n = 2_000_003
    soa = wl.make_snapshot(n, seed=1236, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    grid = engine.get_grid(wl.LIM_XY, 512, 512)
    fx = engine.AccMode("fixed")
    scales = engine.fixed_scales(n, 2.0, 2.0 * float(np.abs(host[3:6]).max()))
    rows = [soa[k] for k in range(6)]
    whole, _ = engine.new_accumulators(grid, 7, fx)
    engine.project_bin(grid, rows, soa[6], mat, 1, 1, 7, fx, scales, whole, None)
    parts = []
    for r in range(2):
        lo, hi = shard_bounds(n, r, 2)
        m, _ = engine.new_accumulators(grid, 7, fx)
        engine.project_bin(grid, [t[lo:hi] for t in rows], soa[6][lo:hi], mat, 1, 1, 7, fx, scales, m, None)
        parts.append(engine.fold_maps(grid, m, fx).clone())
    assert torch.equal(parts[0] + parts[1], engine.fold_maps(grid, whole, fx))
    omaps, _ = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, grid.edges_x, grid.edges_y, None, 7, oc.ACC_FIXED, scales)
    got = engine.fold_maps(grid, whole, fx).cpu().numpy().reshape(512, 512, 4)
    assert np.array_equal(got[..., [2, 0, 1]], omaps[0][..., :3])


def test_config4_views_batched(env):
    """C4 at reduced N: 64 viewing angles (8 inclinations x 8 PAs) from one read == 64 single views."""
    pnm, wl = env
    n = 300_000
    soa = wl.make_snapshot(n, seed=1237, device="cuda")
    host = soa.cpu().numpy()
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="numpy")
    incs = np.repeat(np.linspace(0, 87.5, 8), 8)
    pas = np.tile(np.linspace(0, 157.5, 8), 8)
    views = snap.project_views(pas, incs, weight_name="mass", limXY=wl.LIM_XY, nXY=64)
    assert len(views) == 64
    # the same views binned 16 per read of the particles (pnm_project_bin with nviews = 16)
    batched = snap.project_views(pas[:32], incs[:32], weight_name="mass", limXY=wl.LIM_XY, nXY=64, views_per_launch=16)
    assert all(same(batched[k].M, views[k].M) and rel_err(batched[k].V, views[k].V, floor=1e-3) < 1e-9 for k in range(32))
    ex = on.bin_edges(-15, 15, 64)
    for k in (0, 9, 27, 63):
        mat = on.rotation_matrix(np.float32(pas[k]), np.float32(incs[k]))
        omaps, _ = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, ex, ex, None, 7)
        Mo, Vo, So = oc.finalize_vmaps(omaps[0], None, 7)
        assert same(views[k].M, Mo)                       # unit masses: exact counts per view
        assert rel_err(views[k].V, Vo, floor=1e-3) < 1e-9
        assert float(views[k].PAs[0]) == float(np.float32(pas[k]))


def test_config5_flux_weights_big_cube(env):
    """C5 at reduced N: flux-weighted create_map (256x256) and a 256x256x512 LOSVD cube, weights
    spanning the SSP M/L range (0.07 .. 8.6)."""
    pnm, wl = env
    n = 1_500_000
    soa = wl.make_snapshot(n, seed=1238, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(5)
    ml = 0.067 + (8.6 - 0.067) * torch.rand(n, generator=gen, device="cuda")
    flux = (soa[6] / ml).float()
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], flux=flux, output="numpy")
    snap.rotate(PA=30., inclination=60.)
    host, fh = soa.cpu().numpy(), flux.cpu().numpy()
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    rx, ry, _ = oc.rotate(mat, host[0], host[1], host[2])
    fm = snap.create_map(weight_name="lum", limXY=wl.LIM_XY, nXY=256)           # data = mass, weights = flux
    _, _, W, D, DW = on.points_2map(rx, ry, host[6], weights=fh, limXY=wl.LIM_XY, nXY=256)
    assert close(fm.W, W) and close(fm.DW, DW) and rel_err(fm.D, D, floor=1e-6) < 1e-9
    lv = snap.create_losvds(weight_name="lum", limXY=wl.LIM_XY, nXY=256, limV=wl.LIM_V, nV=512)
    ex, ev = on.bin_edges(-15, 15, 256), on.bin_edges(-500, 500, 512)
    _, ocube = oc.project_bin(list(host[:6]), fh, mat, 1, 1, ex, ex, ev, 8)
    assert lv.losvd.shape == (256, 256, 512)
    assert close(lv.losvd, ocube[0]) and same(lv.losvd == 0, ocube[0] == 0)
    sm = lv.convolve_spatial(0.5, 0.8)
    assert sm.losvd.shape == lv.losvd.shape and abs(sm.losvd.sum() / lv.losvd.sum() - 1) < 1e-2


def test_float64_particles_fused(env):
    """The reference's reader yields float64 positions / velocities / masses (np.loadtxt): the whole
    fused path in float64 arithmetic, 56 B per particle."""
    pnm, wl = env
    from pynmap_b200 import engine
    n = 400_001
    soa32 = wl.make_snapshot(n, seed=1239, device="cuda", unit_mass=False)
    jitter = torch.rand((7, n), generator=torch.Generator(device="cuda").manual_seed(1), device="cuda", dtype=torch.float64)
    soa = soa32.double() * (1.0 + 1e-9 * jitter)          # values that are NOT representable in float32
    host = soa.cpu().numpy()
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="numpy")
    snap.rotate(PA=12.5, inclination=77.)
    vm, lv = snap.project(weight_name="mass", limXY=wl.LIM_XY, nXY=96, limV=wl.LIM_V, nV=64, accumulate="fixed")
    mat = on.rotation_matrix(np.float32(12.5), np.float32(77.))
    ex, ev = on.bin_edges(-15, 15, 96), on.bin_edges(-500, 500, 64)
    scales = snap.last_fixed_scales
    fmaps, fcube = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, ex, ex, ev, 31, oc.ACC_FIXED, scales)
    Mf, Vf, Sf = oc.finalize_vmaps(fmaps[0], fcube[0], 31, oc.ACC_FIXED, scales)
    assert same(vm.M, Mf) and same(vm.V, Vf) and same(vm.S, Sf)
    assert same(lv.losvd, fcube[0].astype(np.float64) * (1.0 / scales[0]))
    assert same(snap.pos, np.stack(oc.rotate(mat, host[0], host[1], host[2]), 1))


def test_non_uniform_edges_through_the_abi(env):
    """pnm_grid_create accepts any increasing edges; far from an affine grid the guard band disables
    the uniform guess and every particle is decided by the exact table search."""
    pnm, wl = env
    from pynmap_b200 import _lib, engine
    rng = np.random.default_rng(11)
    ex = np.sort(rng.uniform(-3, 3, 41)); ey = np.concatenate([[-2.0], np.cumsum(rng.uniform(0.01, 0.3, 30)) - 2.0])
    ev = np.array([-100., -50., -49., 0., 0.5, 200.])
    n = 200_000
    pts = [rng.normal(0, 1.5, n).astype(np.float32), rng.normal(0, 1.5, n).astype(np.float32), rng.normal(0, 80, n).astype(np.float32)]
    pts[0][:41] = ex.astype(np.float32)                   # points on the (rounded) edges
    w = rng.uniform(0.5, 2, n).astype(np.float32)
    h = ctypes.c_void_p()
    _lib.call("pnm_grid_create", 40, _lib.doubles(ex), 30, _lib.doubles(ey), 5, _lib.doubles(ev), ctypes.byref(h))
    lib = _lib.load()
    R = lib.pnm_grid_map_replicas(h)
    maps = torch.zeros(lib.pnm_maps_acc_elems(h), dtype=torch.int64, device="cuda")
    cube = torch.zeros(lib.pnm_cube_acc_elems(h), dtype=torch.int64, device="cuda")
    dev = [torch.from_numpy(a).cuda() for a in pts] + [torch.from_numpy(w).cuda()]
    scales = engine.fixed_scales(n, 2.0, 500.0)
    P = engine.ptr
    _lib.call("pnm_bin_points", h, n, _lib.F32, P(dev[0]), P(dev[1]), P(dev[2]), P(dev[3]), None, 0, 15, _lib.ACC_FIXED,
              _lib.doubles(scales), P(maps), P(cube), engine.stream_ptr())
    _lib.call("pnm_fold_maps", h, P(maps), _lib.ACC_FIXED, 1, engine.stream_ptr())
    out = torch.empty((40, 30, 5), dtype=torch.float64, device="cuda")
    _lib.call("pnm_finalize_cube", h, P(cube), _lib.SUM_CUBE, _lib.ACC_FIXED, float(scales[0]), P(out), engine.stream_ptr())
    omaps, ocube = oc.project_bin(pts, w, None, 1, 1, ex, ey, ev, 15, oc.ACC_FIXED, scales, fused=False)
    got = maps[:40 * 30 * 4].cpu().numpy().reshape(30, 40, 4)
    assert R >= 1 and np.array_equal(got[..., [2, 0, 1]], omaps[0][..., :3])
    assert np.array_equal(out.cpu().numpy(), ocube[0].astype(np.float64) * (1.0 / scales[0]))
    lib.pnm_grid_destroy(h)


def test_host_path_single_and_sharded_flavours(env):
    """hostpath.project_host: the single-call C-ABI path (pnm_project_host) and the sharded path
    (pnm_accumulate_host + reducer + finalise) give the same answer as the device-resident API,
    with smoothing, pinned torch buffers or plain numpy arrays as inputs."""
    pnm, wl = env
    from pynmap_b200 import hostpath
    n = 5_000_011                                        # two staging chunks, ragged
    soa = wl.make_snapshot(n, seed=1240, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    pinned = torch.empty((7, n), dtype=torch.float32, pin_memory=True)
    pinned.copy_(soa)
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="numpy")
    snap.rotate(PA=30., inclination=60.)
    vm, lv = snap.project(weight_name="mass", limXY=wl.LIM_XY, nXY=64, limV=wl.LIM_V, nV=32, accumulate="f64-exact")   # every term a float64 reduction, like the host path
    sm = lv.convolve_spatial(1.0, 1.0)
    single = hostpath.project_host([host[k] for k in range(7)], 30., 60., wl.LIM_XY, 64, wl.LIM_V, 32, smooth_fwhm=(1.0, 1.0))
    calls = []
    def fake_reducer(*tensors):                          # world size 1: nothing to add, rank 0 holds the total
        calls.append([t.numel() for t in tensors])
        return True
    sharded = hostpath.project_host([pinned[k] for k in range(7)], 30., 60., wl.LIM_XY, 64, wl.LIM_V, 32,
                                    smooth_fwhm=(1.0, 1.0), reducer=fake_reducer)
    assert calls == [[64 * 64 * (16 + 1) * 4]]              # one span: the interleaved cube (nV/2 + 1 record slabs of 4)
    for res in (single, sharded):
        assert close(res["M"], vm.M) and rel_err(res["V"], vm.V, floor=1e-3) < 1e-9
        assert np.allclose(res["cube"], sm.losvd, rtol=1e-10, atol=1e-12 * sm.losvd.max())
    unsmoothed = hostpath.project_host([host[k] for k in range(7)], 30., 60., wl.LIM_XY, 64, wl.LIM_V, 32)
    assert close(unsmoothed["cube"], lv.losvd)
    with pytest.raises(ValueError):
        hostpath.project_host([host[k] for k in range(7)], 30., 60., wl.LIM_XY, 64, wl.LIM_V, 32, accumulate="fixed")
    pnm._lib.call("pnm_host_release")


def test_sharded_project_with_post_stream_matches_plain(env):
    """snapshot.project through the sharded code path (fold without clearing, reduce hook, finalise
    from the folded replica) on a post-processing stream == the plain single-GPU call, and two
    half-snapshots summed by the hook == the whole snapshot (fixed point: bit-exact)."""
    import pynmap_b200 as pnm
    import workloads
    n = 3_000_000
    soa = workloads.make_snapshot(n, seed=91, device="cuda")
    kw = dict(weight_name="mass", limXY=workloads.LIM_XY, nXY=96, limV=workloads.LIM_V, nV=64, accumulate="fixed")

    def snap_of(lo, hi):
        s = pnm.snapshot.from_arrays(soa[:3, lo:hi], soa[3:6, lo:hi], mass=soa[6, lo:hi], output="torch")
        s.rotate(PA=30., inclination=60.)
        return s

    whole = snap_of(0, n)
    vm0, lv0 = whole.project(**kw)
    post = torch.cuda.Stream(priority=-1)
    seen = []

    def hook(*tensors):                                   # world size 1: rank 0 already holds the total
        seen.append(sum(t.numel() for t in tensors))
        return True
    hook.n_total = lambda n_local: n
    hook.reduce_max = lambda *v: v
    vm1, lv1 = whole.project(reducer=hook, post_stream=post, **kw)
    post.synchronize()
    assert seen == [96 * 96 * (32 + 1) * 4]               # one span: the interleaved cube (nV/2 + 1 record slabs)
    for a, b in ((vm0.M, vm1.M), (vm0.V, vm1.V), (vm0.S, vm1.S), (lv0.losvd, lv1.losvd)):
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))

    # two "ranks": the second adds the first one's span before finalising
    half = n // 2 + 5
    parts = {}

    def hook_a(*tensors):
        parts["a"] = [t.clone() for t in tensors]
        return False
    hook_a.n_total = lambda n_local: n
    # what the ranks would agree on: global bounds (2 values), typical term sizes (3 values), the layout vote (1 value)
    hook_a.reduce_max = lambda *v: ((max(v[0], bounds[0]), max(v[1], bounds[1])) if len(v) == 2 else
                                    (1.0, 128.0, 16384.0) if len(v) == 3 else v)

    def hook_b(*tensors):
        for t, other in zip(tensors, parts["a"]):
            t += other
        return True
    hook_b.n_total = hook_a.n_total
    hook_b.reduce_max = hook_a.reduce_max
    from pynmap_b200 import engine
    bounds = (engine.absmax(soa[6]), 2.0 * max(engine.absmax(soa[k]) for k in (3, 4, 5)))
    assert snap_of(0, half).project(reducer=hook_a, **kw) == (None, None)
    vm2, lv2 = snap_of(half, n).project(reducer=hook_b, **kw)
    whole2 = snap_of(0, n)
    hook.reduce_max = hook_a.reduce_max
    vm3, lv3 = whole2.project(reducer=hook, **kw)         # same global bounds -> same fixed-point scales
    for a, b in ((vm3.M, vm2.M), (vm3.V, vm2.V), (vm3.S, vm2.S), (lv3.losvd, lv2.losvd)):
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))


def test_ordered_snapshot_uses_plain_layout_same_results(env):
    """A Morton-ordered copy of a snapshot takes the plain cube + replicated maps (run merging), the unordered one
    the interleaved cube; in fixed point both give the same bits."""
    import pynmap_b200 as pnm
    import workloads
    from pynmap_b200 import _lib, engine
    n = 2_000_000
    soa = workloads.make_snapshot(n, seed=17, device="cuda")
    srt = workloads.morton_sort(soa)
    kw = dict(weight_name="mass", limXY=workloads.LIM_XY, nXY=64, limV=workloads.LIM_V, nV=48, accumulate="fixed")
    out = []
    for data, want_ordered in ((soa, False), (srt, True)):
        s = pnm.snapshot.from_arrays(data[:3], data[3:6], mass=data[6], output="torch")
        s.rotate(PA=30., inclination=60.)
        grid = engine.get_grid(workloads.LIM_XY, 64, 64, workloads.LIM_V, 48)
        assert s._memory_ordered(grid) is want_ordered
        out.append(s.project(privatise=False, **kw))      # (the privatised kernel would pick coarser fixed-point scales)
    (vm_a, lv_a), (vm_b, lv_b) = out
    for a, b in ((vm_a.M, vm_b.M), (vm_a.V, vm_b.V), (vm_a.S, vm_b.S), (lv_a.losvd, lv_b.losvd)):
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))


def test_interleaved_cube_odd_and_tiny_velocity_axes(env):
    """snapshot.project on an unordered snapshot takes the interleaved cube (records of two velocity bins):
    odd nV, nV = 2 and a V range that leaves most particles outside it, against the C oracle (fixed point: bit-exact)."""
    import pynmap_b200 as pnm
    import workloads
    from pynmap_b200 import engine
    n = 400_003
    soa = workloads.make_snapshot(n, seed=23, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="numpy")
    snap.rotate(PA=30., inclination=60.)
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    lim = workloads.LIM_XY
    for nv, limv in ((21, [-500., 500.]), (2, [-50., 50.]), (7, [100., 160.])):
        grid = engine.get_grid(lim, 40, 24, limv, nv)
        assert not snap._memory_ordered(grid)
        vm, lv = snap.project(weight_name="mass", limXY=lim, nXY=[40, 24], limV=limv, nV=nv, accumulate="fixed")
        scales = snap.last_fixed_scales
        ex, ey, ev = on.bin_edges(lim[0], lim[1], 40), on.bin_edges(lim[2], lim[3], 24), on.bin_edges(limv[0], limv[1], nv)
        omaps, ocube = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, ex, ey, ev, 15 | 16, oc.ACC_FIXED, scales)
        Mo, Vo, So = oc.finalize_vmaps(omaps[0], ocube[0], 15 | 16, oc.ACC_FIXED, scales)
        assert np.array_equal(vm.M, Mo) and np.array_equal(vm.V, Vo) and np.array_equal(vm.S, So, equal_nan=True)
        assert np.array_equal(lv.losvd, ocube[0].astype(np.float64) * (1.0 / scales[0]))
        assert lv.losvd.shape == (40, 24, nv) and 0 < lv.losvd.sum() <= vm.M.sum() * (1 + 1e-12)


def test_interleaved_cube_two_views_in_one_launch(env):
    """pnm_project_bin with nviews = 2 and PNM_CUBE_MOMENTS (general kernel path, per-view strides of the interleaved
    layout) == two single-view launches, bit for bit in fixed point."""
    import workloads
    from pynmap_b200 import _lib, engine
    n = 300_000
    soa = workloads.make_snapshot(n, seed=41, device="cuda", unit_mass=False)
    rows = [soa[k] for k in range(6)]
    grid = engine.get_grid(workloads.LIM_XY, 32, 48, workloads.LIM_V, 19)
    fx = engine.AccMode("fixed")
    scales = engine.fixed_scales(n, engine.absmax(soa[6]), 2.0 * max(engine.absmax(soa[k]) for k in (3, 4, 5)))
    sums = 31 | _lib.CUBE_MOMENTS
    mats = [on.rotation_matrix(np.float32(30.0), np.float32(60.0)), on.rotation_matrix(np.float32(75.0), np.float32(20.0))]
    maps2, cube2 = engine.new_accumulators(grid, sums, fx, nviews=2)
    engine.project_bin(grid, rows, soa[6], np.stack(mats), 1, 2, sums, fx, scales, maps2, cube2)
    for k, mat in enumerate(mats):
        maps1, cube1 = engine.new_accumulators(grid, sums, fx)
        engine.project_bin(grid, rows, soa[6], mat, 1, 1, sums, fx, scales, maps1, cube1)
        nc = grid.cube_moments_elems
        assert maps1 is None and maps2 is None and torch.equal(cube2[k * nc:(k + 1) * nc], cube1)   # everything lives in the cube records
        got = engine.finalize_vmaps(grid, None, cube2[k * nc:(k + 1) * nc], sums, fx, scales)
        want = engine.finalize_vmaps(grid, None, cube1, sums, fx, scales)
        assert all(torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)) for a, b in zip(got, want))


def test_pipelined_project_recycles_accumulators_correctly(env):
    """Several back-to-back project(post_stream=...) calls (accumulators recycled and cleared on the post-processing
    stream) each return the same bits as a plain call, for both accumulator layouts."""
    import pynmap_b200 as pnm
    import workloads
    n = 1_500_000
    soa = workloads.make_snapshot(n, seed=57, device="cuda")
    kw = dict(weight_name="mass", limXY=workloads.LIM_XY, nXY=64, limV=workloads.LIM_V, nV=40, accumulate="fixed")
    post = torch.cuda.Stream(priority=-1)
    for data in (soa, workloads.morton_sort(soa)):
        snap = pnm.snapshot.from_arrays(data[:3], data[3:6], mass=data[6], output="torch")
        snap.rotate(PA=30., inclination=60.)
        vm0, lv0 = snap.project(**kw)
        outs = [snap.project(post_stream=post, **kw) for _ in range(6)]
        post.synchronize()
        torch.cuda.synchronize()
        for vm, lv in outs:
            for a, b in ((vm0.M, vm.M), (vm0.V, vm.V), (vm0.S, vm.S), (lv0.losvd, lv.losvd)):
                assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
        assert len(next(iter(snap._acc_pool.values()))) == 2          # two sets in rotation


def test_velocity_moments_with_reducer_hook(env):
    """velocity_moments(reducer=...) (BASELINE config 3: maps only, partial grids summed over ranks): the reducer sees
    ONE folded map replica; two particle shards whose partial grids are added give the bits of the whole snapshot with
    int64 accumulators; a rank that does not hold the total gets None."""
    pnm, wl = env
    n = 300_000
    soa = wl.make_snapshot(n, seed=91, device="cuda", unit_mass=False)
    bounds = (float(soa[6].abs().max()), 2.0 * max(float(soa[k].abs().max()) for k in (3, 4, 5)))
    kw = dict(weight_name="mass", limXY=wl.LIM_XY, nXY=[72, 56], accumulate="fixed")

    def snap_of(lo, hi):
        s = pnm.snapshot.from_arrays(soa[:3, lo:hi], soa[3:6, lo:hi], mass=soa[6, lo:hi], output="torch")
        s.rotate(PA=30., inclination=60.)
        return s

    class Hook(object):
        def __init__(self, keep=None, add=None):
            self.keep, self.add, self.seen = keep, add, []
        def n_total(self, n_local): return n
        def reduce_max(self, *v): return bounds
        def __call__(self, *tensors):
            self.seen.append([t.numel() for t in tensors])
            if self.keep is not None:
                self.keep.extend(t.clone() for t in tensors)
                return False
            for t, o in zip(tensors, self.add or []):
                t += o
            return True

    whole_hook = Hook()
    whole = snap_of(0, n).velocity_moments(reducer=whole_hook, **kw)
    assert whole_hook.seen == [[72 * 56 * 4]]                 # one folded replica of 4 slots per pixel
    kept = []
    assert snap_of(0, 120_008).velocity_moments(reducer=Hook(keep=kept), **kw) is None
    both = snap_of(120_008, n).velocity_moments(reducer=Hook(add=kept), **kw)
    for a, b in ((both.M, whole.M), (both.V, whole.V), (both.S, whole.S)):
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
    assert float(whole.M.sum()) > 0
