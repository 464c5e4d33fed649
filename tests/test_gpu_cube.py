"""GPU (-m gpu): the privatised scatter kernel k_project_cube (csrc/pnm_cube.cuh) behind snapshot.project --
the kernel bench.py times.  int64 accumulators: bit-exact against the C oracle for any plan; float64 accumulators:
within the 1e-6 contract (every privatised term is rounded by less than 4.8e-7 of itself).  Reference semantics:
rotation.py:115-117, grid.py:126-131, grid.py:252-259."""
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


def moments_within_contract(V, S, Vo, So, tol=1e-6):
    """k_project_cube (float64 accumulators) rounds every privatised term by less than 4.8e-7 of itself, so each
    pixel's sum w*d is within 4.8e-7 * sum |w*d| <= 4.8e-7 * sqrt(S0 * S2) of the float64 sum (Cauchy-Schwarz):
    V within 1e-6 of the pixel's rms line-of-sight velocity sqrt(V^2 + sigma^2) -- which IS |V| for a pixel with one
    particle -- and sigma^2 = S2/S0 - V^2 within a few 1e-6 of rms^2."""
    V, S, Vo, So = (np.asarray(a, dtype=np.float64) for a in (V, S, Vo, So))
    assert np.array_equal(np.isnan(V), np.isnan(Vo))
    rms2 = Vo ** 2 + np.nan_to_num(So) ** 2
    ok = ~np.isnan(Vo)
    assert (np.abs(V - Vo)[ok] <= tol * np.sqrt(rms2[ok])).all(), float(np.max(np.abs(V - Vo)[ok] / np.sqrt(rms2[ok]).clip(1e-300)))
    okS = ok & np.isfinite(So) & np.isfinite(S)
    assert (np.abs(S ** 2 - So ** 2)[okS] <= 4 * tol * rms2[okS]).all()
    return True


def tsame(a, b):
    return torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))


def oracle_fixed(host, w, mat, grid, scales, sums=15 | 16):
    omaps, ocube = oc.project_bin(list(host[:6]), w, mat, 1, 1, grid.edges_x, grid.edges_y, grid.edges_v, sums,
                                  oc.ACC_FIXED, scales)
    M, V, S = oc.finalize_vmaps(omaps[0], ocube[0], sums, oc.ACC_FIXED, scales)
    return M, V, S, ocube[0].astype(np.float64) * (1.0 / scales[0])


def nasty_weights(w, rng):
    """Weights that exercise every exit of the shared-memory path: NaN, +-inf, huge (scaled term beyond 32 bits),
    tiny (below the float64 mode's floor), negative and zero."""
    w = w.copy()
    idx = rng.choice(w.size, 600, replace=False)
    w[idx[:100]] = np.nan
    w[idx[100:150]] = np.inf
    w[idx[150:200]] = -np.inf
    w[idx[200:300]] *= 3.0e4
    w[idx[300:400]] *= 1.0e-6
    w[idx[400:500]] *= -1.0
    w[idx[500:600]] = 0.0
    return w


@pytest.mark.parametrize("n,nxy,nv", [(2_000_003, (128, 128), 256), (300_001, (40, 24), 21), (5_000, (16, 16), 2), (1023, (8, 8), 7)])
def test_cube_kernel_fixed_is_bit_exact(env, n, nxy, nv):
    """snapshot.project(accumulate='fixed') takes k_project_cube (checked) and equals the oracle bit for bit: ragged
    particle counts (vector loop + scalar end), odd / tiny velocity axes, non-square maps, nasty weights, unit weights."""
    pnm, wl = env
    from pynmap_b200 import engine
    rng = np.random.default_rng(n)
    soa = wl.make_snapshot(n, seed=77, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    host[6] = nasty_weights(host[6], rng) if n > 2000 else host[6]
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    lim, limv = wl.LIM_XY, ([-50., 50.] if nv == 2 else wl.LIM_V)
    grid = engine.get_grid(lim, nxy[0], nxy[1], limv, nv)
    for weights in (host[6], None):
        snap = pnm.snapshot.from_arrays(host[:3].T.copy(), host[3:6].T.copy(), mass=weights, output="numpy")
        snap.rotate(PA=30., inclination=60.)
        timer = engine.KERNEL_TIMER = engine.KernelTimer()
        try:
            wkw = dict(weight_name="mass") if weights is not None else {}        # no weights: the HAS_W = false kernel
            vm, lv = snap.project(limXY=lim, nXY=list(nxy), limV=limv, nV=nv, accumulate="fixed", **wkw)
        finally:
            engine.KERNEL_TIMER = None
        assert len(snap._cube_plans) == 1 and len(timer.pairs) == 1          # the privatised kernel ran, once
        scales = snap.last_fixed_scales
        Mo, Vo, So, co = oracle_fixed(host, weights, mat, grid, scales)
        assert same(vm.M, Mo) and same(vm.V, Vo) and same(vm.S, So) and same(lv.losvd, co)
        # the plain reduction kernel with the same scales gives the same bits (the plan never changes what is summed)
        acc = engine.AccMode("fixed")
        pos, vel = snap._base
        soa6 = (pos[0], pos[1], pos[2], vel[0], vel[1], vel[2])
        sums = 31 | pnm._lib.CUBE_MOMENTS
        w_dev = snap.mass if weights is not None else None
        _, cube = engine.new_accumulators(grid, sums, acc)
        engine.project_bin(grid, soa6, w_dev, mat, 1, 1, sums, acc, scales, None, cube)
        got = engine.finalize_vmaps(grid, None, cube, sums, acc, scales)
        assert same(got[0].cpu().numpy(), Mo) and same(got[1].cpu().numpy(), Vo) and same(got[2].cpu().numpy(), So)


def test_cube_kernel_f64_within_contract(env):
    """accumulate='f64' through k_project_cube against the oracle's sequential float64 sums: per pixel / voxel within
    1e-6 relative (measured ~3e-8), zeros and NaNs in the same places; 'f64-exact' keeps 1e-12."""
    pnm, wl = env
    from pynmap_b200 import engine
    n = 3_000_007
    soa = wl.make_snapshot(n, seed=78, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    rng = np.random.default_rng(5)
    w = host[6].copy()
    idx = rng.choice(n, 300, replace=False)
    w[idx[:100]] *= 3.0e4
    w[idx[100:200]] *= 1.0e-6
    w[idx[200:300]] = 0.0
    host[6] = w
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    lim, limv, nxy, nv = wl.LIM_XY, wl.LIM_V, 128, 64
    grid = engine.get_grid(lim, nxy, nxy, limv, nv)
    snap = pnm.snapshot.from_arrays(host[:3].T.copy(), host[3:6].T.copy(), mass=host[6], output="numpy")
    snap.rotate(PA=30., inclination=60.)
    sums = 15 | 16
    omaps, ocube = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, grid.edges_x, grid.edges_y, grid.edges_v, sums)
    Mo, Vo, So = oc.finalize_vmaps(omaps[0], ocube[0], sums)
    vm, lv = snap.project(weight_name="mass", limXY=lim, nXY=nxy, limV=limv, nV=nv)
    assert len(snap._cube_plans) == 1
    assert same(vm.M == 0, Mo == 0) and same(lv.losvd == 0, ocube[0] == 0)
    assert rel_err(vm.M, Mo) < 1e-6 and rel_err(lv.losvd, ocube[0]) < 1e-6            # contract; sums of positive terms
    # V = S1/S0 and sigma: sums of signed terms -> relative to the pixel's rms velocity (see moments_within_contract)
    assert moments_within_contract(vm.V, vm.S, Vo, So)
    ok = np.isfinite(So) & (So > 5.0) & np.isfinite(vm.S)
    assert rel_err(vm.S[ok], So[ok]) < 1e-5
    vx, lx = snap.project(weight_name="mass", limXY=lim, nXY=nxy, limV=limv, nV=nv, accumulate="f64-exact")
    assert rel_err(vx.M, Mo) < 1e-12 and rel_err(lx.losvd, ocube[0]) < 1e-12 and rel_err(vx.V, Vo, floor=1e-3) < 1e-9


def test_cube_kernel_stale_plan_and_changed_weights_stay_correct(env):
    """A plan built for other weights / another view only costs speed: weights scaled by 1e9 or 1e-9 in place after
    the plan was built (terms no longer fit the shared-memory limbs, or fall below the float64 floor) still give
    oracle-exact sums, and the fixed-point scales follow the new weights (nothing is cached on addresses)."""
    pnm, wl = env
    from pynmap_b200 import engine
    n = 400_000
    soa = wl.make_snapshot(n, seed=79, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    lim, limv, nxy, nv = wl.LIM_XY, wl.LIM_V, 48, 32
    grid = engine.get_grid(lim, nxy, nxy, limv, nv)
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], output="numpy")
    snap.rotate(PA=30., inclination=60.)
    w = soa[6].clone()
    for factor in (1.0, 1.0e9, 1.0e-18):
        w.mul_(factor)                                        # same tensor, same address, new contents
        wh = w.cpu().numpy()
        vm, lv = snap.project(weights=w, limXY=lim, nXY=nxy, limV=limv, nV=nv, accumulate="fixed")
        Mo, Vo, So, co = oracle_fixed(host, wh, mat, grid, snap.last_fixed_scales)
        assert same(vm.M, Mo) and same(vm.V, Vo) and same(vm.S, So) and same(lv.losvd, co)
        assert vm.M.sum() > 0
        vf, lf = snap.project(weights=w, limXY=lim, nXY=nxy, limV=limv, nV=nv)
        omaps, ocube = oc.project_bin(list(host[:6]), wh, mat, 1, 1, grid.edges_x, grid.edges_y, grid.edges_v, 31)
        Mq, _, _ = oc.finalize_vmaps(omaps[0], ocube[0], 31)
        assert rel_err(lf.losvd, ocube[0]) < 1e-6 and rel_err(vf.M, Mq) < 1e-6
    assert len(snap._cube_plans) == 1                         # one plan (built for the first contents) served all calls
    # velocity_moments (maps only, plain kernel): scales are measured again for weights the caller passes in
    a = snap.velocity_moments(weights=soa[6], limXY=lim, nXY=nxy, accumulate="fixed")
    b = snap.velocity_moments(weights=soa[6] * 1.0e6, limXY=lim, nXY=nxy, accumulate="fixed")
    assert np.allclose(b.M, a.M * 1.0e6, rtol=1e-6) and np.allclose(b.V, a.V, rtol=1e-6, atol=1e-6, equal_nan=True)


def test_cube_kernel_view_change_and_compound_rotation(env):
    """A new view gets its own plan; a compounded rotation (rotate(reset=False)) falls back to the plain kernel, which
    re-applies the chain per particle -- both against the oracle."""
    pnm, wl = env
    from pynmap_b200 import engine
    n = 500_000
    soa = wl.make_snapshot(n, seed=80, device="cuda", unit_mass=False)
    host = soa.cpu().numpy()
    lim, limv, nxy, nv = wl.LIM_XY, wl.LIM_V, 64, 40
    grid = engine.get_grid(lim, nxy, nxy, limv, nv)
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="numpy")
    for k, (pa, inc) in enumerate(((30., 60.), (75., 20.), (0., 0.))):
        snap.rotate(PA=pa, inclination=inc)
        vm, lv = snap.project(weight_name="mass", limXY=lim, nXY=nxy, limV=limv, nV=nv, accumulate="fixed")
        assert len(snap._cube_plans) == k + 1
        mat = on.rotation_matrix(np.float32(pa), np.float32(inc))
        Mo, Vo, So, co = oracle_fixed(host, host[6], mat, grid, snap.last_fixed_scales)
        assert same(vm.M, Mo) and same(vm.V, Vo) and same(vm.S, So) and same(lv.losvd, co)
    snap.rotate(PA=30., inclination=60.)
    snap.rotate(PA=10., inclination=15., reset=False)
    nplans = len(snap._cube_plans)
    vm, lv = snap.project(weight_name="mass", limXY=lim, nXY=nxy, limV=limv, nV=nv, accumulate="fixed")
    assert len(snap._cube_plans) == nplans                    # no plan: the chain runs through k_project_bin
    m1, m2 = on.rotation_matrix(np.float32(30.), np.float32(60.)), on.rotation_matrix(np.float32(10.), np.float32(15.))
    scales = snap.last_fixed_scales
    omaps, ocube = oc.project_bin(list(host[:6]), host[6], np.stack([m1, m2]), 2, 1, grid.edges_x, grid.edges_y, grid.edges_v,
                                  15 | 16, oc.ACC_FIXED, scales)
    Mo, Vo, So = oc.finalize_vmaps(omaps[0], ocube[0], 15 | 16, oc.ACC_FIXED, scales)
    assert same(vm.M, Mo) and same(vm.V, Vo) and same(vm.S, So)


def test_bench_path_full_size_against_oracle(env):
    """The exact path bench.py times -- snapshot.project on 10^8 unordered float32 particles, 128 x 128 maps +
    128 x 128 x 256 cube, privatised kernel, float64 AND int64 accumulators -- checked against the C oracle on a
    2 x 10^6-particle slice and through size-independent properties on the whole set: two particle shards summed ==
    the whole snapshot bit for bit (what an N-rank run relies on), float64 == int64 where both are exact (unit
    masses: counts), and float64 moments within the quantisation bound of the int64 ones."""
    pnm, wl = env
    from pynmap_b200 import engine
    n, m = 100_000_000, 2_000_000
    soa = wl.make_snapshot(n, seed=1235, device="cuda")
    lim, limv, nxy, nv = wl.LIM_XY, wl.LIM_V, 128, 256
    grid = engine.get_grid(lim, nxy, nxy, limv, nv)
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    kw = dict(weight_name="mass", limXY=lim, nXY=nxy, limV=limv, nV=nv)
    bounds = (1.0, 2.0 * max(engine.absmax(soa[k]) for k in (3, 4, 5)))

    def snap_of(lo, hi):
        s = pnm.snapshot.from_arrays(soa[:3, lo:hi], soa[3:6, lo:hi], mass=soa[6, lo:hi], output="torch")
        s.rotate(PA=30., inclination=60.)
        return s

    class Hook(object):                                       # stands in for sharded.GridReducer: same scales on every "rank"
        def __init__(self, keep=None, add=None):
            self.keep, self.add = keep, add
        def n_total(self, n_local):
            return n
        def reduce_max(self, *v):
            return bounds if len(v) == 2 else (1.0, 128.0, 16384.0) if len(v) == 3 else v
        def __call__(self, *tensors):
            if self.keep is not None:
                self.keep.extend(t.clone() for t in tensors)
                return False
            if self.add is not None:
                for t, o in zip(tensors, self.add):
                    t += o
            return True

    whole = snap_of(0, n)
    assert not whole._memory_ordered(grid)
    vm_w, lv_w = whole.project(accumulate="fixed", reducer=Hook(), **kw)
    assert len(whole._cube_plans) == 1
    scales = whole.last_fixed_scales
    # (a) a slice against the oracle, bit for bit
    part = snap_of(0, m)
    vm_p, lv_p = part.project(accumulate="fixed", reducer=Hook(), **kw)
    assert part.last_fixed_scales == scales
    host = soa[:, :m].cpu().numpy()
    Mo, Vo, So, co = oracle_fixed(host, host[6], mat, grid, scales)
    assert same(vm_p.M.cpu().numpy(), Mo) and same(vm_p.V.cpu().numpy(), Vo) and same(vm_p.S.cpu().numpy(), So)
    assert same(lv_p.losvd.cpu().numpy(), co)
    # (b) two shards (ragged split), partial grids summed == the whole snapshot
    kept = []
    half = n // 2 + 1029
    assert snap_of(0, half).project(accumulate="fixed", reducer=Hook(keep=kept), **kw) == (None, None)
    vm_s, lv_s = snap_of(half, n).project(accumulate="fixed", reducer=Hook(add=kept), **kw)
    assert tsame(vm_s.M, vm_w.M) and tsame(vm_s.V, vm_w.V) and tsame(vm_s.S, vm_w.S) and torch.equal(lv_s.losvd, lv_w.losvd)
    del kept
    # (c) float64 accumulators through the same kernel
    vm_f, lv_f = whole.project(accumulate="f64", **kw)
    assert torch.equal(vm_f.M, vm_w.M) and torch.equal(lv_f.losvd, lv_w.losvd)       # unit masses: exact counts in both
    assert 0.9 * n < float(vm_f.M.sum()) <= n and float(lv_f.losvd.sum()) <= float(vm_f.M.sum())
    assert float(((vm_f.V - vm_w.V).abs() * 1.0).nan_to_num().max()) < 1e-3          # km/s; typical |V| ~ 100
    okS = torch.isfinite(vm_w.S) & torch.isfinite(vm_f.S) & (vm_w.S > 5.0)
    assert float(((vm_f.S - vm_w.S).abs() / vm_w.S)[okS].max()) < 1e-4               # sqrt(S2/S0 - V^2): cancellation amplifies 1e-7
    # (d) and the oracle's float64 sums on the slice
    vm_q, lv_q = part.project(accumulate="f64", **kw)
    omaps, ocube = oc.project_bin(list(host[:6]), host[6], mat, 1, 1, grid.edges_x, grid.edges_y, grid.edges_v, 15 | 16)
    Mq, Vq, Sq = oc.finalize_vmaps(omaps[0], ocube[0], 15 | 16)
    assert same(vm_q.M.cpu().numpy(), Mq) and same(lv_q.losvd.cpu().numpy(), ocube[0])
    assert moments_within_contract(vm_q.V.cpu().numpy(), vm_q.S.cpu().numpy(), Vq, Sq)


def test_slab_finalisation_equals_whole_accumulator(env):
    """pnm_slab_sums / pnm_finalize_sums / pnm_finalize_cube_slabs on the chunks a reduce-scatter over 1, 2, 3 and 8
    ranks would leave == pnm_finalize_vmaps / pnm_finalize_cube on the whole interleaved accumulator: bit for bit with
    int64 accumulators (odd and even velocity axes; ranks that only hold padding), to rounding with float64."""
    pnm, wl = env
    from pynmap_b200 import engine, _lib
    n = 600_000
    soa = wl.make_snapshot(n, seed=81, device="cuda", unit_mass=False)
    rows = [soa[k] for k in range(6)]
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    sums = 31 | _lib.CUBE_MOMENTS
    for nv in (33, 8, 2):
        grid = engine.get_grid(wl.LIM_XY, 40, 24, wl.LIM_V, nv)
        for mode in ("fixed", "f64"):
            acc = engine.AccMode(mode)
            scales = engine.fixed_scales(n, engine.absmax(soa[6]), 2.0 * max(engine.absmax(r) for r in rows[3:])) if mode == "fixed" else None
            for world in (1, 2, 3, 8):
                _, cube = engine.new_accumulators(grid, sums, acc, slab_multiple=world)
                c = -(-grid.record_slabs // world)
                assert cube.numel() == c * world * grid.slab_elems
                engine.project_bin(grid, rows, soa[6], mat, 1, 1, sums, acc, scales, None, cube)
                Mw, Vw, Sw = engine.finalize_vmaps(grid, None, cube, sums, acc, scales)
                cw = engine.finalize_cube(grid, cube, acc, scales, sums)
                per = c * grid.slab_elems
                total, parts = None, []
                for r in range(world):
                    chunk = cube[r * per:(r + 1) * per]
                    share = engine.slab_sums(grid, chunk, r * c, (r + 1) * c, acc)
                    total = share if total is None else total + share
                    parts.append(engine.finalize_cube_slabs(grid, chunk, r * c, (r + 1) * c, acc, scales))
                    assert parts[-1].shape[2] == max(0, min(nv, 2 * (r + 1) * c) - min(nv, 2 * r * c))
                M, V, S = engine.finalize_sums(grid, total, acc, scales)
                assert torch.equal(torch.cat(parts, dim=2), cw)
                if mode == "fixed":
                    assert tsame(M, Mw) and tsame(V, Vw) and tsame(S, Sw)
                else:
                    assert torch.allclose(M, Mw, rtol=1e-13, atol=0) and torch.allclose(V, Vw, rtol=1e-9, atol=1e-9, equal_nan=True)


def test_project_through_slab_reducer_single_rank(env):
    """snapshot.project(reducer=SlabReducer-like) with one rank: the reduce-scatter path (padded accumulator, slab sums,
    finalisation of the own velocity bins, smoothing) returns the plain project() bits, also pipelined on a side stream."""
    pnm, wl = env
    n = 700_001
    soa = wl.make_snapshot(n, seed=82, device="cuda", unit_mass=False)

    class OneRank(object):                                    # sharded.SlabReducer without a process group
        scatter_slabs, world, rank, reserve_sms = True, 1, 0, 0
        def chunk_slabs(self, grid): return grid.record_slabs
        def slab_range(self, grid, rank=None): return 0, grid.record_slabs
        def begin(self, acc, grid, mode): return (acc, grid)
        def finish(self, token): return token[0]
        def map_sums(self, token, mine, grid, acc):
            from pynmap_b200 import engine
            return engine.slab_sums(grid, mine, 0, grid.record_slabs, acc)
        def n_total(self, n_local): return n_local
        def reduce_max(self, *v): return v

    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
    snap.rotate(PA=30., inclination=60.)
    kw = dict(weight_name="mass", limXY=wl.LIM_XY, nXY=[48, 40], limV=wl.LIM_V, nV=37, accumulate="fixed")
    vm0, lv0 = snap.project(**kw)
    sm0 = lv0.convolve_spatial(1.0, 1.0)
    post = torch.cuda.Stream()
    for stream in (None, post, post, post):
        vm1, lv1 = snap.project(reducer=OneRank(), post_stream=stream, **kw)
        with (torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())):
            sm1 = lv1.convolve_spatial(1.0, 1.0)
        if stream is None:                                    # two-phase form: scatter now, the rest on demand
            pend = snap.project(reducer=OneRank(), deferred=True, **kw)
            assert pend.result() is pend.result()
            assert tsame(pend.result()[0].V, vm0.V) and torch.equal(pend.result()[1].losvd, lv0.losvd)
        torch.cuda.synchronize()
        assert lv1.vrange == (0, 37)
        assert tsame(vm1.M, vm0.M) and tsame(vm1.V, vm0.V) and tsame(vm1.S, vm0.S)
        assert torch.equal(lv1.losvd, lv0.losvd) and torch.equal(sm1.losvd, sm0.losvd)
