"""CPU, world_size 2, gloo: the multi-rank host logic -- shard boundaries, agreement on global
bounds / particle count, and the sum-reduction of partial grids -- with the oracle standing in for
the per-rank kernel.  Fixed-point partial grids must reduce to the single-process bits."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_c as oc
from oracle import oracle_np as on


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(n, seed=12):
    rng = np.random.default_rng(seed)
    soa = [rng.normal(0, 4, n).astype(np.float32) for _ in range(3)] + [rng.normal(0, 120, n).astype(np.float32) for _ in range(3)]
    return soa, rng.uniform(0.5, 2, n).astype(np.float32)


def _worker(rank, world, port, n, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pynmap_b200 import engine
    from pynmap_b200.sharded import GridReducer, shard_bounds
    soa, w = _make(n)
    lo, hi = shard_bounds(n, rank, world)
    red = GridReducer()
    n_total = red.n_total(hi - lo)
    max_w, max_v = red.reduce_max(float(np.abs(w[lo:hi]).max()), 2.0 * max(float(np.abs(a[lo:hi]).max()) for a in soa[3:]))
    scales = engine.fixed_scales(n_total, max_w, max_v)
    mat = on.rotation_matrix(np.float32(30), np.float32(60))
    ex, ev = on.bin_edges(-10, 10, 24), on.bin_edges(-300, 300, 16)
    maps, cube = oc.project_bin([a[lo:hi] for a in soa], w[lo:hi], mat, 1, 1, ex, ex, ev, 15, oc.ACC_FIXED, scales)
    tm, tc = torch.from_numpy(maps), torch.from_numpy(cube)
    holds = red(tm, tc)
    assert holds == (rank == 0)
    if rank == 0:
        np.savez(out_path, maps=tm.numpy(), cube=tc.numpy(), scales=np.array(scales), n_total=n_total)
    # all-reduce flavour: every rank ends with the same float64 total
    f = torch.from_numpy(oc.project_bin([a[lo:hi] for a in soa], w[lo:hi], mat, 1, 1, ex, ex, ev, 15)[0])
    assert GridReducer(all_ranks=True)(f)
    gathered = [torch.empty_like(f) for _ in range(world)]
    dist.all_gather(gathered, f)
    assert all(torch.equal(g, f) for g in gathered)
    dist.destroy_process_group()


def test_shard_bounds_cover_and_align():
    from pynmap_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 1000, 10 ** 9 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(lo % 8 == 0 or lo == n for lo, _ in spans)


def test_two_rank_reduce_is_bit_identical_to_one_rank(tmp_path):
    n, world = 50001, 2
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    z = np.load(out)
    soa, w = _make(n)
    mat = on.rotation_matrix(np.float32(30), np.float32(60))
    ex, ev = on.bin_edges(-10, 10, 24), on.bin_edges(-300, 300, 16)
    assert int(z["n_total"]) == n
    maps, cube = oc.project_bin(soa, w, mat, 1, 1, ex, ex, ev, 15, oc.ACC_FIXED, z["scales"])
    assert np.array_equal(maps, z["maps"]) and np.array_equal(cube, z["cube"])


# ---------------------------------------------------------------------------------------------- slab reducer
class _FakeGrid(object):
    """The attributes sharded.SlabReducer reads from engine.Grid (no device needed)."""
    def __init__(self, nx, ny, nv):
        self.nx, self.ny, self.nv = nx, ny, nv
        self.record_slabs = (nv + 1) // 2 + 1
        self.slab_elems = nx * ny * 4


def _interleave(maps_rest, s1, s2, cube, nslab_pad):
    """numpy restatement of the PNM_CUBE_MOMENTS layout (include/pnm_b200.h): records [g][q][4] =
    { w of bin 2g, w of bin 2g+1, sum w*d, sum w*d*d }; the moments and the out-of-range weights sit in slab G."""
    nv, npix = cube.shape
    G = (nv + 1) // 2
    acc = np.zeros((nslab_pad, npix, 4), dtype=cube.dtype)
    for iv in range(nv):
        acc[iv // 2, :, iv & 1] = cube[iv]
    acc[G, :, 0], acc[G, :, 2], acc[G, :, 3] = maps_rest, s1, s2
    return acc


def _slab_worker(rank, world, port, nv, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pynmap_b200.sharded import SlabReducer
    nx, ny = 6, 5
    npix = nx * ny
    grid = _FakeGrid(nx, ny, nv)
    red = SlabReducer()
    c = red.chunk_slabs(grid)
    assert c * world >= grid.record_slabs and red.slab_range(grid) == (rank * c, (rank + 1) * c)
    rng = np.random.default_rng(100 + rank)
    cube = rng.integers(0, 1000, (nv, npix)).astype(np.int64)
    rest, s1, s2 = (rng.integers(-500, 500, npix).astype(np.int64) for _ in range(3))
    mine = _interleave(rest, s1, s2, cube, c * world)
    total = torch.from_numpy(mine.copy())
    dist.all_reduce(total)                                   # what the ranks together hold
    chunk = red.scatter(torch.from_numpy(mine.reshape(-1).copy()), grid)
    g0, g1 = red.slab_range(grid)
    assert torch.equal(chunk.reshape(c, npix, 4), total[g0:g1])
    # per-rank share of the map sums, all-reduced == the sums over every slab of the total
    part = chunk.reshape(c, npix, 4)
    share = torch.stack([part[:, :, 0].sum(0) + part[:, :, 1].sum(0), part[:, :, 2].sum(0), part[:, :, 3].sum(0)])
    sums3 = red.all_sum(share.clone())
    want = torch.stack([total[:, :, 0].sum(0) + total[:, :, 1].sum(0), total[:, :, 2].sum(0), total[:, :, 3].sum(0)])
    assert torch.equal(sums3, want)
    # own velocity bins -> the whole cube on rank 0 only / on every rank
    iv0, iv1 = min(2 * g0, nv), min(2 * g1, nv)
    local = torch.empty((nx, ny, iv1 - iv0), dtype=torch.float64)
    for iv in range(iv0, iv1):
        local[:, :, iv - iv0] = part[iv // 2 - g0, :, iv & 1].reshape(nx, ny).double()
    whole = red.gather_losvd(local, grid, dst=0)
    everywhere = red.gather_losvd(local, grid)
    ref = torch.stack([total[iv // 2, :, iv & 1].reshape(nx, ny).double() for iv in range(nv)], dim=2)
    assert (whole is None) == (rank != 0) and torch.equal(everywhere, ref)
    if rank == 0:
        assert torch.equal(whole, ref)
        np.savez(out_path, ok=np.array(1))
    dist.destroy_process_group()


def test_slab_reducer_two_and_three_ranks(tmp_path):
    """sharded.SlabReducer on gloo: chunking of the padded record slabs, reduce-scatter (all-reduce + slice on gloo),
    the all-reduced map sums and the gather of the per-rank velocity bins -- odd and even velocity axes, more ranks than
    some chunk boundaries like."""
    for world, nv in ((2, 8), (3, 7), (2, 1)):
        out = str(tmp_path / ("slab_%d_%d.npz" % (world, nv)))
        mp.spawn(_slab_worker, args=(world, _free_port(), nv, out), nprocs=world, join=True)
        assert os.path.isfile(out)
