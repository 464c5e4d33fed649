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
