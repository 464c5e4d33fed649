"""Particle-sharded multi-GPU execution (new: the reference is single process).

One process per GPU.  The path shards naturally: every rank bins a contiguous slice of the particles into its own
full-size accumulators (no halo, no sorting) and the only exchange is a SUM of those partial grids.  Three flavours:

  GridReducer       reduce / all-reduce of the whole partial grids (NCCL on the GPU box, gloo in the CPU tests): one rank
                    (or every rank) finalises everything.  velocity_moments(reducer=) and the host path use it.
  SlabReducer       reduce-scatter along the record slabs (pairs of velocity bins) of the interleaved cube: every rank
                    finalises and smooths its own velocity bins, the maps come from an all-reduce of per-rank shares.
  PeerSlabReducer   the same exchange without NCCL kernels or SMs: accumulators in symmetric memory, copy-engine pulls over
                    NVLink that run under the next scatter kernel (snapshot.project(deferred=True)), local sum in rank order.

With int64 fixed-point accumulators the reduced grids are bit-identical for any number of ranks and any flavour, because
integer addition is associative (bench.py checks it on hardware every run).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib

ALIGN = 8   # particles: keeps every shard's base pointer 32-byte aligned for float32 arrays


def shard_bounds(n, rank, world, align=ALIGN):
    """Contiguous [lo, hi) slice of n particles for `rank`; interior boundaries are multiples of
    `align` so that each shard keeps the 16-byte alignment the vector-load kernel wants."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


class GridReducer(object):
    """Sums partial accumulators over the process group.  Passed as ``reducer=`` to
    ``snapshot.project``; also usable on bare tensors.

    all_ranks=False: reduce to ``dst`` (only that rank finalises); True: all-reduce."""

    def __init__(self, group=None, dst=0, all_ranks=False, reserve_sms=0):
        if not dist.is_initialized():
            raise RuntimeError("GridReducer needs an initialised torch.distributed process group")
        self.group, self.dst, self.all_ranks = group, dst, all_ranks
        self.reserve_sms = int(reserve_sms)      # SMs the persistent scatter kernel leaves to the collective (pipelined runs)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def n_total(self, n_local):
        t = torch.tensor([int(n_local)], dtype=torch.int64, device=self._device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def reduce_max(self, *values):
        """Global bounds for the fixed-point scales: every rank must quantise identically."""
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=self._device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return tuple(float(v) for v in t.tolist())

    def _device(self):
        return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")

    def __call__(self, *tensors):
        """In-place sum of each tensor over the ranks.  Returns True on ranks that hold the total."""
        for t in tensors:
            if t is None:
                continue
            if self.all_ranks:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.reduce(t, dst=self.dst, op=dist.ReduceOp.SUM, group=self.group)
        return self.all_ranks or self.rank == self.dst


class SlabReducer(GridReducer):
    """Reduce-scatter flavour for ``snapshot.project``: the interleaved cube accumulator is an array of record slabs
    (pairs of velocity bins x all pixels), padded to a multiple of the world size; a reduce-scatter leaves every rank
    with the TOTAL of its own contiguous slab range.  Each rank then finalises and smooths its own velocity bins
    (smoothing acts per velocity slice), and the moment maps come from a small all-reduce of per-rank partial sums
    (3 values per pixel).  Compared with a reduce to one owner the post-processing is spread evenly over the ranks
    and nothing has to rotate.  ``project`` returns, on every rank, the full SnapMap and the LOSVD of the rank's
    velocity bins (``gather_losvd`` assembles the whole cube where it is wanted)."""

    scatter_slabs = True

    def chunk_slabs(self, grid):
        """Record slabs per rank."""
        return -(-grid.record_slabs // self.world)

    def slab_range(self, grid, rank=None):
        c = self.chunk_slabs(grid)
        r = self.rank if rank is None else rank
        return r * c, (r + 1) * c

    def scatter(self, acc_padded, grid):
        """acc_padded: world * chunk_slabs record slabs (this rank's partial sums) -> this rank's slabs of the total."""
        per = self.chunk_slabs(grid) * grid.slab_elems
        if acc_padded.numel() != per * self.world:
            raise ValueError("accumulator of %d elements, expected %d (padded to the world size)" % (acc_padded.numel(), per * self.world))
        if dist.get_backend(self.group) == "gloo":          # gloo has no reduce-scatter: sum everything, keep the own chunk
            dist.all_reduce(acc_padded, op=dist.ReduceOp.SUM, group=self.group)
            return acc_padded[self.rank * per:(self.rank + 1) * per]
        out = torch.empty(per, dtype=acc_padded.dtype, device=acc_padded.device)
        dist.reduce_scatter_tensor(out, acc_padded, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def all_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    # two-phase form (snapshot.project(deferred=True)): ``begin`` right after the scatter kernel has been launched,
    # ``finish`` when the rank's slabs of the total are wanted.  Here the collective simply runs in ``finish``.
    def begin(self, acc_padded, grid, acc):
        return (acc_padded, grid)

    def finish(self, token):
        return self.scatter(*token)

    def map_sums(self, token, mine, grid, acc):
        """Total (sum w, sum w*d, sum w*d*d) per pixel, (3, nx*ny) of the accumulator type, on every rank: the rank's
        share from its own slabs of the total, all-reduced."""
        from . import engine
        g0, g1 = self.slab_range(grid)
        return self.all_sum(engine.slab_sums(grid, mine, g0, g1, acc))

    def gather_losvd(self, local_cube, grid, dst=None):
        """(nX, nY, own bins) float64 cubes of all ranks -> the whole (nX, nY, nV) cube on rank ``dst`` (None: on every
        rank); other ranks get None."""
        c = self.chunk_slabs(grid)
        pad = torch.zeros((grid.nx, grid.ny, 2 * c), dtype=local_cube.dtype, device=local_cube.device)
        pad[:, :, :local_cube.shape[2]] = local_cube
        parts = torch.empty(self.world * pad.numel(), dtype=pad.dtype, device=pad.device)
        dist.all_gather_into_tensor(parts, pad.reshape(-1), group=self.group)
        if dst is not None and dst != self.rank:
            return None
        parts = parts.reshape((self.world,) + tuple(pad.shape))
        return parts.permute(1, 2, 0, 3).reshape(grid.nx, grid.ny, self.world * 2 * c)[:, :, :grid.nv].contiguous()


class PeerSlabReducer(SlabReducer):
    """SlabReducer whose reduce-scatter uses no SMs and no NCCL kernel: the accumulators live in symmetric memory
    (torch.distributed._symmetric_memory: every rank can address every rank's buffer over NVLink), and after a
    device-side barrier every rank PULLS its own chunk of every rank's partial grid with the copy engines
    (cudaMemcpyAsync peer copies on a side stream) into a local staging area; ``finish`` sums the staged chunks in rank
    order (pnm_sum_chunks -- deterministic, and exact for int64).  The scatter kernel is persistent with one CTA per SM,
    so nothing that needs SMs can run next to it (profiles/r02_timeline_*.log): the copy engines can.  With
    ``snapshot.project(deferred=True)`` the pulls of step k overlap the scatter kernel of step k + 1:

        pending = snap.project(..., reducer=r, deferred=True)      # scatter k, barrier, pulls start
        nxt = snap.project(..., reducer=r, deferred=True)          # scatter k + 1 runs while step k's chunks arrive
        vmap, losvd = pending.result()                             # sum of the staged chunks, finalisation, ...

    Buffers: ``nbuf`` (>= 2) accumulators used round-robin + two staging areas, per (size, dtype).  An accumulator is
    cleared after the barrier of the FOLLOWING step, when every rank has finished pulling from it."""

    def __init__(self, group=None, nbuf=2, barrier_timeout_ms=60000):
        super().__init__(group)
        import torch.distributed._symmetric_memory as symm
        if dist.get_backend(self.group) != "nccl":
            raise RuntimeError("PeerSlabReducer needs CUDA devices with peer access (nccl process group)")
        self._symm = symm
        self._group_name = (self.group if self.group is not None else dist.group.WORLD).group_name
        self.nbuf = max(2, int(nbuf))
        self.timeout_ms = int(barrier_timeout_ms)
        self._rings = {}
        self._copy_stream = torch.cuda.Stream()
        self.peer_sums = os.environ.get("PNM_PEER_SUMS", "nccl") == "peer"    # map sums over peer memory or an NCCL all-reduce
        self.trace = None                    # development: a list collects (name, timing event) pairs (scripts/timeline_n.py)

    def _mark(self, name, stream):
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            self.trace.append((name, e))

    def _ring(self, numel, dtype, device):
        key = (int(numel), dtype)
        ring = self._rings.get(key)
        if ring is None:                     # collective: every rank creates its rings in the same order
            bufs, handles = [], []
            for _ in range(self.nbuf):
                t = self._symm.empty(int(numel), dtype=dtype, device=device)
                handles.append(self._symm.rendezvous(t, self._group_name))
                t.zero_()
                bufs.append(t)
            stage = [torch.empty(int(numel), dtype=dtype, device=device) for _ in range(2)]
            ptrs = [(ctypes.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs]) for h in handles]
            ring = self._rings[key] = dict(bufs=bufs, handles=handles, stage=stage, next=0, step=0, pending=None, pulled=None,
                                           ptrs=ptrs, index={b.data_ptr(): i for i, b in enumerate(bufs)})
        return ring

    def acquire(self, grid, acc):
        """A zeroed, padded accumulator in symmetric memory for the next scatter (round-robin)."""
        ring = self._ring(self.chunk_slabs(grid) * self.world * grid.slab_elems, acc.dtype, grid.device)
        i = ring["next"]
        ring["next"] = (i + 1) % self.nbuf
        if ring["pending"] is not None and ring["bufs"][i].data_ptr() == ring["pending"].data_ptr():
            raise RuntimeError("PeerSlabReducer: accumulator still in use (call result() of the pending projection "
                               "before starting more than %d others)" % (self.nbuf - 1))
        return ring["bufs"][i]

    def begin(self, acc_padded, grid, acc):
        ring = self._ring(acc_padded.numel(), acc_padded.dtype, acc_padded.device)
        idx = ring["index"][acc_padded.data_ptr()]
        per = self.chunk_slabs(grid) * grid.slab_elems
        cur = torch.cuda.current_stream()
        self._mark("begin", cur)
        if ring["pulled"] is not None:
            cur.wait_event(ring["pulled"])               # my pulls of the previous step are done before I say so
        ring["handles"][0].barrier(channel=0, timeout_ms=self.timeout_ms)    # every rank: scatter done, previous pulls done
        self._mark("barrier", cur)
        if ring["pending"] is not None:
            ring["pending"].zero_()                      # nobody reads the previous accumulator any more
        self._mark("zeroed", cur)
        ready = torch.cuda.Event()
        ready.record(cur)
        si = ring["step"] % 2
        if ring.setdefault("busy", [False, False])[si]:
            raise RuntimeError("PeerSlabReducer: more than two projections in flight -- call result() of the older "
                               "pending projections first (their staged chunks would be overwritten)")
        ring["busy"][si] = True
        stage = ring["stage"][si]
        ring["step"] += 1
        cs = self._copy_stream
        cs_ptr = ctypes.c_void_p(cs.cuda_stream)
        cs.wait_event(ready)
        ring["pending"] = acc_padded
        # copy-engine pulls through the C ABI: one call issues the cudaMemcpyAsync of every rank's chunk on the copy stream
        esz = acc_padded.element_size()
        bases = ring["ptrs"][idx]
        _lib.call("pnm_pull_chunks", ctypes.c_void_p(stage.data_ptr()), bases, self.world, self.rank, per * esz, cs_ptr)
        pulled = torch.cuda.Event()
        pulled.record(cs)
        self._mark("pulled", cs)
        ring["pulled"] = pulled
        return (stage, pulled, per, acc, ring, si)

    def finish(self, token):
        from . import engine
        stage, pulled, per, acc = token[:4]
        cur = torch.cuda.current_stream()
        cur.wait_event(pulled)
        self._mark("finish", cur)
        out = engine.sum_chunks(stage, self.world, per, acc)
        token[4]["busy"][token[5]] = False               # (stream order protects the staged chunks from here on)
        self._mark("summed", cur)
        return out

    def map_sums(self, token, mine, grid, acc):
        """One-shot all-reduce over peer memory instead of an NCCL call: every rank writes its share of the per-pixel
        sums into its symmetric buffer, a device-side barrier, then every rank adds the shares of all ranks in rank
        order straight from peer memory (pnm_sum_peers: 8 x 0.4 MB of NVLink loads) -- the same bits on every rank."""
        from . import engine
        if not self.peer_sums:
            return SlabReducer.map_sums(self, token, mine, grid, acc)
        g0, g1 = self.slab_range(grid)
        npix3 = 3 * grid.nx * grid.ny
        key = ("shares", npix3, acc.dtype)
        sh = self._rings.get(key)
        if sh is None:
            bufs, handles = [], []
            for _ in range(2):
                t = self._symm.empty(npix3, dtype=acc.dtype, device=grid.device)
                handles.append(self._symm.rendezvous(t, self._group_name))
                bufs.append(t)
            sh = self._rings[key] = dict(bufs=bufs, handles=handles, step=0)
        i = sh["step"] % 2               # double buffered: a rank may be one barrier ahead of the readers of the other buffer
        sh["step"] += 1
        engine.slab_sums(grid, mine, g0, g1, acc, out=sh["bufs"][i].view(3, -1))
        sh["handles"][0].barrier(channel=1, timeout_ms=self.timeout_ms)
        return engine.sum_peers([int(p) for p in sh["handles"][i].buffer_ptrs], npix3, acc, grid.device).view(3, -1)

    def scatter(self, acc_padded, grid):
        raise RuntimeError("PeerSlabReducer works through begin() / finish() on accumulators from acquire()")
