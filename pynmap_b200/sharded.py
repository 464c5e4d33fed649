"""Particle-sharded multi-GPU execution (new: the reference is single process).

One process per GPU.  The path shards naturally: every rank bins a contiguous slice of the
particles into its own full-size accumulators (no halo, no sorting) and the only exchange is a
SUM of those partial grids -- `torch.distributed.reduce` (NCCL over NVLink on the GPU box, gloo
in the CPU tests).  With int64 fixed-point accumulators the reduced grids are bit-identical for
any number of ranks, because integer addition is associative.
"""
import torch
import torch.distributed as dist

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
