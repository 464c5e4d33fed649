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

    def __init__(self, group=None, dst=0, all_ranks=False):
        if not dist.is_initialized():
            raise RuntimeError("GridReducer needs an initialised torch.distributed process group")
        self.group, self.dst, self.all_ranks = group, dst, all_ranks
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
