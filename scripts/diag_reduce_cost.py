"""Diagnostic (not a bench value): what the multi-GPU step costs beyond the scatter.  Runs bench.py's
step loop under torchrun with the exchange replaced piece by piece:
    PNM_DIAG=none       the normal step
    PNM_DIAG=noreduce   GridReducer returns without calling NCCL (results are wrong on purpose)
    PNM_DIAG=fixeddst   normal reduce, owner fixed to rank 0
Usage: PNM_DIAG=noreduce python -m torch.distributed.run --nproc-per-node 2 ... scripts/diag_reduce_cost.py --gpus 2 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pynmap_b200 import sharded  # noqa: E402

mode = os.environ.get("PNM_DIAG", "none")
if mode == "noreduce":
    def skip(self, *tensors):
        return self.all_ranks or self.rank == self.dst
    sharded.GridReducer.__call__ = skip
elif mode == "fixeddst":
    sharded.GridReducer.dst = property(lambda self: 0, lambda self, v: None)
sys.exit(bench.main())
