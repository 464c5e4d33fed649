"""Diagnostic on ONE GPU: cost of the sharded path's extra pieces (replica fold, owner / non-owner
post-processing) when they overlap the next step's scatter.  No process group: a dummy reducer."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
import pynmap_b200 as pnm
from pynmap_b200 import engine


class Dummy(object):
    def __init__(self, keep):
        self.keep, self.dst = keep, 0
    def __call__(self, *tensors):
        k = self.keep if not callable(self.keep) else self.keep()
        return k


n = 100_000_000
soa = workloads.make_snapshot(n, bulge_fraction=0.2, seed=1234, device="cuda")
snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
snap.rotate(PA=30., inclination=60.)
post = torch.cuda.Stream(priority=-1)
sink = io.StringIO()
flip = [0]
def alternate():
    flip[0] ^= 1
    return bool(flip[0])
variants = {"no reducer": None, "fold, owner every step": Dummy(True), "fold, never owner": Dummy(False),
            "fold, owner every other step": Dummy(alternate)}
if len(sys.argv) > 1:
    variants = {k: v for k, v in variants.items() if sys.argv[1] in k}
for name, red in variants.items():
    def step():
        with contextlib.redirect_stdout(sink):
            vmap, losvd = snap.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=128, limV=workloads.LIM_V, nV=256,
                                       reducer=red, post_stream=post)
            if losvd is not None:
                with torch.cuda.stream(post):
                    losvd.convolve_spatial(1.0, 1.0)
        sink.seek(0); sink.truncate()
    for _ in range(6):
        step()
    torch.cuda.synchronize()
    engine.KERNEL_TIMER = timer = engine.KernelTimer()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        step()
    torch.cuda.current_stream().wait_stream(post)
    b.record(); torch.cuda.synchronize()
    engine.KERNEL_TIMER = None
    print("%-32s ms/step %.4f   scatter kernel %.4f" % (name, a.elapsed_time(b) / 40, timer.total_ms() / max(1, len(timer.pairs))))
