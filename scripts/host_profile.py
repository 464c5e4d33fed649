"""Development: cProfile of the host side of the sharded (peer, deferred) bench step -- where do the ~0.6 ms of Python /
launch time per step go?  usage: torchrun --nproc-per-node N scripts/host_profile.py"""
import contextlib, cProfile, io, os, pstats, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import pynmap_b200 as pnm, workloads
from pynmap_b200.sharded import PeerSlabReducer
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
soa = workloads.make_snapshot(n, bulge_fraction=0.2, seed=1235 + 1000 * rank, device="cuda")
snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
snap.rotate(PA=30., inclination=60.)
reducer = PeerSlabReducer() if world > 1 else None
sink = io.StringIO()
pending = [None]
def finish(p):
    with contextlib.redirect_stdout(sink):
        vm, lv = p.result(); lv.convolve_spatial(1.0, 1.0)
    sink.seek(0); sink.truncate()
def step():
    with contextlib.redirect_stdout(sink):
        cur = snap.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=128, limV=workloads.LIM_V, nV=256, reducer=reducer, deferred=True)
    if pending[0] is not None: finish(pending[0])
    pending[0] = cur
for _ in range(10): step()
torch.cuda.synchronize()
prof = cProfile.Profile()
t0 = time.perf_counter()
prof.enable()
for _ in range(200): step()
prof.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
if rank == 0:
    print("host time per step %.3f ms (profiler on)" % ((t1 - t0) * 1e3 / 200))
    out = io.StringIO()
    pstats.Stats(prof, stream=out).sort_stats("tottime").print_stats(45)
    print(out.getvalue()[:9000])
if world > 1:
    dist.barrier(); dist.destroy_process_group()
