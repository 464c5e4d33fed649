"""Development (round 2): CUDA-event timeline of the pipelined sharded step -- when do the scatter kernel, the reduce
over ranks and the owner's post-processing of consecutive steps run relative to each other?
usage: torchrun --nproc-per-node N scripts/timeline_n.py [--nccl-max-ctas C] [--reserve-sms R] [--steps K]"""
import argparse, contextlib, io, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--nccl-max-ctas", type=int, default=4)
ap.add_argument("--reserve-sms", type=int, default=0)
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--reduce", default="scatter", choices=["scatter", "owner", "peer"])
ap.add_argument("--particles", type=int, default=100_000_000)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
if args.nccl_max_ctas > 0:
    opts.config.max_ctas = args.nccl_max_ctas
    opts.config.min_ctas = 1
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
import pynmap_b200 as pnm, workloads
from pynmap_b200 import engine
from pynmap_b200.sharded import GridReducer, PeerSlabReducer, SlabReducer
soa = workloads.make_snapshot(args.particles, bulge_fraction=0.2, seed=1235 + 1000 * rank, device="cuda")
snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
snap.rotate(PA=30., inclination=60.)
reducer = (GridReducer if args.reduce == 'owner' else SlabReducer)(reserve_sms=args.reserve_sms) if args.reduce != 'peer' else None
post = torch.cuda.Stream(priority=-1)
sink = io.StringIO()
marks = []          # (step, name, event)
class Timed(object):
    def __init__(self, inner): self.inner = inner
    def __getattr__(self, k): return getattr(self.inner, k)
    def __call__(self, *t):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = self.inner(*t); b.record()
        marks.append((counter[0], "reduce0", a)); marks.append((counter[0], "reduce1", b))
        return out
    def scatter(self, *t):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = self.inner.scatter(*t); b.record()
        marks.append((counter[0], "reduce0", a)); marks.append((counter[0], "reduce1", b))
        return out
timed = Timed(reducer)
counter = [0]
def step():
    reducer.dst = counter[0] % world
    with contextlib.redirect_stdout(sink):
        vm, lv = snap.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=128, limV=workloads.LIM_V, nV=256, reducer=timed, post_stream=post)
        if lv is not None:
            with torch.cuda.stream(post):
                lv.convolve_spatial(1.0, 1.0)
    e = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(post): e.record()
    marks.append((counter[0], "post1", e))
    counter[0] += 1
    sink.seek(0); sink.truncate()
if args.reduce == "peer":
    # deferred pipeline on one stream: trace events come from the reducer itself
    reducer = PeerSlabReducer()
    pend = [None]
    def finish(p):
        with contextlib.redirect_stdout(sink):
            vm, lv = p.result(); lv.convolve_spatial(1.0, 1.0)
        sink.seek(0); sink.truncate()
        reducer._mark("post_end", torch.cuda.current_stream())
    def step():
        with contextlib.redirect_stdout(sink):
            cur = snap.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=128, limV=workloads.LIM_V, nV=256, reducer=reducer, deferred=True)
        if pend[0] is not None: finish(pend[0])
        pend[0] = cur
    for _ in range(8): step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    reducer.trace = []
    engine.KERNEL_TIMER = timer = engine.KernelTimer()
    origin = torch.cuda.Event(enable_timing=True); origin.record()
    for _ in range(args.steps): step()
    torch.cuda.synchronize()
    engine.KERNEL_TIMER = None
    ev = [("scat0", a) for a, b in timer.pairs] + [("scat1", b) for a, b in timer.pairs] + reducer.trace
    rows = sorted((origin.elapsed_time(e), name) for name, e in ev)
    for r in range(world):
        dist.barrier()
        if r == rank and rank in (0, world - 1):
            print("rank %d (peer, deferred): events in time order [ms from origin]" % rank)
            line = []
            for t, name in rows:
                if name == "scat0" and line:
                    print("   " + "  ".join(line)); line = []
                line.append("%s %.3f" % (name, t))
            print("   " + "  ".join(line), flush=True)
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0)
for _ in range(2 * world + 4): step()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
marks.clear()
engine.KERNEL_TIMER = timer = engine.KernelTimer()
origin = torch.cuda.Event(enable_timing=True); origin.record()
first = counter[0]
for _ in range(args.steps): step()
torch.cuda.synchronize()
engine.KERNEL_TIMER = None
rows = {}
for k, (a, b) in enumerate(timer.pairs):
    rows.setdefault(first + k, {})["scat0"] = origin.elapsed_time(a); rows[first + k]["scat1"] = origin.elapsed_time(b)
for s, name, e in marks:
    rows.setdefault(s, {})[name] = origin.elapsed_time(e)
for r in range(world):
    dist.barrier()
    if r == rank and rank in (0, 1):
        print("rank %d (%s, ctas %d, reserved %d): step owner | scatter start..end | reduce start..end | post end   [ms from origin]" % (rank, args.reduce, args.nccl_max_ctas, args.reserve_sms))
        for s in sorted(rows):
            v = rows[s]
            print("  %3d  own=%d | %7.3f .. %7.3f | %7.3f .. %7.3f | %7.3f" % (s, s % world, v.get("scat0", -1), v.get("scat1", -1), v.get("reduce0", -1), v.get("reduce1", -1), v.get("post1", -1)), flush=True)
dist.barrier()
dist.destroy_process_group()
