"""Development: BASELINE config 4 (64 viewing angles over one snapshot, 256 x 256 maps) -- how many views per launch?
usage: python scripts/views_bench.py [n_particles]"""
import contextlib, io, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
import pynmap_b200 as pnm
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
soa = workloads.make_snapshot(n, seed=1238, device="cuda")
snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
incs = np.repeat(np.linspace(0, 87.5, 8), 8); pas = np.tile(np.linspace(0, 157.5, 8), 8)
sink = io.StringIO()
def run(per, acc="f64"):
    with contextlib.redirect_stdout(sink):
        return snap.project_views(pas, incs, weight_name="mass", limXY=workloads.LIM_XY, nXY=256, views_per_launch=per, accumulate=acc)
ref = run(1, "fixed")
for per in (1, 2, 4, 8, 16, 32, 64):
    out = run(per, "fixed")
    same = all(torch.equal(a.M, b.M) and torch.equal(a.V.nan_to_num(), b.V.nan_to_num()) for a, b in zip(out, ref))
    for _ in range(2): run(per)
    torch.cuda.synchronize(); ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(per); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("views per launch %2d: %7.3f ms for 64 views of %.0e particles = %.1f G particle-views/s  (fixed-point maps identical to per-view launches: %s)"
          % (per, min(ts), n, 64 * n / min(ts) / 1e6, same), flush=True)
