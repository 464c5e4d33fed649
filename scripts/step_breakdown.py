"""Development: CUDA-event breakdown of one bench step (zero, bin, finalise, cube_out, smooth) + host time."""
import contextlib, io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
import pynmap_b200 as pnm
from pynmap_b200 import engine, _lib, rotation, hostpath

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
soa = workloads.make_snapshot(n, seed=1235, device="cuda")
rows = [soa[k] for k in range(6)]
mat = rotation.rotation_matrix(np.float32(30), np.float32(60))
acc = engine.AccMode("f64")
grid = engine.get_grid(workloads.LIM_XY, 128, 128, workloads.LIM_V, 256)
sums = 31
t0, t1 = hostpath.smoothing_taps(1.0, 1.0, 30 / 127, 30 / 127)
ev = lambda: torch.cuda.Event(enable_timing=True)
stages = ["zero", "bin", "finalize_vmaps", "cube_out", "smooth"]
tot = {s: 0.0 for s in stages}
reps = 10
for it in range(reps + 3):
    e = [ev() for _ in range(len(stages) + 1)]
    e[0].record()
    maps, cube = engine.new_accumulators(grid, sums, acc); e[1].record()
    engine.project_bin(grid, rows, soa[6], mat, 1, 1, sums, acc, None, maps, cube); e[2].record()
    M, V, S = engine.finalize_vmaps(grid, maps, cube, sums, acc, None); e[3].record()
    c = engine.finalize_cube(grid, cube, acc, None); e[4].record()
    sm = engine.smooth_cube(c, t0, t1); e[5].record()
    torch.cuda.synchronize()
    if it >= 3:
        for k, s in enumerate(stages):
            tot[s] += e[k].elapsed_time(e[k + 1])
print("  ".join("%s=%.3f" % (s, tot[s] / reps) for s in stages), " sum=%.3f ms" % (sum(tot.values()) / reps))
# host-side cost of the public API step (no GPU wait)
snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
snap.rotate(PA=30., inclination=60.)
sink = io.StringIO()
def step():
    with contextlib.redirect_stdout(sink):
        vmap, losvd = snap.project(weight_name="mass", limXY=workloads.LIM_XY, nXY=128, limV=workloads.LIM_V, nV=256)
        return losvd.convolve_spatial(1.0, 1.0)
for _ in range(3): step()
torch.cuda.synchronize()
a, b = ev(), ev()
h0 = time.perf_counter(); a.record()
for _ in range(20): step()
h1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
print("public API: host enqueue %.3f ms/step, gpu %.3f ms/step" % ((h1 - h0) * 50, a.elapsed_time(b) / 20))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
