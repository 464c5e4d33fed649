"""Development experiment: one workload (bench config), knobs from the environment."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from pynmap_b200 import engine, rotation

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
soa = workloads.make_snapshot(n, seed=1235, device="cuda")
srt = workloads.morton_sort(soa) if "--sorted" in sys.argv else None
mat = rotation.rotation_matrix(np.float32(30), np.float32(60))
f64 = engine.AccMode("f64")
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("PNM_") and k != "PNM_B200_LIB")
lib = os.path.basename(os.environ.get("PNM_B200_LIB", "default"))
out = []
for name, data, nxy, nv, sums, limV in (("none", soa, 128, 256, 8, [1e6, 1e6 + 1]), ("cube", soa, 128, 256, 8, workloads.LIM_V),
                                        ("maps3", soa, 128, 0, 7, workloads.LIM_V), ("full31", soa, 128, 256, 31, workloads.LIM_V),
                                        ("maps3_256", soa, 256, 0, 7, workloads.LIM_V), ("sorted31", srt, 128, 256, 31, workloads.LIM_V),
                                        ("moments287", soa, 128, 256, 287, workloads.LIM_V), ("moments287_256x512", soa, 256, 512, 287, workloads.LIM_V)):
    if data is None: continue
    grid = engine.get_grid(workloads.LIM_XY, nxy, nxy, limV, nv)
    maps, cube = engine.new_accumulators(grid, sums | 1, f64)
    if sums & 8 and not sums & 256: _, cube = engine.new_accumulators(grid, 8, f64)
    rows = [data[k] for k in range(6)]
    t = timeit(lambda: engine.project_bin(grid, rows, data[6], mat, 1, 1, sums, f64, None, maps, cube))
    out.append("%s=%.3f" % (name, t))
print("%-22s %-50s R=%d  %s" % (lib, tag, grid.map_replicas, "  ".join(out)), flush=True)
