"""Development experiment: the scatter kernel on a UNIFORM box of particles (no hot pixels) and on the disk+bulge
snapshot, per accumulator layout -- separates address contention from unit throughput."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from pynmap_b200 import engine

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

n = 100_000_000
gen = torch.Generator(device="cuda").manual_seed(5)
soa = torch.empty((7, n), dtype=torch.float32, device="cuda")
for k in range(3): soa[k] = (torch.rand(n, generator=gen, device="cuda") - 0.5) * 29.0
for k in range(3, 6): soa[k] = (torch.rand(n, generator=gen, device="cuda") - 0.5) * 950.0
soa[6] = 0.5 + 1.5 * torch.rand(n, generator=gen, device="cuda")
disk = workloads.make_snapshot(n, seed=1235, device="cuda")
mat = np.eye(3, dtype=np.float32)
f64 = engine.AccMode("f64")
grid = engine.get_grid(workloads.LIM_XY, 128, 128, workloads.LIM_V, 256)
out = []
for name, data in (("uniform", soa), ("disk+bulge", disk)):
    rows = [data[k] for k in range(6)]
    for label, sums in (("lane pairs, plain cube", 31), ("lane quads, interleaved cube", 31 | 256), ("maps only, lane quads", 7),
                        ("cube only", 8)):
        maps, cube = engine.new_accumulators(grid, sums | 1, f64)
        if sums == 8:
            _, cube = engine.new_accumulators(grid, 8, f64)
        t = timeit(lambda: engine.project_bin(grid, rows, data[6], mat, 1, 1, sums, f64, None, maps, cube))
        out.append("%-12s %-30s %.3f ms" % (name, label, t))
print("\n".join(out))
