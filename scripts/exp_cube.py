"""Development experiment (round 2): k_project_cube (shared-memory privatised scatter) against k_project_bin on the
bench workload -- timing, plan statistics and a bit-exact comparison of the int64 accumulators.
usage: python scripts/exp_cube.py [n_particles]"""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
from pynmap_b200 import engine, rotation, _lib

def timeit(fn, reps=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nxy, nv = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 256)
soa = workloads.make_snapshot(n, seed=1235, device="cuda")
rows = [soa[k] for k in range(6)]; w = soa[6]
mat = rotation.rotation_matrix(np.float32(30), np.float32(60))
grid = engine.get_grid(workloads.LIM_XY, nxy, nxy, workloads.LIM_V, nv)
f64, fx = engine.AccMode("f64"), engine.AccMode("fixed")
max_w, max_d = engine.absmax(w), max(engine.absmax(r) for r in rows[3:6]) * 3 ** 0.5
plan = engine.CubePlan(grid)
t_plan = timeit(lambda: plan.build(rows, w, mat))
info = plan.info()
scales = engine.cube_fixed_scales(n, max_w, max_d, info["mean_abs"])
print("scales fixed (capped)", [math.log2(v) for v in scales], " uncapped", [math.log2(v) for v in engine.fixed_scales(n, max_w, max_d)],
      " smem (f64)", [math.log2(v) for v in info["smem_scales"]], flush=True)
print("n=%d grid %dx%dx%d  plan build %.3f ms  %s  captured %.3f of the sample" % (n, nxy, nxy, nv, t_plan[0], info, info["captured"] / max(info["sampled"], 1)), flush=True)
sums = 31 | _lib.CUBE_MOMENTS
res = {}
for acc, sc, name in ((fx, scales, "fixed"), (f64, None, "f64")):
    ref = torch.zeros(grid.cube_moments_elems, dtype=acc.dtype, device="cuda")
    engine.project_bin(grid, rows, w, mat, 1, 1, sums, acc, sc, None, ref)
    new = torch.zeros_like(ref)
    engine.project_cube(grid, plan, rows, w, mat, acc, sc, new)
    torch.cuda.synchronize()
    # finalised maps / cube agree?
    Mr, Vr, Sr = engine.finalize_vmaps(grid, None, ref, sums, acc, sc)
    Mn, Vn, Sn = engine.finalize_vmaps(grid, None, new, sums, acc, sc)
    cr, cn = engine.finalize_cube(grid, ref, acc, sc, sums), engine.finalize_cube(grid, new, acc, sc, sums)
    rel = lambda a, b: float(((a - b).abs() / b.abs().clamp_min(1e-300)).max())
    ok = torch.isfinite(Sr) & torch.isfinite(Sn)
    print("%s: finalised bit-identical M %s V %s S %s cube %s | max rel diff  M %.2e  cube %.2e  V(abs) %.2e  S(abs) %.2e" % (
        name, bool(torch.equal(Mn, Mr)), bool(torch.equal(Vn.nan_to_num(), Vr.nan_to_num())), bool(torch.equal(Sn.nan_to_num(), Sr.nan_to_num())),
        bool(torch.equal(cn, cr)), rel(Mn, Mr), rel(cn, cr), float((Vn - Vr).nan_to_num().abs().max()), float((Sn - Sr)[ok].abs().max())), flush=True)
    buf = torch.zeros_like(ref)
    t_old = timeit(lambda: engine.project_bin(grid, rows, w, mat, 1, 1, sums, acc, sc, None, buf))
    t_new = timeit(lambda: engine.project_cube(grid, plan, rows, w, mat, acc, sc, buf))
    print("%s: k_project_bin %.3f ms (med %.3f)   k_project_cube %.3f ms (med %.3f)   roofline frac %.3f" % (name, t_old[0], t_old[1], t_new[0], t_new[1], 28 * n / t_new[0] / 1e6 / 6448.7), flush=True)
for dbg, what in ((1, "no shared-memory adds"), (2, "no in-box REDs"), (4, "no queue REDs"), (7, "loads + arithmetic only"), (3, "queue REDs only"), (5, "in-box REDs only"), (6, "shared-memory adds only")):
    os.environ["PNM_CUBE_DBG"] = str(dbg)
    buf = torch.zeros(grid.cube_moments_elems, dtype=torch.float64, device="cuda")
    t = timeit(lambda: engine.project_cube(grid, plan, rows, w, mat, f64, None, buf))
    print("dbg %d (%s): %.3f ms" % (dbg, what, t[0]), flush=True)
os.environ["PNM_CUBE_DBG"] = "0"
