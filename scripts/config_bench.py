"""Development: device-resident timings of the BASELINE.json configurations (CUDA events, best of 5).
Not the bench line -- numbers for DESIGN.md."""
import contextlib, io, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads
import pynmap_b200 as pnm

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

def report(name, n, ms, bytes_per=28, extra=""):
    print("%-58s %9.3f ms  %8.2f Gpart/s  %6.1f GB/s (%4.1f%% of 6449) %s" % (
        name, ms, n / ms / 1e6, bytes_per * n / ms / 1e6, 100 * bytes_per * n / ms / 1e6 / 6448.7, extra), flush=True)

sink = io.StringIO()
with contextlib.redirect_stdout(sink):
    pass
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
soa = workloads.make_snapshot(N, seed=1235, device="cuda")
snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], output="torch")
snap.rotate(30., 60.)
q = lambda f: (lambda: contextlib.redirect_stdout(sink).__enter__() and None or f())
def quiet(f):
    def g():
        with contextlib.redirect_stdout(sink):
            return f()
    return g
L = workloads.LIM_XY; V = workloads.LIM_V
report("C1-like 256^2 maps (weighted)", N, timeit(quiet(lambda: snap.velocity_moments(weight_name="mass", limXY=L, nXY=256))))
report("C1-like 256^2 maps (unweighted, 24 B)", N, timeit(quiet(lambda: snap.velocity_moments(limXY=L, nXY=256))), 24)
report("C2 128^2 maps + 128^2x256 cube", N, timeit(quiet(lambda: snap.project(weight_name="mass", limXY=L, nXY=128, limV=V, nV=256))))
def c2_full():
    m, l = snap.project(weight_name="mass", limXY=L, nXY=128, limV=V, nV=256)
    return l.convolve_spatial(1.0, 1.0)
report("C2 + smoothing (the bench step)", N, timeit(quiet(c2_full)))
report("C2 fixed-point accumulators", N, timeit(quiet(lambda: snap.project(weight_name="mass", limXY=L, nXY=128, limV=V, nV=256, accumulate="fixed"))))
report("C2 cube only (create_losvds)", N, timeit(quiet(lambda: snap.create_losvds(weight_name="mass", limXY=L, nXY=128, limV=V, nV=256))))
report("C3 per-GPU share: 512^2 maps", N, timeit(quiet(lambda: snap.velocity_moments(weight_name="mass", limXY=L, nXY=512))))
report("C5 256^2 map (create_map) ", N, timeit(quiet(lambda: snap.create_map(weight_name="mass", limXY=L, nXY=256))))
report("C5 256^2x512 cube (create_losvds)", N, timeit(quiet(lambda: snap.create_losvds(weight_name="mass", limXY=L, nXY=256, limV=V, nV=512))))
report("C5 256^2 maps + 256^2x512 cube (project)", N, timeit(quiet(lambda: snap.project(weight_name="mass", limXY=L, nXY=256, limV=V, nV=512))))
srt = workloads.morton_sort(soa)
snap_s = pnm.snapshot.from_arrays(srt[:3], srt[3:6], mass=srt[6], output="torch")
snap_s.rotate(30., 60.)
report("C2 on a Morton-ordered snapshot", N, timeit(quiet(lambda: snap_s.project(weight_name="mass", limXY=L, nXY=128, limV=V, nV=256))))
report("C1-like 256^2 maps, Morton-ordered", N, timeit(quiet(lambda: snap_s.velocity_moments(weight_name="mass", limXY=L, nXY=256))))
del snap_s, srt
incs = np.repeat(np.linspace(0, 87.5, 8), 8); pas = np.tile(np.linspace(0, 157.5, 8), 8)
n4 = min(N, 20_000_000)
sub = pnm.snapshot.from_arrays(soa[:3, :n4], soa[3:6, :n4], mass=soa[6, :n4], output="torch")
t = timeit(quiet(lambda: sub.project_views(pas, incs, weight_name="mass", limXY=L, nXY=256)), reps=2, warm=1)
report("C4 64 views x 256^2 maps, one read (n=%.0e)" % n4, n4, t, 28, "= %.2f G particle-views/s" % (64 * n4 / t / 1e6))
def loop64():
    for p, i in zip(pas, incs):
        sub.rotate(float(p), float(i)); sub.velocity_moments(weight_name="mass", limXY=L, nXY=256)
t2 = timeit(quiet(loop64), reps=2, warm=1)
report("C4 as 64 separate rotate+velocity_moments calls", n4, t2, 28 * 64, "= %.2f G particle-views/s" % (64 * n4 / t2 / 1e6))
d = soa[:, :50_000_000].double()
snap64 = pnm.snapshot.from_arrays(d[:3], d[3:6], mass=d[6], output="torch"); snap64.rotate(30., 60.)
report("C2 float64 particles (56 B), n=5e7", 50_000_000, timeit(quiet(lambda: snap64.project(weight_name="mass", limXY=L, nXY=128, limV=V, nV=256))), 56)
