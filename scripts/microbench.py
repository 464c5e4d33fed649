"""Development microbenchmark: times pnm_project_bin variants with CUDA events (not a bench line).
usage: python scripts/microbench.py [n_particles]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads  # noqa: E402
from pynmap_b200 import engine, rotation  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    soa = workloads.make_snapshot(n, seed=1235, device="cuda")
    rows = [soa[k] for k in range(6)]
    w = soa[6]
    mat = rotation.rotation_matrix(np.float32(30), np.float32(60))
    f64, fx = engine.AccMode("f64"), engine.AccMode("fixed")
    scales = engine.fixed_scales(n, 1.0, 2000.0)

    def case(name, nxy, nv, sums, acc=f64, limV=workloads.LIM_V, rows=rows, w=w, limXY=workloads.LIM_XY):
        grid = engine.get_grid(limXY, nxy, nxy, limV, nv)
        maps, cube = engine.new_accumulators(grid, sums | 1, acc)
        if sums & 8 and cube is None:
            _, cube = engine.new_accumulators(grid, 8, acc)
        def fn():
            engine.project_bin(grid, rows, w, mat, 1, 1, sums, acc, scales if acc is fx else None, maps, cube)
        best, med = timeit(fn)
        print("%-44s %8.3f ms (med %8.3f)  %7.1f GB/s  %6.2f Gpart/s" % (name, best, med, 28 * n / best / 1e6, n / best / 1e6), flush=True)

    print("n = %d" % n)
    case("no atomics (cube, V range empty)", 128, 256, 8, limV=[1e6, 1e6 + 1])
    case("no atomics, xy range empty", 128, 256, 15, limXY=[1e6, 1e6 + 1, 1e6, 1e6 + 1])
    case("cube only 128x128x256", 128, 256, 8)
    case("maps S0 only 128^2", 128, 0, 1)
    case("maps S1+S2 128^2", 128, 0, 6)
    case("maps S0+S1+S2 128^2", 128, 0, 7)
    case("maps S0+S1+S2 256^2", 256, 0, 7)
    case("maps S0+S1+S2 512^2", 512, 0, 7)
    case("full 128^2 x256 (4 REDs)", 128, 256, 15)
    case("full, W from cube (3 REDs)", 128, 256, 31)
    case("full, W from cube, FIXED", 128, 256, 31, acc=fx)
    case("cube only 256x256x512", 256, 512, 8)
    case("unweighted full, W from cube", 128, 256, 31, w=None)
    t = time.time()
    srt = workloads.morton_sort(soa)
    torch.cuda.synchronize()
    print("morton sort: %.2f s" % (time.time() - t))
    rows_s = [srt[k] for k in range(6)]
    case("SORTED cube only", 128, 256, 8, rows=rows_s, w=srt[6])
    case("SORTED maps S0+S1+S2 128^2", 128, 0, 7, rows=rows_s, w=srt[6])
    case("SORTED full, W from cube", 128, 256, 31, rows=rows_s, w=srt[6])
    # uniform box: no hot pixels
    uni = torch.empty_like(soa)
    uni[:3] = (torch.rand((3, n), device="cuda") - 0.5) * 29.0
    uni[3:6] = (torch.rand((3, n), device="cuda") - 0.5) * 600.0
    uni[6] = 1.0
    rows_u = [uni[k] for k in range(6)]
    case("UNIFORM cube only", 128, 256, 8, rows=rows_u, w=uni[6])
    case("UNIFORM maps S0+S1+S2 128^2", 128, 0, 7, rows=rows_u, w=uni[6])
    case("UNIFORM full, W from cube", 128, 256, 31, rows=rows_u, w=uni[6])


if __name__ == "__main__":
    main()
