"""Row f4 timing: Gauss-Hermite fit of a 128 x 128 x 256 cube (one kernel) against the oracle's
scipy checker on a sample of spaxels (CPU, 1 core).  Run on the GPU box:  python scripts/fit_bench.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynmap_b200 import fit_losvd  # noqa: E402
from oracle import oracle_np as on  # noqa: E402


def main():
    nx = ny = 128
    nv = 256
    edges = np.linspace(-500.0, 500.0, nv + 1)
    x = 0.5 * (edges[1:] + edges[:-1])
    rng = np.random.default_rng(0)
    n = nx * ny
    truth = np.stack([rng.uniform(5, 500, n), rng.uniform(-200, 200, n), rng.uniform(30, 120, n),
                      rng.uniform(-0.15, 0.15, n), rng.uniform(-0.15, 0.15, n)], 1)
    w = (x[None, :] - truth[:, 1:2]) / truth[:, 2:3]
    h3 = (2 * w * w - 3) * w / np.sqrt(3.0)
    h4 = (np.sqrt(2.0) * h3 * w - (2 * w * w - 1) / np.sqrt(2.0)) / 2.0
    model = truth[:, :1] * (1 + truth[:, 3:4] * h3 + truth[:, 4:5] * h4) * np.exp(-0.5 * w * w)
    cube_h = rng.poisson(np.maximum(model, 0)).astype(np.float64).reshape(nx, ny, nv)
    cube = torch.from_numpy(cube_h).cuda()
    out = {}
    for deg in (2, 4, 6):
        for _ in range(2):
            fit_losvd.fit_profiles(x, cube, deg_gh=deg)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            gh, best, status, niter, chi2 = fit_losvd.fit_profiles(x, cube, deg_gh=deg)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        st = status.cpu().numpy()
        out["deg%d" % deg] = {"ms": round(ms, 3), "spaxels_per_s": n / ms * 1e3, "mean_niter": float(niter.double().mean()),
                              "max_niter": int(niter.max()), "status": {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}}
    sample = rng.choice(n, 64, replace=False)
    t0 = time.perf_counter()
    ref = np.stack([on.fit_gh_scipy(x, cube_h.reshape(n, nv)[i], 4)[0] for i in sample])
    cpu_s = (time.perf_counter() - t0) / len(sample)
    gh4 = fit_losvd.fit_profiles(x, cube, deg_gh=4)[0].reshape(5, -1).cpu().numpy().T[sample]
    scale = np.ones_like(ref); scale[:, 0] = ref[:, 0]; scale[:, 1] = scale[:, 2] = ref[:, 2]
    out["cpu_scipy_s_per_spaxel"] = cpu_s
    out["cpu_scipy_spaxels_per_s"] = 1.0 / cpu_s
    out["max_scaled_diff_vs_scipy"] = float(np.max(np.abs(gh4 - ref) / scale))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
