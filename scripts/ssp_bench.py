"""Development: throughput of the SSP M/L interpolation kernel (pnm_ct_interp)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynmap_b200 import ssp
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ssp.npz"))
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 64
interp = ssp.MLInterpolator(g["node_age"], g["node_MH"], g["node_ML"], ncell=ncell)
print("ncell", ncell, "candidates/cell %.1f" % (interp.cells["tris"].size / ncell ** 2))
n = 100_000_000
gen = torch.Generator(device="cuda").manual_seed(3)
age = 0.03 + 13.97 * torch.rand(n, generator=gen, device="cuda")
mh = -2.27 + 2.67 * torch.rand(n, generator=gen, device="cuda")
for _ in range(2): interp(age, mh)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = interp(age, mh); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
t = min(ts)
print("pnm_ct_interp: %d particles in %.3f ms = %.1f Gpart/s, %.0f GB/s of 16 B/particle (8 in + 8 out)" % (n, t, n / t / 1e6, 16 * n / t / 1e6))
print("outside hull: %.3f" % float(torch.isnan(interp(age[:1000000], mh[:1000000], fill_nan=False)).float().mean()))
