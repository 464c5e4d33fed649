"""Synthetic N-body snapshots for the tests and bench.py (the reference ships no data).

Recipe = SURVEY.md section 8(d): units kpc and km/s; an exponential disk (R ~ Gamma(2, R_d = 3),
sech^2 vertical profile h_z = 0.3, V_c = 220 R / sqrt(R^2 + 1), sigma_R = 55 exp(-R / 2 R_d) + 5,
sigma_phi = sigma_R / sqrt 2, sigma_z = 0.6 sigma_R) plus, for "disk+bulge", a Plummer bulge
(fraction 0.2, a = 0.5, isotropic sigma = 120 / (1 + (r/a)^2)^(1/4)).  Particles are i.i.d., so
their order is random (the worst case for warp-level aggregation); ``morton_sort`` provides the
spatially sorted variant.  Generated with a seeded torch generator on the target device in chunks,
as float32 SoA rows (x, y, z, vx, vy, vz, mass).
"""
import math

import torch

R_D, H_Z = 3.0, 0.3
BULGE_A = 0.5
LIM_XY = [-15.0, 15.0, -15.0, 15.0]
LIM_V = [-500.0, 500.0]


def make_snapshot(n, bulge_fraction=0.2, seed=1234, device="cpu", dtype=torch.float32, chunk=1 << 24,
                  unit_mass=True):
    """-> (7, n) tensor: rows x, y, z, vx, vy, vz, mass."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    out = torch.empty((7, n), dtype=dtype, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        u = torch.rand((12, m), generator=gen, device=device, dtype=torch.float32)
        g = torch.randn((6, m), generator=gen, device=device, dtype=torch.float32)
        is_bulge = u[0] < bulge_fraction
        # ---- disk
        R = -R_D * (torch.log(u[1].clamp_min(1e-12)) + torch.log(u[2].clamp_min(1e-12)))
        phi = 2.0 * math.pi * u[3]
        z = H_Z * torch.atanh((2.0 * u[4] - 1.0).clamp(-0.999999, 0.999999))
        vc = 220.0 * R / torch.sqrt(R * R + 1.0)
        sr = 55.0 * torch.exp(-R / (2.0 * R_D)) + 5.0
        vr, vp, vz = sr * g[0], vc + sr / math.sqrt(2.0) * g[1], 0.6 * sr * g[2]
        c, sn = torch.cos(phi), torch.sin(phi)
        dx, dy, dz = R * c, R * sn, z
        dvx, dvy, dvz = vr * c - vp * sn, vr * sn + vp * c, vz
        # ---- bulge (Plummer)
        r = BULGE_A / torch.sqrt(u[5].clamp(1e-6, 1.0 - 1e-6) ** (-2.0 / 3.0) - 1.0)
        ct = 2.0 * u[6] - 1.0
        st = torch.sqrt((1.0 - ct * ct).clamp_min(0.0))
        ph = 2.0 * math.pi * u[7]
        sig = 120.0 / (1.0 + (r / BULGE_A) ** 2) ** 0.25
        bx, by, bz = r * st * torch.cos(ph), r * st * torch.sin(ph), r * ct
        sel = lambda a, b: torch.where(is_bulge, a, b).to(dtype)  # noqa: E731
        out[0, s:s + m] = sel(bx, dx)
        out[1, s:s + m] = sel(by, dy)
        out[2, s:s + m] = sel(bz, dz)
        out[3, s:s + m] = sel(sig * g[3], dvx)
        out[4, s:s + m] = sel(sig * g[4], dvy)
        out[5, s:s + m] = sel(sig * g[5], dvz)
        out[6, s:s + m] = 1.0 if unit_mass else (0.5 + 1.5 * u[8]).to(dtype)
        del u, g
    return out


def morton_sort(soa, bits=10, extent=40.0):
    """Reorder particles along a 3-D Morton (Z-order) curve of their positions -- a one-off
    per-snapshot preprocessing step (like a tree-code's own particle order), not part of the
    timed path."""
    def spread(v):
        v = v & 0x3FF
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    q = [((soa[k].float() / extent + 0.5).clamp(0.0, 0.999999) * (1 << bits)).to(torch.int64) for k in range(3)]
    key = spread(q[0]) | (spread(q[1]) << 1) | (spread(q[2]) << 2)
    order = torch.argsort(key)
    return soa[:, order].contiguous()
