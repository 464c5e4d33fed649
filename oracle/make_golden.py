"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.npz by executing the UNMODIFIED
reference (``/root/reference/pynmap``) under ``oracle/refshim.py``.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference has no tests, fixtures or golden vectors of its own, so
these files ARE the pin: inputs and the reference's outputs for every function on the hot
path (rotation.py:73-119, :14-71; grid.py:70-262; losvd.py:94-141; pynmap.py:246-455),
including the quirks listed in SURVEY.md 8(c).  Everything is stored bit-exactly (npz).
numpy 2.3.5 / scipy 1.18.1 produced the committed vectors.  The smoothing vectors go through
the shim's restatement of astropy (astropy is absent) and are therefore NOT a pin of astropy.
"""
import contextlib
import io
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.refshim import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def toy_snapshot(n, seed, dtype):
    """Small exponential disk + bulge in numpy (kpc, km/s); only used to make fixtures."""
    rng = np.random.default_rng(seed)
    nb = n // 5
    nd = n - nb
    R = rng.gamma(2.0, 3.0, nd)
    phi = rng.uniform(0, 2 * np.pi, nd)
    z = 0.3 * np.arctanh(rng.uniform(-0.999, 0.999, nd))
    vc = 220.0 * R / np.sqrt(R * R + 1.0)
    sr = 55.0 * np.exp(-R / 6.0) + 5.0
    vr, vp, vz = sr * rng.standard_normal(nd), vc + sr / np.sqrt(2) * rng.standard_normal(nd), 0.6 * sr * rng.standard_normal(nd)
    disk_pos = np.stack([R * np.cos(phi), R * np.sin(phi), z], 1)
    disk_vel = np.stack([vr * np.cos(phi) - vp * np.sin(phi), vr * np.sin(phi) + vp * np.cos(phi), vz], 1)
    r = 0.5 / np.sqrt(rng.uniform(1e-6, 1, nb) ** (-2.0 / 3.0) - 1.0 + 1e-12)
    u = rng.standard_normal((nb, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    sig = 120.0 / (1.0 + (r / 0.5) ** 2) ** 0.25
    pos = np.concatenate([disk_pos, r[:, None] * u])
    vel = np.concatenate([disk_vel, sig[:, None] * rng.standard_normal((nb, 3))])
    perm = rng.permutation(n)
    mass = rng.uniform(0.5, 2.0, n)
    return pos[perm].astype(dtype), vel[perm].astype(dtype), mass.astype(dtype)


def main():
    ref = load_reference()
    rot, grid, losvd_mod = ref.rotation, ref.grid, ref.losvd
    os.makedirs(OUT, exist_ok=True)
    saved = {}

    def save(name, **arrays):
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        saved[name] = sum(np.asarray(a).nbytes for a in arrays.values())

    # ---- rotation: float32 / float64 data x python-float / float32 angles x direct / inverse
    for dt in (np.float32, np.float64):
        pos, vel, _ = toy_snapshot(1024, 11, dt)
        pos[7] = [np.inf, 1.0, -2.0]      # non-finite coordinates travel through np.dot
        pos[9] = [np.nan, 0.0, 0.0]
        out = {"pos": pos, "vel": vel}
        for tag, (pa, inc) in {"f64ang": (30.0, 60.0), "f32ang": (np.float32(30.0), np.float32(60.0)),
                               "odd": (12.3, 77.7)}.items():
            for direct in (True, False):
                p, v = rot.rotate_mat(pos, vel, pa, inc, direct)
                k = "%s_%s" % (tag, "dir" if direct else "inv")
                out["pos_" + k], out["vel_" + k] = np.ascontiguousarray(p), np.ascontiguousarray(v)
        p, v = rot.rotateXYZ_mat(pos, vel, 10.0, 20.0, 30.0, rotorder=[2, 0, 1])
        out["pos_xyz"], out["vel_xyz"] = np.ascontiguousarray(p), np.ascontiguousarray(v)
        p, _ = rot.rotate_mat(pos, None, 45.0, 45.0)
        out["pos_novel"] = np.ascontiguousarray(p)
        save("rotate_%s" % np.dtype(dt).name, **out)

    # ---- KAT: edge inclusivity, NaN / inf dropping, matrices, orientation
    xk = np.array([-1.5, -0.5, 0.5, 1.5, 1.5000001, -1.5000001, np.nan, np.inf, -np.inf, 0.0])
    gx, gy, M, V, S = quiet(grid.points_2vmaps, xk, np.zeros_like(xk), np.ones_like(xk), limXY=[-1, 1, -1, 1], nXY=3)
    vq = np.array([1.0, np.nan, 3.0])
    _, _, M_nomask, V_nomask, _ = quiet(grid.points_2vmaps, np.zeros(3), np.zeros(3), vq, limXY=[-1, 1, -1, 1], nXY=3)
    _, _, M_mask, V_mask, _ = quiet(grid.points_2vmaps, np.zeros(3), np.zeros(3), vq, limXY=[-1, 1, -1, 1], nXY=3,
                                    mask=np.array([True, True, False]))
    save("kat", x_edge=xk, M_edge=M, V_edge=V, S_edge=S, gridX_edge=gx, gridY_edge=gy,
         v_quirk=vq, M_nomask=M_nomask, V_nomask=V_nomask, M_mask=M_mask, V_mask=V_mask,
         mat_f64ang=_matrix(rot, 30.0, 60.0), mat_f32ang=_matrix(rot, np.float32(30.0), np.float32(60.0)))

    # ---- maps and cubes on a toy snapshot, float32 and float64, weighted and not
    for dt in (np.float32, np.float64):
        pos, vel, mass = toy_snapshot(12000, 5, dt)
        p, v = rot.rotate_mat(pos, vel, np.float32(30.0), np.float32(60.0))
        x, y, d = p[:, 0], p[:, 1], v[:, 2]
        lim, nxy, limv, nv = [-15, 15, -12, 12], (40, 32), [-400, 400], 20
        out = {"pos": pos, "vel": vel, "mass": mass, "x": np.ascontiguousarray(x), "y": np.ascontiguousarray(y),
               "d": np.ascontiguousarray(d), "limXY": np.array(lim, float), "nXY": np.array(nxy), "limV": np.array(limv, float),
               "nV": np.array(nv)}
        for wtag, w in (("w", mass), ("u", None)):
            gx, gy, M, V, S = quiet(grid.points_2vmaps, x, y, d, weights=w, limXY=lim, nXY=nxy)
            out.update({"vm_%s_M" % wtag: M, "vm_%s_V" % wtag: V, "vm_%s_S" % wtag: S})
            _, _, W, D, DW = quiet(grid.points_2map, x, y, d, weights=w, limXY=lim, nXY=nxy)
            out.update({"map_%s_W" % wtag: W, "map_%s_D" % wtag: D, "map_%s_DW" % wtag: DW})
            L = quiet(grid.points_2losvds, x, y, d, weights=w, limXY=lim, nXY=nxy, limV=limv, nV=nv)
            out["cube_%s" % wtag] = L.losvd
        out.update({"gridX": gx, "gridY": gy, "binx": L.binx, "biny": L.biny, "binv": L.binv})
        # raw sums, for tolerance tests that do not go through the ill-conditioned sigma
        ex = np.linspace(lim[0] - (lim[1] - lim[0]) / (nxy[0] - 1) / 2., lim[1] + (lim[1] - lim[0]) / (nxy[0] - 1) / 2., nxy[0] + 1)
        ey = np.linspace(lim[2] - (lim[3] - lim[2]) / (nxy[1] - 1) / 2., lim[3] + (lim[3] - lim[2]) / (nxy[1] - 1) / 2., nxy[1] + 1)
        out["s1_w"] = np.histogram2d(x, y, [ex, ey], weights=mass * d)[0].T
        out["s2_w"] = np.histogram2d(x, y, [ex, ey], weights=mass * d ** 2)[0].T
        out["edges_x"], out["edges_y"] = ex, ey
        # mask with weights=None on the cube path (the legal combination)
        msk = np.zeros(x.size, bool)
        msk[::7] = True
        out["mask"] = msk
        out["cube_masked_u"] = quiet(grid.points_2losvds, x, y, d, limXY=lim, nXY=nxy, limV=limv, nV=nv, mask=msk).losvd
        dq = d.copy()
        dq[::11] = np.nan
        out["d_nan"] = dq
        _, _, Mq, Vq, Sq = quiet(grid.points_2vmaps, x, y, dq, weights=mass, limXY=lim, nXY=nxy, mask=msk)
        out.update({"vm_q_M": Mq, "vm_q_V": Vq, "vm_q_S": Sq})
        # smoothing (shim restatement of astropy): equal and unequal widths
        Lw = losvd_mod.LOSVD(L.binx, L.biny, L.binv, out["cube_w"])
        c1 = quiet(Lw.convolve_spatial, 1.5, 1.5)
        c2 = quiet(Lw.convolve_spatial, 1.2, 2.4)
        c3 = quiet(c1.convolve_spatial, 2.5, 2.0)
        out.update({"smooth_eq": c1.losvd, "smooth_neq": c2.losvd, "smooth_twice": c3.losvd})
        quiet(Lw.vmoments)
        out.update({"mom0": Lw.mom0, "mom1": Lw.mom1, "mom2": Lw.mom2})
        # integer-factor rebin (losvd.py:69-92): sum and mean, coordinates through the same function
        rs, rm = quiet(Lw.rebin_spatial, 4, 2), quiet(Lw.rebin_spatial, 2, 4, "mean")
        out.update({"rebin_sum": rs.losvd, "rebin_sum_binx": rs.binx, "rebin_sum_biny": rs.biny,
                    "rebin_mean": rm.losvd, "rebin_mean_binx": rm.binx, "rebin_mean_biny": rm.biny})
        save("grid_%s" % np.dtype(dt).name, **out)

    # ---- snapshot level: ascii file -> rotate -> velocity_moments / create_map / create_losvds
    pos, vel, mass = toy_snapshot(4000, 21, np.float64)
    with tempfile.TemporaryDirectory() as tmp:
        np.savetxt(os.path.join(tmp, "snap.txt"), np.hstack([pos, vel, mass[:, None]]), fmt="%.17g")
        snap = quiet(ref.pynmap.snapshot, folder=tmp, snap_name="snap.txt", MLrecipe="none", extra_labels=["mass"], verbose=False)
        quiet(snap.rotate, PA=30., inclination=60.)
        vm = quiet(snap.velocity_moments, weight_name="mass", limXY=[-12, 12, -12, 12], nXY=32)
        mp = quiet(snap.create_map, weight_name="mass", limXY=[-12, 12, -12, 12], nXY=32)
        lv = quiet(snap.create_losvds, weight_name="mass", limXY=[-12, 12, -12, 12], nXY=16, limV=[-350, 350], nV=20)
        pos1, vel1 = np.ascontiguousarray(snap.pos), np.ascontiguousarray(snap.vel)
        quiet(snap.rotate, PA=15., inclination=20., reset=False)         # compounded on rounded coordinates
        vm2 = quiet(snap.velocity_moments, weight_name="mass", limXY=[-12, 12, -12, 12], nXY=32)
        pos2 = np.ascontiguousarray(snap.pos)
        PAs2, incs2 = np.array(snap.PAs), np.array(snap.inclinations)
    save("snapshot", pos=pos, vel=vel, mass=mass, M=vm.M, V=vm.V, S=vm.S, X=vm.X, Y=vm.Y, W=mp.W, D=mp.D, DW=mp.DW,
         cube=lv.losvd, binv=lv.binv, pos_rot=pos1, vel_rot=vel1, M2=vm2.M, V2=vm2.V, S2=vm2.S, pos_rot2=pos2,
         PAs2=PAs2, incs2=incs2)

    # ---- f2: SSP mass-to-light ratios (misc_functions.py:115-207) with the reference's E-MILES tables.
    # Stored: the (age, [M/H]) -> M/L nodes the reference interpolates (its own intermediate `sspML`,
    # observed by replaying misc_functions.py:176-193 on its tables), query points, and its output.
    rng = np.random.default_rng(31)
    qa, qm = rng.uniform(0.0, 14.5, 4000), rng.uniform(-2.4, 0.5, 4000)
    ML = quiet(ref.misc_functions.compute_ml_ssps, np.ones(4000), qa, qm)
    ML_nofill = quiet(ref.misc_functions.compute_ml_ssps, np.ones(4000), qa, qm, fill_nan=False)
    mf = ref.misc_functions
    libdir = os.path.join(os.path.dirname(os.path.realpath(mf.__file__)), "SSPLibraries")
    mData = quiet(mf.SSPMass, os.path.join(libdir, "out_mass_KB_BASTI.txt"))
    lData = quiet(mf.SSPPhot, os.path.join(libdir, "out_phot_KB_BASTI.txt"))
    mw = np.isclose(1.30, mData['slope'], atol=1e-2)
    lw = np.isclose(1.30, lData['slope'], atol=1e-2)
    nodes_ml = mData['mSRem'][mw] / 10 ** (-0.4 * (lData['V'][lw] - 4.820))
    save("ssp", age=qa, MH=qm, ML=ML, ML_nofill=ML_nofill, node_age=mData['age'][mw], node_MH=mData['MH'][mw], node_ML=nodes_ml)

    make_fit_golden(ref, save)
    make_amr_golden(ref, save)
    make_profiles_golden(ref, save)

    for k, v in saved.items():
        print("%-16s %8.1f kB (uncompressed)" % (k, v / 1e3))


def make_fit_golden(ref, save):
    """Row f4: the reference's own Gauss-Hermite fits (fit_losvd.fitgh -> mpfit, fit_losvd.py:84-235;
    LOSVD.fitlosvd, losvd.py:53-67) on noisy synthetic profiles and on a cube of a toy snapshot."""
    fl, mf = ref.fit_losvd, ref.misc_functions
    edges = np.linspace(-500.0, 500.0, 65)
    x = 0.5 * (edges[1:] + edges[:-1])
    rng = np.random.default_rng(77)

    def run(profiles, deg, errs=None):
        gh = np.zeros((len(profiles), deg + 1))
        status, niter, fnorm = (np.zeros(len(profiles), dtype=t) for t in (np.int32, np.int32, np.float64))
        best = np.zeros_like(profiles)
        for i, d in enumerate(profiles):
            with np.errstate(all="ignore"):
                gh[i], res, best[i] = quiet(fl.fitgh, x, d, deg_gh=deg, err=None if errs is None else errs[i],
                                            verbose=False)
            status[i], niter[i], fnorm[i] = res.status, res.niter, res.fnorm
        return gh, status, niter, fnorm, best

    out = {"x": x}
    for deg, count in ((4, 96), (2, 16), (3, 16), (5, 16), (6, 16)):
        truth = np.zeros((count, 7))
        truth[:, 0] = rng.uniform(2.0, 200.0, count)
        truth[:, 1] = rng.uniform(-300.0, 300.0, count)
        truth[:, 2] = rng.uniform(20.0, 180.0, count)
        truth[:, 3:] = rng.uniform(-0.25, 0.25, (count, 4))       # beyond +-0.2: fits end pegged at a limit
        prof = np.stack([rng.poisson(np.maximum(mf.gausshermite(x, t[:deg + 1]), 0.0)).astype(np.float64)
                         for t in truth])
        if deg == 4:
            prof[-1] = 0.0                                        # all-zero spaxel: I = NaN, status -16
        gh, status, niter, fnorm, best = run(prof, deg)
        out.update({"data%d" % deg: prof, "gh%d" % deg: gh, "status%d" % deg: status, "niter%d" % deg: niter,
                    "fnorm%d" % deg: fnorm, "best%d" % deg: best})
    # with an error array (fit_losvd.py:180-185, :442-446)
    truth = np.stack([rng.uniform(20.0, 200.0, 16), rng.uniform(-200.0, 200.0, 16), rng.uniform(40.0, 150.0, 16),
                      rng.uniform(-0.15, 0.15, 16), rng.uniform(-0.15, 0.15, 16)], 1)
    prof = np.stack([rng.poisson(np.maximum(mf.gausshermite(x, t), 0.0)).astype(np.float64) for t in truth])
    errs = np.sqrt(np.maximum(prof, 1.0))
    gh, status, niter, fnorm, best = run(prof, 4, errs)
    out.update(data_err=prof, err=errs, gh_err=gh, status_err=status, niter_err=niter, fnorm_err=fnorm, best_err=best)

    # non-default parinfo (fit_losvd.py:166-176): partial start values, a fixed parameter, unlimited h_n, tight limits
    truth = np.stack([rng.uniform(20.0, 200.0, 12), rng.uniform(-90.0, 90.0, 12), rng.uniform(40.0, 110.0, 12),
                      rng.uniform(-0.15, 0.15, 12), rng.uniform(-0.15, 0.15, 12)], 1)
    prof = np.stack([rng.poisson(np.maximum(mf.gausshermite(x, t), 0.0)).astype(np.float64) for t in truth])
    out["data_custom"] = prof
    custom = {"A": dict(params=[10.0, 80.0]),
              "B": dict(fixed=[False, False, True, False]),
              "C": dict(limitedmin=[True, True, False, False], limitedmax=[True, True, False, False]),
              "D": dict(minpars=[-100.0, 30.0, -0.05, -0.05], maxpars=[100.0, 120.0, 0.05, 0.05], params=[0.0, 60.0, 0.0, 0.0])}
    for tag, kw in custom.items():
        gh = np.zeros((len(prof), 5))
        fn = np.zeros(len(prof))
        for i, d in enumerate(prof):
            with np.errstate(all="ignore"):
                gh[i], res, _ = quiet(fl.fitgh, x, d, deg_gh=4, verbose=False, **kw)
            fn[i] = res.fnorm
        out["gh_custom" + tag] = gh
        out["fnorm_custom" + tag] = fn

    # LOSVD.fitlosvd on a populated cube
    pos, vel, mass = toy_snapshot(60000, 5, np.float64)
    lim = [-3.0, 3.0, -3.0, 3.0]
    lv = quiet(ref.grid.points_2losvds, pos[:, 0], pos[:, 1], vel[:, 2], weights=mass, limXY=lim, nXY=8,
               limV=[-450.0, 450.0], nV=36)
    with np.errstate(all="ignore"):
        quiet(lv.fitlosvd, degree=4, verbose=False)
    out.update(cube=lv.losvd, cube_binv=lv.binv, cube_GH=lv.GH, cube_bestfit=lv.bestfit,
               cube_status=np.array([r.status for r in lv.resfit], dtype=np.int32).reshape(8, 8),
               cube_fnorm=np.array([r.fnorm for r in lv.resfit]).reshape(8, 8))
    save("fit", **out)


def make_amr_golden(ref, save):
    """Row f3, AMR part: misc_io.spread_amr (misc_io.py:184-224) and snapshot.amr_velocity_moments
    (pynmap.py:342-419, smear=False) with numpy's global generator seeded before each call."""
    pos, vel, mass = toy_snapshot(3000, 9, np.float64)
    rng = np.random.default_rng(4)
    level = rng.integers(6, 11, 3000).astype(np.float64)
    out = {"pos": pos, "vel": vel, "mass": mass, "level": level, "seed": np.array(1234)}
    lim = [-12.0, 12.0, -12.0, 12.0]
    with tempfile.TemporaryDirectory() as tmp:
        np.savetxt(os.path.join(tmp, "snap.txt"), np.hstack([pos, vel, mass[:, None], level[:, None]]), fmt="%.17g")
        snap = quiet(ref.pynmap.snapshot, folder=tmp, snap_name="snap.txt", MLrecipe="none",
                     extra_labels=["mass", "level"], verbose=False)
    quiet(snap.rotate, PA=30., inclination=60.)
    with np.errstate(all="ignore"):
        np.random.seed(1234)
        a = quiet(snap.amr_velocity_moments, weight_name="mass", box=100., limXY=lim, nXY=32)
        np.random.seed(1234)
        b = quiet(snap.amr_velocity_moments, box=100., limXY=lim, nXY=[24, 40])            # unit weights
        c = quiet(snap.amr_velocity_moments, weight_name="mass", box=100., limXY=lim, nXY=32, spread=False)
    out.update(M=a.M, V=a.V, S=a.S, Mu=b.M, Vu=b.V, Su=b.S, Mn=c.M, Vn=c.V, Sn=c.S)
    # spread_amr itself: float32 coordinates / values, integer levels, a stretch, an integer constant
    x32, w32 = pos[:500, 0].astype(np.float32), mass[:500].astype(np.float32)
    lev_i = level[:500].astype(np.int64)
    np.random.seed(77)
    cr, vr, kr = ref.misc_io.spread_amr([x32, pos[:500, 1]], lev_i, list_val=[w32, lev_i], list_const=[w32, lev_i],
                                        nsample=7, box=50., stretch=[1., 0.5])
    out.update(sp_x=cr[0], sp_y=cr[1], sp_w=vr[0], sp_li=vr[1], sp_cw=kr[0], sp_cl=kr[1])
    save("amr", **out)


def make_profiles_golden(ref, save):
    """grid.point_2vprofiles (grid.py:16-68).  One float32 particle set is stored; the float64 case
    runs on its exact upcast."""
    pos32, vel32, _ = toy_snapshot(20000, 13, np.float32)
    out = {"pos": pos32, "vel": vel32}
    for dt in (np.float32, np.float64):
        pos, vel = pos32.astype(dt), vel32.astype(dt)
        tag = np.dtype(dt).name
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Rbin, v, s = ref.grid.point_2vprofiles(pos[:, 0], pos[:, 1], pos[:, 2], vel[:, 0], vel[:, 1], vel[:, 2], nbins=40)
            Rb2, v2, s2 = ref.grid.point_2vprofiles(pos[:, 0], pos[:, 1], pos[:, 2], vel[:, 0], vel[:, 1], vel[:, 2], nbins=25,
                                                    dz=np.array([-0.5, 0.1]), Rmin=0.05, Rmax=20.0)
        out.update({"Rbin_" + tag: Rbin, "v_" + tag: v, "s_" + tag: s, "Rbin2_" + tag: Rb2, "v2_" + tag: v2, "s2_" + tag: s2})
    save("profiles", **out)


def _matrix(rot, pa, inc):
    """The float32 matrix rotate_mat builds, observed through its action on the unit vectors."""
    eye = np.eye(3, dtype=np.float32)
    p, _ = rot.rotate_mat(eye, None, pa, inc)
    return np.ascontiguousarray(p.T)      # p = (mat . I)^T


if __name__ == "__main__":
    main()
