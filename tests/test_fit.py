"""f4 -- Gauss-Hermite fit per spaxel (pynmap/losvd.py:53-67, fit_losvd.py:57-235).

tests/golden/fit.npz holds noisy synthetic profiles and a toy-snapshot cube with the parameters the
reference itself fitted (fitgh -> mpfit), made by oracle/make_golden.py.  The fit is tolerance
defined (mpfit stops at ftol = xtol = gtol = 1e-10 with a finite-difference Jacobian), so the bars
are: parameters within TOL of the reference's (I relative; V and S in units of S; h_n absolute),
chi2 not worse than the reference's, and first-order optimality of the restated objective.
"""
import numpy as np
import pytest

from oracle import oracle_np as on

TOL = 2e-5          # observed: < 3e-6 against mpfit, < 6e-6 against scipy's trust-region solver


def scaled_diff(gh, ref):
    """max |gh - ref| in the units described in the module docstring; rows are profiles."""
    scale = np.ones_like(ref)
    scale[:, 0] = np.abs(ref[:, 0])
    scale[:, 1] = scale[:, 2] = ref[:, 2]
    return np.max(np.abs(gh - ref) / scale, axis=1)


# ------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_objective_matches_reference(golden):
    g = golden("fit")
    x = g["x"]
    for deg in (2, 3, 4, 5, 6):
        gh, data, best, fnorm = g["gh%d" % deg], g["data%d" % deg], g["best%d" % deg], g["fnorm%d" % deg]
        for i in range(len(gh)):
            if not np.isfinite(gh[i, 0]):
                continue
            assert np.array_equal(on.gausshermite(x, gh[i]), best[i])                  # bit-exact profile
            assert on.gh_amplitude(data[i], on.gausshermite(x, np.r_[1.0, gh[i, 1:]])) == gh[i, 0]
            chi2 = np.sum(on.gh_residuals(gh[i, 1:], x, data[i]) ** 2)
            assert abs(chi2 - fnorm[i]) <= 1e-12 * fnorm[i]
    i = 3
    chi2 = np.sum(on.gh_residuals(g["gh_err"][i, 1:], x, g["data_err"][i], g["err"][i]) ** 2)
    assert abs(chi2 - g["fnorm_err"][i]) <= 1e-12 * chi2


def test_oracle_start_and_limits(golden):
    g = golden("fit")
    x = g["x"]
    p0, lo, hi = on.gh_guess_and_limits(x, g["data4"][-1], 4)           # the all-zero profile
    assert g["status4"][-1] == -16 and np.isnan(g["gh4"][-1, 0])
    assert np.array_equal(p0, g["gh4"][-1, 1:])                          # the reference returns its start values
    assert lo[1] == np.min(np.diff(x)) / 3 and hi[1] == (x[-1] - x[0]) / 3 and hi[2] == 0.2
    inside = (g["gh4"][:-1, 1:] >= lo) & (g["gh4"][:-1, 1:] <= hi)
    assert inside.all()
    assert (np.abs(g["gh4"][:-1, 3:]) == 0.2).any()                      # some fits end pegged at a limit


def test_scipy_checker_agrees_with_reference(golden):
    g = golden("fit")
    x = g["x"]
    for deg, idx in ((4, range(0, 24)), (2, range(4)), (6, range(4))):
        ref = g["gh%d" % deg]
        got = np.stack([on.fit_gh_scipy(x, g["data%d" % deg][i], deg)[0] for i in idx])
        assert scaled_diff(got, ref[list(idx)]).max() < TOL
    got = np.stack([on.fit_gh_scipy(x, g["data_err"][i], 4, g["err"][i])[0] for i in range(4)])
    assert scaled_diff(got, g["gh_err"][:4]).max() < TOL


# ------------------------------------------------------------------------------------ GPU: kernel
@pytest.mark.gpu
def test_gpu_fit_matches_reference_fits(golden):
    from pynmap_b200 import fit_losvd
    g = golden("fit")
    x = g["x"]
    for deg in (2, 3, 4, 5, 6):
        data, ref, ref_chi2 = g["data%d" % deg], g["gh%d" % deg], g["fnorm%d" % deg]
        gh, best, status, niter, chi2 = fit_losvd.fit_profiles(x, data, deg_gh=deg)
        gh, best, status, chi2 = gh.cpu().numpy().T, best.cpu().numpy(), status.cpu().numpy(), chi2.cpu().numpy()
        ok = np.isfinite(ref[:, 0])
        assert np.isin(status[ok], (1, 2, 3, 4)).all(), status
        assert scaled_diff(gh[ok], ref[ok]).max() < TOL, (deg, scaled_diff(gh[ok], ref[ok]).max())
        assert (chi2[ok] <= ref_chi2[ok] * (1 + 1e-9)).all()
        for i in np.nonzero(ok)[0][:8]:                                   # outputs are self-consistent
            assert np.allclose(best[i], on.gausshermite(x, gh[i]), rtol=1e-12, atol=1e-300)
            assert abs(np.sum((data[i] - best[i]) ** 2) - chi2[i]) <= 1e-10 * chi2[i]
        if deg == 4:                                                      # all-zero profile
            assert status[-1] == -16 and np.isnan(gh[-1, 0]) and np.isnan(best[-1]).all()
            assert np.array_equal(gh[-1, 1:], ref[-1, 1:])


@pytest.mark.gpu
def test_gpu_fit_with_errors(golden):
    from pynmap_b200 import fit_losvd
    g = golden("fit")
    gh, best, status, niter, chi2 = fit_losvd.fit_profiles(g["x"], g["data_err"], deg_gh=4, err=g["err"])
    assert scaled_diff(gh.cpu().numpy().T, g["gh_err"]).max() < TOL
    assert (chi2.cpu().numpy() <= g["fnorm_err"] * (1 + 1e-9)).all()


@pytest.mark.gpu
def test_gpu_fit_with_user_parinfo(golden):
    """Partial start values, a fixed parameter, unlimited h_n, tight limits (fit_losvd.py:166-176)."""
    from pynmap_b200 import fit_losvd
    g = golden("fit")
    x, data = g["x"], g["data_custom"]
    custom = {"A": dict(params=[10.0, 80.0]),
              "B": dict(fixed=[False, False, True, False]),
              "C": dict(limitedmin=[True, True, False, False], limitedmax=[True, True, False, False]),
              "D": dict(minpars=[-100.0, 30.0, -0.05, -0.05], maxpars=[100.0, 120.0, 0.05, 0.05],
                        params=[0.0, 60.0, 0.0, 0.0])}
    for tag, kw in custom.items():
        gh, best, status, niter, chi2 = fit_losvd.fit_profiles(x, data, deg_gh=4, **kw)
        gh = gh.cpu().numpy().T
        diff = scaled_diff(gh, g["gh_custom" + tag])
        assert diff.max() < TOL, (tag, diff.max())
        assert (chi2.cpu().numpy() <= g["fnorm_custom" + tag] * (1 + 1e-9)).all()
    assert (gh[:, 3:] >= -0.05).all() and (gh[:, 3:] <= 0.05).all() and (np.abs(gh[:, 3:]) == 0.05).any()
    one = fit_losvd.fitgh(x, data[2], deg_gh=4, verbose=False, **custom["B"])
    assert one[0][3] == 0.02 and scaled_diff(one[0][None], g["gh_customB"][2:3]).max() < TOL


@pytest.mark.gpu
def test_gpu_fitlosvd_on_cube(golden):
    import torch
    from pynmap_b200 import LOSVD
    g = golden("fit")
    nx, ny, nv = g["cube"].shape
    lv = LOSVD(np.arange(nx) * 1.0, np.arange(ny) * 1.0, g["cube_binv"], g["cube"])
    lv.fitlosvd(degree=4, verbose=False)
    assert lv.GH.shape == (5, nx, ny) and lv.bestfit.shape == (nx, ny, nv) and isinstance(lv.GH, np.ndarray)
    diff = scaled_diff(lv.GH.reshape(5, -1).T, g["cube_GH"].reshape(5, -1).T)
    assert diff.max() < TOL, diff.max()
    assert np.allclose(lv.bestfit, g["cube_bestfit"], rtol=0, atol=2e-4 * g["cube"].max())
    assert len(lv.resfit) == nx * ny
    r = lv.resfit[nx + 3]                                                 # loop order: ix outer, iy inner
    assert r.status in (1, 2, 3, 4) and abs(r.fnorm - g["cube_fnorm"][1, 3]) <= 1e-8 * r.fnorm
    assert np.array_equal(r.params, lv.GH[1:, 1, 3])
    # device in, device out
    lv_dev = LOSVD(lv.binx, lv.biny, g["cube_binv"], torch.from_numpy(g["cube"]).cuda())
    lv_dev.fitlosvd(degree=4)
    assert lv_dev.GH.is_cuda and np.array_equal(lv_dev.GH.cpu().numpy(), lv.GH)


@pytest.mark.gpu
def test_gpu_fitgh_single_profile_and_errors(golden, capsys):
    from pynmap_b200 import fit_losvd, LOSVD
    g = golden("fit")
    x = g["x"]
    gh, res, best = fit_losvd.fitgh(x, g["data4"][5], deg_gh=4, verbose=True)
    assert "Chi2:" in capsys.readouterr().out
    assert scaled_diff(gh[None], g["gh4"][5:6]).max() < TOL and res.status in (1, 2, 3, 4)
    assert np.array_equal(res.params, gh[1:]) and best.shape == x.shape
    assert np.allclose(fit_losvd.gausshermite(x, gh), best, rtol=1e-12, atol=1e-300)
    # a single occupied channel: sigma = 0 < lower limit -> the reference raises out of mpfit
    spike = np.zeros_like(x)
    spike[20] = 7.0
    with pytest.raises(Exception, match="not within PARINFO limits"):
        fit_losvd.fitgh(x, spike, deg_gh=4, verbose=False)
    cube = np.stack([g["data4"][0], spike]).reshape(1, 2, -1)
    lv = LOSVD(np.arange(2.0), np.arange(2.0), x, cube)
    with pytest.raises(Exception, match="not within PARINFO limits"):
        lv.fitlosvd(degree=4)
    lv.fitlosvd(degree=4, on_error="flag")
    assert lv.resfit[1].status == 0 and np.isnan(lv.GH[:, 0, 1]).all() and np.isfinite(lv.GH[:, 0, 0]).all()
    out = fit_losvd.fitgh(x, g["data4"][0], deg_gh=4, params=[0.0, 50.0, 0.0, 0.0, 0.0], verbose=False)
    assert out == (0., 0., 0., 0.) and "larger than expected" in capsys.readouterr().out     # fit_losvd.py:140-146
    with pytest.raises(Exception, match="no free parameters"):
        fit_losvd.fitgh(x, g["data4"][0], deg_gh=2, params=[0.0, 50.0], fixed=[True, True], verbose=False)
    with pytest.raises(Exception, match="limits are not consistent"):
        fit_losvd.fitgh(x, g["data4"][0], deg_gh=2, minpars=[50.0, 10.0], maxpars=[-50.0, 100.0], verbose=False)
    with pytest.raises(Exception):
        fit_losvd.fit_profiles(x, g["data4"], deg_gh=7)                  # outside the compiled range, loud


@pytest.mark.gpu
def test_gpu_fit_full_size_recovers_exact_models():
    """128 x 128 spaxels x 256 channels (BASELINE configs[1]'s cube): noiseless Gauss-Hermite
    profiles inside the limits must be recovered (chi2 -> 0), whatever the optimiser's path."""
    import torch
    from pynmap_b200 import fit_losvd
    nx = ny = 128
    edges = np.linspace(-500.0, 500.0, 257)
    x = 0.5 * (edges[1:] + edges[:-1])
    rng = np.random.default_rng(3)
    truth = np.stack([rng.uniform(1.0, 1e4, nx * ny), rng.uniform(-150, 150, nx * ny), rng.uniform(30, 110, nx * ny),
                      rng.uniform(-0.12, 0.12, nx * ny), rng.uniform(-0.12, 0.12, nx * ny)])
    xt = torch.from_numpy(x).cuda()
    t = torch.from_numpy(truth).cuda()
    w = (xt[None, :] - t[1][:, None]) / t[2][:, None]
    h3 = (2 * w * w - 3) * w / np.sqrt(3.0)
    h4 = (np.sqrt(2.0) * h3 * w - (2 * w * w - 1) / np.sqrt(2.0)) / 2.0
    cube = (t[0][:, None] * (1 + t[3][:, None] * h3 + t[4][:, None] * h4) * torch.exp(-0.5 * w * w)).reshape(nx, ny, -1)
    gh, best, status, niter, chi2 = fit_losvd.fit_profiles(x, cube, deg_gh=4)
    status = status.cpu().numpy()
    assert np.isin(status, (1, 2, 3, 4)).all(), np.unique(status, return_counts=True)
    diff = scaled_diff(gh.reshape(5, -1).cpu().numpy().T, truth.T)
    assert diff.max() < 1e-6, diff.max()
    assert float((chi2.reshape(-1) / (cube * cube).sum(-1).reshape(-1)).max()) < 1e-12
