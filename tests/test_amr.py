"""f3, AMR part -- spread_amr (pynmap/misc_io.py:184-224) and snapshot.amr_velocity_moments
(pynmap/pynmap.py:342-419, smear=False).  tests/golden/amr.npz holds the reference's outputs with
numpy's global generator seeded before each call; the host layer draws the same deviates, so the
jittered coordinates are bit-exact and the maps match like any other points_2vmaps call."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import oracle_np as on

LIM = [-12.0, 12.0, -12.0, 12.0]


def _spread_inputs(g):
    x32, w32 = g["pos"][:500, 0].astype(np.float32), g["mass"][:500].astype(np.float32)
    return x32, g["pos"][:500, 1], w32, g["level"][:500].astype(np.int64)


def test_oracle_spread_amr_matches_reference(golden):
    g = golden("amr")
    x32, y, w32, lev = _spread_inputs(g)
    np.random.seed(77)
    cr, vr, kr = on.spread_amr([x32, y], lev, [w32, lev], [w32, lev], nsample=7, box=50., stretch=[1., 0.5])
    for got, key in ((cr[0], "sp_x"), (cr[1], "sp_y"), (vr[0], "sp_w"), (vr[1], "sp_li"), (kr[0], "sp_cw"), (kr[1], "sp_cl")):
        assert got.dtype == g[key].dtype and np.array_equal(got, g[key]), key


def test_oracle_amr_maps_match_reference(golden):
    """spread + points_2vmaps restated: the reference's maps, bit for bit."""
    g = golden("amr")
    mat = on.rotation_matrix(np.float32(30.), np.float32(60.))
    pos = np.dot(mat, g["pos"].T).T
    vel = np.dot(mat, g["vel"].T).T
    np.random.seed(int(g["seed"]))
    (x, y), (w,), (d, _) = on.spread_amr([pos[:, 0], pos[:, 1]], g["level"], [g["mass"]], [vel[:, 2], g["level"]],
                                          nsample=10, box=100., stretch=[1., 1.])
    _, _, M, V, S = on.points_2vmaps(x, y, d, weights=w, limXY=LIM, nXY=32)
    assert np.array_equal(M, g["M"]) and np.array_equal(V, g["V"]) and np.array_equal(S, g["S"], equal_nan=True)


@pytest.mark.gpu
def test_gpu_spread_amr_bit_exact(golden):
    import torch
    from pynmap_b200 import amr
    g = golden("amr")
    x32, y, w32, lev = _spread_inputs(g)
    np.random.seed(77)
    cr, vr, kr = amr.spread_amr([x32, torch.from_numpy(y).cuda()], lev, list_val=[w32, lev], list_const=[w32, lev],
                                nsample=7, box=50., stretch=[1., 0.5])
    for got, key in ((cr[0], "sp_x"), (cr[1], "sp_y"), (vr[0], "sp_w"), (vr[1], "sp_li"), (kr[0], "sp_cw"), (kr[1], "sp_cl")):
        h = got.cpu().numpy()
        assert h.dtype == g[key].dtype and np.array_equal(h, g[key]), key


@pytest.mark.gpu
def test_gpu_amr_velocity_moments(golden, capsys):
    import pynmap_b200 as pnm
    g = golden("amr")
    snap = pnm.snapshot.from_arrays(g["pos"], g["vel"], mass=g["mass"], level=g["level"])
    snap.rotate(PA=30., inclination=60.)
    np.random.seed(int(g["seed"]))
    a = snap.amr_velocity_moments(weight_name="mass", box=100., limXY=LIM, nXY=32)
    np.random.seed(int(g["seed"]))
    b = snap.amr_velocity_moments(box=100., limXY=LIM, nXY=[24, 40])
    c = snap.amr_velocity_moments(weight_name="mass", box=100., limXY=LIM, nXY=32, spread=False)
    for got, (km, kv, ks) in ((a, ("M", "V", "S")), (b, ("Mu", "Vu", "Su")), (c, ("Mn", "Vn", "Sn"))):
        assert isinstance(got.M, np.ndarray) and got.M.shape == g[km].shape
        assert rel_err(got.M, g[km]) < 1e-12
        occupied = g[km] > 0
        assert np.max(np.abs(got.V - g[kv])[occupied]) < 1e-9 * 300.0
        var_ref = np.nan_to_num(g[ks]) ** 2
        assert np.max(np.abs(np.nan_to_num(got.S) ** 2 - var_ref)[occupied]) < 1e-9 * 300.0 ** 2
    assert (b.M.sum() - round(b.M.sum())) < 1e-9 and b.M.sum() <= 3000          # unit weights / 10 x 10 samples
    assert a.PAs == snap.PAs
    # error behaviour (pynmap.py:348-350) and the unavailable branch
    bare = pnm.snapshot.from_arrays(g["pos"], g["vel"], mass=g["mass"])
    assert bare.amr_velocity_moments() is None and "missing some input info" in capsys.readouterr().out
    with pytest.raises(NotImplementedError):
        snap.amr_velocity_moments(smear=True)


@pytest.mark.gpu
def test_gpu_amr_float32_snapshot_against_oracle(golden):
    """float32 particles: the jittered coordinates are float64 while velocities and weights / 10 stay
    float32, so the weighted terms are float32 products (numpy result types) -- checked against the
    oracle's restatement of the whole chain on the same seeded deviates."""
    import pynmap_b200 as pnm
    g = golden("amr")
    pos, vel, mass = (g[k].astype(np.float32) for k in ("pos", "vel", "mass"))
    level = g["level"].astype(np.int64)
    snap = pnm.snapshot.from_arrays(pos, vel, mass=mass, level=level)
    snap.rotate(PA=30., inclination=60.)
    np.random.seed(5)
    got = snap.amr_velocity_moments(weight_name="mass", box=100., limXY=LIM, nXY=32)
    p_rot, v_rot = on.rotate_mat(pos, vel, np.float32(30.), np.float32(60.))
    np.random.seed(5)
    (x, y), (w,), (d, _) = on.spread_amr([p_rot[:, 0], p_rot[:, 1]], level, [mass], [v_rot[:, 2], level],
                                          nsample=10, box=100., stretch=[1., 1.])
    assert x.dtype == np.float64 and d.dtype == np.float32 and w.dtype == np.float32
    _, _, M, V, S = on.points_2vmaps(x, y, d, weights=w, limXY=LIM, nXY=32)
    occupied = M > 0
    assert rel_err(got.M, M) < 1e-12 and np.max(np.abs(got.V - V)[occupied]) < 1e-10 * 300.0
    assert np.max(np.abs(np.nan_to_num(got.S) ** 2 - np.nan_to_num(S) ** 2)[occupied]) < 1e-10 * 300.0 ** 2
