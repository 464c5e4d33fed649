"""Radial velocity-ellipsoid profiles -- grid.point_2vprofiles (pynmap/grid.py:16-68) and
snapshot.get_tensor_profiles (pynmap/pynmap.py:457-474).  tests/golden/profiles.npz: one float32
particle set and the reference's outputs on it and on its float64 upcast.  Bars: radii and per-bin
membership bit-exact; means and dispersions (float32 in the reference, computed there with float32
pairwise sums and libm cos(arctan2)) within 2e-5 of the bin's velocity scale."""
import numpy as np
import pytest

from oracle import oracle_np as on

TOL = 2e-5


def _cols(g, dt):
    pos, vel = g["pos"].astype(dt), g["vel"].astype(dt)
    return [pos[:, 0], pos[:, 1], pos[:, 2], vel[:, 0], vel[:, 1], vel[:, 2]]


def test_oracle_profiles_match_reference(golden):
    g = golden("profiles")
    for dt in (np.float32, np.float64):
        tag = np.dtype(dt).name
        Rbin, v, s, counts = on.point_2vprofiles(*_cols(g, dt), nbins=40)
        assert np.array_equal(Rbin, g["Rbin_" + tag])
        assert np.array_equal(v, g["v_" + tag], equal_nan=True) and np.array_equal(s, g["s_" + tag], equal_nan=True)
        assert v.dtype == np.float32 and counts.sum() > 10000 and (counts == 0).any()
        Rb2, v2, s2, _ = on.point_2vprofiles(*_cols(g, dt), nbins=25, dz=(-0.5, 0.1), Rmin=0.05, Rmax=20.0)
        assert np.array_equal(Rb2, g["Rbin2_" + tag]) and np.array_equal(s2, g["s2_" + tag], equal_nan=True)


def _close(v, s, v_ref, s_ref):
    assert np.array_equal(np.isnan(v), np.isnan(v_ref)) and np.array_equal(np.isnan(s), np.isnan(s_ref))
    scale = np.nan_to_num(np.abs(v_ref)) + np.nan_to_num(s_ref) + 1e-3
    worst_v = np.nanmax(np.abs(v - v_ref) / scale)
    worst_s = np.nanmax(np.abs(s - s_ref) / scale)
    assert worst_v < TOL and worst_s < TOL, (worst_v, worst_s)


@pytest.mark.gpu
def test_gpu_profiles_match_reference(golden):
    import torch
    import pynmap_b200 as pnm
    from pynmap_b200 import _lib, engine
    g = golden("profiles")
    for dt in (np.float32, np.float64):
        tag = np.dtype(dt).name
        cols = _cols(g, dt)
        Rbin, v, s = pnm.grid.point_2vprofiles(*cols, nbins=40)
        assert np.array_equal(Rbin, g["Rbin_" + tag]) and v.dtype == np.float32 and v.shape == (3, 40)
        _close(v, s, g["v_" + tag], g["s_" + tag])
        Rb2, v2, s2 = pnm.grid.point_2vprofiles(*cols, nbins=25, dz=np.array([-0.5, 0.1]), Rmin=0.05, Rmax=20.0)
        assert np.array_equal(Rb2, g["Rbin2_" + tag])
        _close(v2, s2, g["v2_" + tag], g["s2_" + tag])
        # per-bin membership through the C ABI: exact
        _, _, _, counts_ref = on.point_2vprofiles(*cols, nbins=40)
        tdt = torch.float32 if dt == np.float32 else torch.float64
        dev = [torch.from_numpy(np.ascontiguousarray(c)).cuda() for c in cols]
        rb = torch.from_numpy(Rbin).cuda()
        acc = torch.empty(int(_lib.load().pnm_profile_acc_elems(40)), dtype=torch.float64, device="cuda")
        vv, ss = (torch.empty((3, 40), dtype=torch.float32, device="cuda") for _ in range(2))
        cnt = torch.empty(40, dtype=torch.int64, device="cuda")
        P = engine.ptr
        _lib.call("pnm_radial_profiles", *[P(t) for t in dev], None, dev[0].numel(), _lib.F32 if tdt == torch.float32 else _lib.F64,
                  -0.2, 0.2, P(rb), 40, P(acc), P(vv), P(ss), P(cnt), engine.stream_ptr())
        assert np.array_equal(cnt.cpu().numpy(), counts_ref)
        assert np.array_equal(vv.cpu().numpy(), v, equal_nan=True)          # deterministic membership, same sums to rounding
    # CUDA tensors in -> CUDA tensors out
    dev = [torch.from_numpy(np.ascontiguousarray(c)).cuda() for c in _cols(g, np.float32)]
    _, vt, st = pnm.grid.point_2vprofiles(*dev, nbins=40)
    assert vt.is_cuda and st.dtype == torch.float32


@pytest.mark.gpu
def test_gpu_get_tensor_profiles_and_mask(golden):
    import pynmap_b200 as pnm
    g = golden("profiles")
    pos, vel = g["pos"], g["vel"]
    snap = pnm.snapshot.from_arrays(pos, vel, mass=np.ones(len(pos), dtype=np.float32))
    Rbin, v, s = snap.get_tensor_profiles(nbins=40)
    assert np.array_equal(Rbin, g["Rbin_float32"])
    _close(v, s, g["v_float32"], g["s_float32"])
    mask = pos[:, 0] > 0
    Rm, vm, sm = snap.get_tensor_profiles(nbins=30, mask=mask)
    cols = [pos[mask, 0], pos[mask, 1], pos[mask, 2], vel[mask, 0], vel[mask, 1], vel[mask, 2]]
    Ro, vo, so, _ = on.point_2vprofiles(*cols, nbins=30)
    assert np.array_equal(Rm, Ro)
    _close(vm, sm, vo, so)
    with pytest.raises(NameError):
        snap.get_tensor_profiles(plot=True)
