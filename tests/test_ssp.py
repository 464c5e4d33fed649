"""f2 -- SSP mass-to-light weights (pynmap/misc_functions.py:115-207).

tests/golden/ssp.npz holds the (age, [M/H]) -> M/L nodes the reference interpolates (built from its
E-MILES tables, KB / BaSTI / slope 1.30 / V band), 4000 query points and the reference's own output
of compute_ml_ssps.  CPU: the oracle (scipy griddata, like the reference) and the host-side
Clough-Tocher tables reproduce it.  GPU: the pnm_ct_interp kernel reproduces it and scipy on 10^6
points, in and outside the convex hull."""
import numpy as np
import pytest

from oracle import oracle_np as on


def test_oracle_reproduces_reference(golden):
    g = golden("ssp")
    assert g["node_ML"].size == 636 and 0.06 < g["node_ML"].min() < g["node_ML"].max() < 8.7
    ML = on.ml_from_nodes(g["node_age"], g["node_MH"], g["node_ML"], g["age"], g["MH"])
    assert np.array_equal(ML, g["ML"])                       # same scipy, same calls: bit-exact
    nofill = on.ml_from_nodes(g["node_age"], g["node_MH"], g["node_ML"], g["age"], g["MH"], fill_nan=False)
    assert np.array_equal(np.isnan(nofill), np.isnan(g["ML_nofill"])) and 0.05 < np.isnan(nofill).mean() < 0.2


def test_host_tables_reproduce_reference(golden):
    """The table-only precomputation (Delaunay, gradients, 19 Bezier ordinates per triangle, cell
    index) evaluated in numpy == the reference's output to rounding."""
    from pynmap_b200 import ssp
    g = golden("ssp")
    tabs = ssp.clough_tocher_tables(np.stack([g["node_age"], g["node_MH"]], 1), g["node_ML"])
    cells = ssp.cell_index(tabs["points"], tabs["simplices"], 64)
    assert tabs["coef"].shape[1] == 19 and tabs["xform"].shape[1] == 6 and cells["start"][-1] == cells["tris"].size
    out = ssp.eval_tables_numpy(tabs, cells, g["age"][:1500], g["MH"][:1500])
    assert np.allclose(out, g["ML"][:1500], rtol=1e-13, atol=0)
    nofill = ssp.eval_tables_numpy(tabs, cells, g["age"][:1500], g["MH"][:1500], fill_nearest=False)
    assert np.array_equal(np.isnan(nofill), np.isnan(g["ML_nofill"][:1500]))


@pytest.mark.gpu
def test_gpu_interpolator_vs_reference_and_scipy(golden):
    import torch
    from pynmap_b200 import ssp
    g = golden("ssp")
    interp = ssp.MLInterpolator(g["node_age"], g["node_MH"], g["node_ML"])
    ML = interp(g["age"], g["MH"])
    assert isinstance(ML, np.ndarray) and ML.dtype == np.float64
    assert np.allclose(ML, g["ML"], rtol=1e-12, atol=0)                       # the reference's own output
    nofill = interp(g["age"], g["MH"], fill_nan=False)
    assert np.array_equal(np.isnan(nofill), np.isnan(g["ML_nofill"]))
    # 10^6 random points (float32 tensors in, CUDA tensor out) against scipy
    gen = torch.Generator(device="cuda").manual_seed(3)
    age = 14.5 * torch.rand(1_000_000, generator=gen, device="cuda")
    mh = -2.4 + 2.9 * torch.rand(1_000_000, generator=gen, device="cuda")
    out = interp(age, mh)
    assert out.is_cuda and out.dtype == torch.float64
    ref = on.ml_from_nodes(g["node_age"], g["node_MH"], g["node_ML"], age.cpu().numpy(), mh.cpu().numpy())
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-12, atol=0)
    # table nodes themselves are reproduced
    assert np.allclose(interp(g["node_age"], g["node_MH"]), g["node_ML"], rtol=1e-12)


@pytest.mark.gpu
def test_snapshot_flux_from_ssp(golden):
    """snapshot(ml_function=...) -> flux = mass / ML (pynmap.py:234-236) -> flux-weighted map."""
    import torch
    import pynmap_b200 as pnm
    import workloads
    from pynmap_b200 import ssp
    g = golden("ssp")
    interp = ssp.MLInterpolator(g["node_age"], g["node_MH"], g["node_ML"])
    n = 200_000
    soa = workloads.make_snapshot(n, seed=77, device="cuda", unit_mass=False)
    gen = torch.Generator(device="cuda").manual_seed(4)
    age = 0.03 + 13.97 * torch.rand(n, generator=gen, device="cuda")
    mh = -2.27 + 2.67 * torch.rand(n, generator=gen, device="cuda")
    snap = pnm.snapshot.from_arrays(soa[:3], soa[3:6], mass=soa[6], age=age, MH=mh, output="numpy",
                                    ml_function=lambda mass, a, m: interp(a, m))
    ML = on.ml_from_nodes(g["node_age"], g["node_MH"], g["node_ML"], age.cpu().numpy(), mh.cpu().numpy())
    flux = soa[6].cpu().numpy().astype(np.float64) / ML
    assert np.allclose(snap.flux.cpu().numpy(), flux.astype(np.float32), rtol=2e-7)
    snap.rotate(30., 60.)
    fm = snap.create_map(weight_name="lum", limXY=workloads.LIM_XY, nXY=64)
    assert fm.W.sum() > 0 and abs(fm.W.sum() / snap.flux.double().sum().item() - 1) < 0.05
