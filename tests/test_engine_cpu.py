"""CPU: engine-level host logic that needs no device -- joint accumulator layout / reduce span,
accumulate-mode parsing, smoothing taps of the host path."""
import types

import numpy as np
import pytest
import torch

from oracle import oracle_np as on
from pynmap_b200 import engine, hostpath


def fake_grid(nx, ny, nv, replicas):
    g = types.SimpleNamespace(nx=nx, ny=ny, nv=nv, device="cpu")
    g.cube_elems = nx * ny * nv
    g.maps_fold_elems = nx * ny * 4
    g.maps_elems = g.maps_fold_elems * replicas
    return g


@pytest.mark.parametrize("shape", [(4, 3, 5), (3, 3, 3), (8, 8, 0)])
def test_joint_accumulators_and_reduce_span(shape):
    nx, ny, nv = shape
    g = fake_grid(nx, ny, nv, replicas=4)
    acc = engine.AccMode("fixed")
    sums = 7 | (8 if nv else 0)
    maps, cube = engine.new_accumulators(g, sums, acc)
    assert maps.numel() == g.maps_elems and maps.dtype == torch.int64 and int(maps.abs().sum()) == 0
    if nv:
        assert cube.numel() == g.cube_elems
        assert maps.data_ptr() % 16 == 0 or cube.data_ptr() % 16 != 0      # maps stay 16-byte aligned after the cube
        span = engine.reduce_span(g, maps, cube)
        assert len(span) == 1                                               # one contiguous view -> one NCCL call
        span[0].add_(1)                                                     # touches cube + first replica only
        assert int(cube.sum()) == g.cube_elems and int(maps[:g.maps_fold_elems].sum()) == g.maps_fold_elems
        assert int(maps[g.maps_fold_elems:].sum()) == 0
    else:
        assert cube is None
        span = engine.reduce_span(g, maps, None)
        assert len(span) == 1 and span[0].numel() == g.maps_fold_elems
    # separately allocated tensors fall back to two pieces
    m2, c2 = torch.zeros(g.maps_elems, dtype=torch.int64), torch.zeros(max(g.cube_elems, 1), dtype=torch.int64)
    assert [t.numel() for t in engine.reduce_span(g, m2, c2)] == [g.maps_fold_elems, max(g.cube_elems, 1)]
    # several views: plain per-view buffers
    mv, cv = engine.new_accumulators(g, sums, acc, nviews=3)
    assert mv.numel() == 3 * g.maps_elems and (cv is None or cv.numel() == 3 * g.cube_elems)


def test_accmode_names():
    assert engine.AccMode("f64").dtype == torch.float64 and engine.AccMode("float64").name == "f64"
    assert engine.AccMode("fixed").dtype == torch.int64 and engine.AccMode("int64").name == "fixed"
    with pytest.raises(ValueError):
        engine.AccMode("f32")


def test_host_path_taps_follow_losvd_axis_quirk():
    """hostpath.smoothing_taps == the factors of the oracle's 2-D kernel, with fwhmx acting along
    cube axis 1 (losvd.py:127-131: the kernel array is [y, x], slice axis 0 is X)."""
    stepx, stepy = 30 / 127, 30 / 95
    t0, t1 = hostpath.smoothing_taps(1.2, 2.4, stepx, stepy)
    k2 = on.gaussian_kernel2d(1.2 * on.FWHM_TO_SIGMA / stepx, 2.4 * on.FWHM_TO_SIGMA / stepy)
    assert np.allclose(np.outer(t0, t1), k2, rtol=1e-13, atol=1e-300)
    assert len(t0) == len(t1) == k2.shape[0] and len(t0) % 2 == 1
