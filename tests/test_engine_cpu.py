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


def test_accmode_exact_flavour_and_privatisation_flag():
    a = engine.AccMode("f64-exact")
    assert a.name == "f64" and a.dtype == torch.float64 and a.privatise is False
    assert engine.AccMode("f64").privatise is True and engine.AccMode("fixed").privatise is True
    assert engine.AccMode("f64_exact").privatise is False


def test_cube_fixed_scales_cap_typical_terms():
    """engine.cube_fixed_scales: powers of two, never above the no-overflow scales, typical term in [2^24, 2^25)."""
    import math
    n, max_w, max_d = 10 ** 8, 2.0, 900.0
    mean = (1.0, 106.3, 15090.0)
    base = engine.fixed_scales(n, max_w, max_d)
    capped = engine.cube_fixed_scales(n, max_w, max_d, mean)
    for b, c, m in zip(base, capped, mean):
        assert c <= b and math.log2(c) == int(math.log2(c))
        assert 2.0 ** engine.CUBE_TYP_BITS <= m * c < 2.0 ** (engine.CUBE_TYP_BITS + 1)
        assert n * max_w * max_d ** 0 * c < 2.0 ** 63
    assert engine.cube_scales_ok(capped) and engine.cube_scales_ok(None)
    assert not engine.cube_scales_ok([2.0 ** 130, 1.0, 1.0]) and not engine.cube_scales_ok([1.0, 2.0 ** -130, 1.0])
    # huge particle counts: the no-overflow bound wins over the typical-term cap
    tight = engine.cube_fixed_scales(2 ** 40, 1.0e6, 1.0e3, (1.0, 1.0, 1.0))
    assert tight == engine.fixed_scales(2 ** 40, 1.0e6, 1.0e3)
    # degenerate means (0, inf, NaN) leave the base scales alone
    assert engine.cube_fixed_scales(n, max_w, max_d, (0.0, float("inf"), float("nan"))) == base


def test_interleaved_accumulator_padding_for_slab_reducers():
    g = types.SimpleNamespace(nx=5, ny=4, nv=7, device="cpu")
    g.record_slabs = (g.nv + 1) // 2 + 1                     # 4 pairs of velocity bins + the out-of-range slab
    g.slab_elems = g.nx * g.ny * 4
    g.cube_moments_elems = g.record_slabs * g.slab_elems
    acc = engine.AccMode("f64")
    sums = 31 | engine._lib.CUBE_MOMENTS
    maps, cube = engine.new_accumulators(g, sums, acc)
    assert maps is None and cube.numel() == 5 * g.slab_elems
    for world, slabs in ((2, 6), (3, 6), (4, 8), (8, 8)):
        _, padded = engine.new_accumulators(g, sums, acc, slab_multiple=world)
        assert padded.numel() == slabs * g.slab_elems and float(padded.abs().sum()) == 0.0
    assert engine.slab_vrange(g, 0, 2) == (0, 4) and engine.slab_vrange(g, 2, 4) == (4, 7) and engine.slab_vrange(g, 4, 6) == (7, 7)


def test_host_path_rejects_mismatched_arrays_before_touching_the_device():
    """project_host only passes pointers down, so dtype / length mismatches must be caught on the host."""
    x = np.zeros(10, dtype=np.float32)
    good = [x] * 6 + [None]
    for bad in ([x] * 5 + [np.zeros(10, dtype=np.float64), None],       # another dtype
                [x] * 5 + [np.zeros(9, dtype=np.float32), None],        # shorter
                [x] * 6 + [np.zeros((10, 1), dtype=np.float32)],        # not 1-D
                [x] * 5 + [None, None],                                 # a coordinate missing
                [x] * 6):                                               # no weights slot
        with pytest.raises(ValueError):
            hostpath.project_host(bad, 30., 60., [-1, 1, -1, 1], 8)
    with pytest.raises(ValueError):
        hostpath.project_host(good, 30., 60., [-1, 1, -1, 1], 8, accumulate="fixed")   # fixed point needs scales
