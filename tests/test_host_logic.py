"""CPU: host-side logic of the product (edges, matrices, scales, taps, shapes) against the golden
vectors and the oracle -- no device needed."""
import math

import numpy as np
import pytest

from oracle import oracle_np as on
from pynmap_b200 import engine, rotation
from pynmap_b200.pynmap import convert_to_list


def test_matrices_match_reference(golden):
    g = golden("kat")
    assert np.array_equal(rotation.rotation_matrix(30.0, 60.0), g["mat_f64ang"])
    pa, inc = convert_to_list(30.0)[0], convert_to_list(60.0)[0]
    assert isinstance(pa, np.float32)
    assert np.array_equal(rotation.rotation_matrix(pa, inc), g["mat_f32ang"])
    assert np.array_equal(rotation.rotation_matrix(30.0, 60.0, direct=False), g["mat_f64ang"].T)
    assert np.array_equal(rotation.rotation_matrix_xyz(10., 20., 30., [2, 0, 1]), on.rotation_matrix_xyz(10., 20., 30., [2, 0, 1]))
    with pytest.raises(ValueError):            # SURVEY 8c KAT 6: a list of angles does not work in the reference either
        rotation.rotation_matrix(convert_to_list([30, 40])[0], convert_to_list([60, 70])[0])


def test_edges_and_shapes():
    for lo, hi, n in ((-15, 15, 256), (-1000, 1000, 100), (-1, 1, 3), (-12.5, 7.25, 511)):
        e = engine.bin_edges(lo, hi, n)
        half = (hi - lo) / (n - 1) / 2.
        assert np.array_equal(e, np.linspace(lo - half, hi + half, n + 1))   # grid.py:119-124 verbatim
        assert np.array_equal(e, on.bin_edges(lo, hi, n))
    assert engine.expand_nxy(7) == (7, 7) and engine.expand_nxy((64, 32)) == (64, 32)
    assert engine.expand_nxy(np.array([5])) == (5, 5) and engine.expand_nxy((1, 2, 3)) is None


def test_fixed_scales_cannot_overflow():
    for n, mw, md in ((10 ** 9, 3.0, 1200.0), (10 ** 6, 1e-3, 5.0), (1, 1.0, 1.0), (10 ** 8, 1e6, 1e4)):
        s = engine.fixed_scales(n, mw, md)
        for scale, bound in zip(s, (mw, mw * md, mw * md * md)):
            assert math.frexp(scale)[0] == 0.5                 # a power of two
            assert n * bound * scale <= 2.0 ** 62
            assert n * bound * scale > 2.0 ** 60


def test_gaussian_taps_match_astropy_statement():
    for sx, sy in ((1.8, 1.8), (1.2, 2.7), (0.1, 0.1), (6.0, 2.0)):
        s = engine.gaussian_support(sx, sy)
        assert s % 2 == 1 and s == on.kernel_support(sx, sy)
        t0, t1 = engine.gaussian_taps(sy, s), engine.gaussian_taps(sx, s)
        k2 = on.gaussian_kernel2d(sx, sy)
        assert np.allclose(np.outer(t0, t1), k2, rtol=1e-13, atol=1e-300)
    assert engine.gaussian_support(0.1, 0.1) == 1
    assert abs(engine.FWHM_TO_SIGMA - on.FWHM_TO_SIGMA) == 0


def test_result_dtype_follows_numpy():
    import torch
    f32, f64, i64 = np.zeros(2, np.float32), np.zeros(2, np.float64), np.zeros(2, np.int64)
    assert engine.result_dtype(f32) == torch.float32 and engine.result_dtype(f32, f32) == torch.float32
    assert engine.result_dtype(f32, f64) == torch.float64 and engine.result_dtype(i64) == torch.float64
    assert engine.result_dtype(torch.zeros(2), None) == torch.float32
    assert engine.result_dtype(torch.zeros(2, dtype=torch.float64)) == torch.float64


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm) prints ONE JSON line with the contract's keys; a small
    sample keeps this a few seconds.  Ranks other than 0 print nothing and exit 0."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PNM_CPU_SAMPLE="100000")
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["unit"] == "particles/s" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "workload" in j["config"]
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    other = subprocess.run(cmd, env=dict(env, RANK="1", WORLD_SIZE="2"), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True, timeout=600)
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_fitgh_parinfo_defaults_and_overrides():
    """fit_losvd._parinfo: the start / limits / fixed arrays of fit_losvd.py:157-176 (NaN start = 'from the moments')."""
    from pynmap_b200 import fit_losvd
    x = np.linspace(-490.0, 490.0, 50)
    info = fit_losvd._parinfo(x, 4)
    assert np.isnan(info["start"][:2]).all() and list(info["start"][2:]) == [0.02, 0.02]
    assert list(info["lo"]) == [-490.0, (x[1] - x[0]) / 3.0, -0.2, -0.2] or np.allclose(info["lo"], [-490.0, 20.0 / 3.0, -0.2, -0.2])
    assert np.allclose(info["hi"], [490.0, 980.0 / 3.0, 0.2, 0.2]) and not info["fixed"].any()
    part = fit_losvd._parinfo(x, 4, params=[10.0], fixed=[False, True], limitedmin=[True, False], maxpars=[100.0])
    assert part["start"][0] == 10.0 and np.isnan(part["start"][1]) and list(part["start"][2:]) == [0.02, 0.02]
    assert list(part["fixed"]) == [False, True, False, False]
    assert part["lo"][0] == -490.0 and part["lo"][1] == -np.inf and part["lo"][2] == -0.2          # limitedmin[1] = False
    assert part["hi"][0] == 100.0 and np.isclose(part["hi"][1], 980.0 / 3.0)
    assert fit_losvd._set_ghparameters([1, 2, 3, 4, 5], [9] * 4, 4) == [1, 2, 3, 4]                 # longer lists are cut


def test_shard_bounds_and_span_helpers():
    from pynmap_b200 import sharded
    for n, world in ((100, 3), (1_000_003, 8), (7, 8), (0, 2)):
        cuts = [sharded.shard_bounds(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:])) and all(lo <= hi for lo, hi in cuts)
        assert all(lo % sharded.ALIGN == 0 for lo, _ in cuts[1:] if lo < n)


def test_dropin_pynmap_import_name():
    """``dropin/`` on PYTHONPATH makes ``import pynmap`` (and its five submodules, pynmap/__init__.py:61) resolve to
    pynmap_b200 -- checked in a child process so that this process can still import the real reference."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import pynmap, pynmap_b200\n"
            "from pynmap.misc_functions import gausshermite, moment012, compute_ml, compute_ml_ssps\n"
            "from pynmap.misc_io import convert_to_list, spread_amr, rebin_1darray, rebin_2darray, read_ascii_files\n"
            "from pynmap import pynmap as p, grid, losvd, misc_io, misc_functions\n"
            "assert p.snapshot is pynmap_b200.snapshot and losvd.LOSVD is pynmap_b200.LOSVD\n"
            "assert grid.points_2vmaps is pynmap_b200.grid.points_2vmaps\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "dropin"), root]))
    out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout


def test_misc_helpers_match_their_definitions():
    """misc_io.rebin_1darray / rebin_2darray and misc_functions.moment012 (small host helpers of the drop-in surface)."""
    import numpy as np
    from pynmap_b200 import misc_functions, misc_io
    a = np.arange(12.0)
    assert np.array_equal(misc_io.rebin_1darray(a, 4), a.reshape(4, 3).sum(-1))
    assert np.array_equal(misc_io.rebin_1darray(a, 3, 'mean'), a.reshape(3, 4).mean(-1))
    b = np.arange(48.0).reshape(6, 8)
    assert np.array_equal(misc_io.rebin_2darray(b, (3, 2)), b.reshape(3, 2, 2, 4).sum(-1).sum(1))
    assert np.array_equal(misc_io.rebin_2darray(b, (2, 4), 'mean'), b.reshape(2, 3, 4, 2).mean(-1).mean(1))
    x = np.linspace(-300., 300., 31)
    d = np.exp(-0.5 * ((x - 40.) / 60.) ** 2)
    m0, m1, m2 = misc_functions.moment012(x, d)
    assert np.isclose(m0, d.sum()) and np.isclose(m1, (x * d).sum() / d.sum())
    assert np.isclose(m2, np.sqrt((x * x * d).sum() / d.sum() - m1 ** 2))
    assert misc_functions.moment012(x, np.zeros_like(x)) == (0., 0., 20. / 3.0)
    assert misc_io.return_dummy(3, 1.5) == [1.5, 1.5, 1.5]
    with __import__("pytest").raises(NotImplementedError):
        misc_io.gas_to_cube(None, None, None, None, None)


def test_low_limb_carry_scheme_of_the_privatised_kernel():
    """The arithmetic k_project_cube (csrc/pnm_cube.cuh) uses for its shared-memory sums, restated with numpy: a signed
    32-bit term a is added to an unsigned 32-bit low limb (ATOMS.ADD returns the old value); the adding thread owes the
    high part h = (a >> 31) + carry_out(old + a), which is -1, 0 or +1 and is sent to the global accumulator as
    h * 2^32.  Low limb + 2^32 * sum(h) must equal the exact 64-bit sum for any order of the terms."""
    import numpy as np
    rng = np.random.default_rng(3)
    for scale in (2 ** 24, 2 ** 30, 2 ** 31 - 1):
        a = rng.integers(-scale, scale, 20000, dtype=np.int64)
        a[:10] = [scale - 1, -(scale - 1), 0, 1, -1, scale - 1, scale - 1, -(scale - 1), -(scale - 1), 5]
        for order in (np.arange(a.size), rng.permutation(a.size)):
            low, high = 0, 0
            for t in a[order].tolist():
                old = low
                low = (old + t) & 0xFFFFFFFF                           # ATOMS.ADD on the low limb (two's complement)
                carry = 1 if old + (t & 0xFFFFFFFF) > 0xFFFFFFFF else 0   # add.cc.u32 of old and the term's low word
                h = (-1 if t < 0 else 0) + carry                       # addc.s32 with the sign extension of the term
                assert h in (-1, 0, 1)
                high += h
            assert low + (high << 32) == int(a.sum())
    # float64 accumulators: h * 2^32 * 2^-k is built from the high word of the double 2^(32-k) and the sign of h
    import struct
    for k in (-20, 0, 11, 24, 40):
        hi = struct.unpack("<q", struct.pack("<d", 4294967296.0 * 2.0 ** -k))[0] >> 32
        for h in (1, -1):
            word = (hi ^ ((h & 0xFFFFFFFF) & 0x80000000)) << 32
            assert struct.unpack("<d", struct.pack("<q", word if word < 2 ** 63 else word - 2 ** 64))[0] == h * 2.0 ** (32 - k)
