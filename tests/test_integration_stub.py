"""The ctypes stub INTEGRATION.md shows to a pynmap maintainer is executed as written (only the
library path is substituted) and must reproduce the oracle -- the document cannot rot."""
import os
import re

import numpy as np
import pytest

from oracle import oracle_np as on

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if "def vmaps_b200" in b]
    assert len(stub) == 1
    return stub[0]


def test_stub_binds_only_declared_symbols():
    from pynmap_b200 import _lib
    used = set(re.findall(r"lib\.(pnm_[a-z0-9_]+)", stub_source()))
    assert used and used <= set(_lib.EXPORTED_SYMBOLS), used - set(_lib.EXPORTED_SYMBOLS)


@pytest.mark.gpu
def test_stub_runs_and_matches_the_oracle():
    import torch
    from pynmap_b200 import _lib
    src = stub_source().replace('ctypes.CDLL("libpnm_b200.so")', "ctypes.CDLL(%r)" % _lib.LIB_PATH)
    ns = {}
    exec(compile(src, "INTEGRATION.md:_b200.py", "exec"), ns)
    rng = np.random.default_rng(2)
    n = 50000
    x, y = rng.normal(0, 3, n).astype(np.float32), rng.normal(0, 3, n).astype(np.float32)
    v, w = rng.normal(0, 100, n).astype(np.float32), rng.uniform(0.5, 2, n).astype(np.float32)
    lim = [-6.0, 6.0, -5.0, 7.0]
    M, V, S = ns["vmaps_b200"](*(torch.from_numpy(a).cuda() for a in (x, y, v, w)), lim, 24, 18)
    _, _, Mo, Vo, So = on.points_2vmaps(x, y, v, weights=w, limXY=lim, nXY=[24, 18])
    assert M.shape == (18, 24)
    assert np.allclose(M.cpu().numpy(), Mo, rtol=1e-12, atol=0)
    ok = Mo > 0
    assert np.max(np.abs(V.cpu().numpy() - Vo)[ok]) < 1e-9 * 100
