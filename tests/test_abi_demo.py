"""examples/abi_demo.cpp: a compiled C++ host that uses only include/pnm_b200.h and the CUDA runtime
(no Python, no torch) runs the hot path and checks it against its own plain loop."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_demo_builds_and_links_against_the_library():
    from pynmap_b200.csrc import build
    exe = build.build_demo()
    assert os.access(exe, os.X_OK)
    needed = subprocess.run(["readelf", "-d", exe], stdout=subprocess.PIPE, text=True).stdout
    assert "libpnm_b200.so" in needed and "$ORIGIN/../pynmap_b200" in needed
    assert "torch" not in needed and "python" not in needed.lower()


@pytest.mark.gpu
def test_demo_passes_on_the_gpu():
    from pynmap_b200.csrc import build
    try:
        exe = build.build_demo()
    except RuntimeError:                       # no nvcc on this box: the prebuilt binary travelled with the snapshot
        exe = build.DEMO_BIN
        assert os.path.isfile(exe)
    proc = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert proc.returncode == 0 and proc.stdout.strip().endswith("PASS"), proc.stdout
