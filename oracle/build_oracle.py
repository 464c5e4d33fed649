"""TEST INFRASTRUCTURE ONLY -- builds oracle/_build/libpnm_oracle.so from oracle/pnm_oracle.c.

The reference is pure Python (no compilable sources), so there is no ``oracle/_ref`` binary:
the reference itself is exercised through ``oracle/refshim.py`` in the build container and its
outputs are frozen in ``tests/golden``.  This C file is the fast restatement used as checker
and as the ``cpu_baseline`` "port".
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pnm_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libpnm_oracle.so")


def build(force=False):
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
           "-o", LIB + ".tmp", SRC, "-lm"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed:\n%s\n%s" % (" ".join(cmd), proc.stdout))
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
