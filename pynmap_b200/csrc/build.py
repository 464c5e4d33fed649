"""Build libpnm_b200.so (sm_100a only) in-tree with nvcc.

Run as ``python -m pynmap_b200.csrc.build`` or through ``__graft_entry__.build()``.
nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libpnm_b200.so")
SOURCES = ["pnm_api.cu", "pnm_host.cu"]
HEADERS = ["pnm_common.cuh", "pnm_bin.cuh", "pnm_post.cuh", "pnm_fit.cuh", "pnm_profile.cuh", "pnm_cube.cuh", "pnm_plan.cuh",
           os.path.join("..", "..", "include", "pnm_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--diag-suppress", "177",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libpnm_b200.so cannot be built (there is no CPU fallback)")


def is_stale():
    if not os.path.isfile(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.isfile(d) and os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources; returns its path."""
    if not force and not is_stale():
        return LIB
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.isfile(os.path.join(HERE, s))]
    extra = os.environ.get("PNM_NVCC_EXTRA", "").split()
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB + ".tmp"] + srcs
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stdout))
    if verbose:
        print(proc.stdout)
    os.replace(LIB + ".tmp", LIB)
    return LIB


DEMO_SRC = os.path.join(os.path.dirname(PKG), "examples", "abi_demo.cpp")
DEMO_BIN = os.path.join(os.path.dirname(PKG), "examples", "abi_demo")


def build_demo(force=False):
    """examples/abi_demo: a compiled host on the C ABI alone (no Python, no torch), linked against the
    in-tree library with an $ORIGIN-relative rpath so that it runs from the repository snapshot."""
    build()
    if not force and os.path.isfile(DEMO_BIN) and os.path.getmtime(DEMO_BIN) > max(os.path.getmtime(DEMO_SRC), os.path.getmtime(LIB)):
        return DEMO_BIN
    cmd = [find_nvcc(), "-std=c++17", "-O2", "-Wno-deprecated-gpu-targets", "-I", os.path.join(os.path.dirname(PKG), "include"),
           DEMO_SRC, "-o", DEMO_BIN, "-L", PKG, "-lpnm_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../pynmap_b200"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stdout))
    return DEMO_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
