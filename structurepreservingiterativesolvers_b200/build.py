"""Build the sm_100a shared library in-tree with nvcc (no torch, no JIT cache).

    python -m structurepreservingiterativesolvers_b200.build [--force] [--verbose]

The library is written to structurepreservingiterativesolvers_b200/lib/libspis_b200.so so that it
travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libspis_b200.so")
SOURCES = [os.path.join(CSRC, "spis_capi.cu")]
HEADERS = [os.path.join(CSRC, "spis_kernels.cuh"),
           os.path.join(os.path.dirname(PKG_DIR), "include", "spis_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-DSPIS_BUILD",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the sm_100a library cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a; returns the path of the shared library."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + SOURCES + ["-o", LIB_PATH + ".tmp"]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose or ptxas_info:
        sys.stdout.write(res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or True,
                 ptxas_info="--ptxas" in sys.argv)
    print("built", path)
