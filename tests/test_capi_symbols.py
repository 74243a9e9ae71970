"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol of include/spis_b200.h."""
import ctypes
import os
import re

import pytest

from structurepreservingiterativesolvers_b200 import _native as nat
from structurepreservingiterativesolvers_b200 import build as builder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spis_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spis_[a-z0-9_]+)\s*\(", text, flags=re.I)) - {"spis_allreduce_fn", "spis_halo_fn"})


@pytest.fixture(scope="module")
def lib_path():
    return builder.build()          # no-op when up to date; nvcc cross-compiles without a GPU


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(nat.SIGNATURES)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} missing from {lib_path}"
    assert lib.spis_abi_version() == nat.ABI_VERSION


def test_library_is_sm100a_only(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(nat.NativeLibraryError):
        nat.load_library(str(tmp_path / "nope.so"))


def test_no_cpu_fallback_without_device(lib_path):
    """Without a GPU, creating a context must fail with a CUDA error -- never fall back."""
    from structurepreservingiterativesolvers_b200.device import KrylovContext
    try:
        ndev = nat.device_count()
    except nat.SpisError:
        ndev = 0
    if ndev > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nat.SpisError):
        KrylovContext(16, 4)
