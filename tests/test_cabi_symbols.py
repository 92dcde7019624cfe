"""CPU tier: libb2048.so builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every
symbol include/b2048.h declares; the ctypes table mirrors the header; no compute call is made."""
import ctypes as C
import importlib
import os
import re
import sys

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, os.path.join(ROOT, "2048_b200"))
    spec = importlib.util.spec_from_file_location("b2048_build", os.path.join(ROOT, "2048_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def header_functions():
    text = open(os.path.join(ROOT, "include", "b2048.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2048_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(built)
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b2048.h but not exported"
    lib.b2048_abi_version.restype = C.c_int
    assert lib.b2048_abi_version() == 1


def test_ctypes_table_mirrors_header(built):
    from game2048 import cabi
    assert sorted(cabi.SIGNATURES) == header_functions()
    L = cabi.lib()
    # host-only layout queries need no GPU
    assert [cabi.num_feat(n) for n in (2, 3, 4, 5, 6)] == [24, 52, 17, 21, 33]
    assert [cabi.num_weights(n) for n in (2, 3, 4, 5, 6)] == [6144, 212992, 1114112, 5308416, 95662848]
    assert L.b2048_num_feat(9) == -1
    assert cabi.table_offsets(5)[17:] == [17 * 65536 + k * 1048576 for k in range(5)]
    assert C.sizeof(cabi.Games) == 14 * 8 and C.sizeof(cabi.Replay) == 24


def test_product_fails_loudly_without_gpu(built):
    """no CPU fallback: without a CUDA device the product refuses to construct a context"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from game2048 import engine
    with pytest.raises(engine.B2048Error):
        engine.Context.get()
