"""CPU tier: the C-ABI library is built in-tree, loads without a GPU, and exports exactly the entry points that
include/aletsch_gpu.h declares (no compute call is made here); the product refuses to run without a device."""
import ctypes as C
import os
import re

import pytest

from aletsch_b200 import gpu as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    h = open(os.path.join(ROOT, "include", "aletsch_gpu.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(agpu_\w+)\s*\(", h)))


def test_header_and_binding_agree():
    assert declared() == sorted(G.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(G.GPU_SO):
        import __graft_entry__ as ge
        ge.build_gpu()
    lib = C.CDLL(G.GPU_SO)
    for name in declared():
        assert getattr(lib, name) is not None, name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(G.AgpuError):
        G.Context(0)          # the real library: agpu_create must fail without a device
    with pytest.raises(RuntimeError):
        G.load(os.path.join(ROOT, "aletsch_b200", "does_not_exist.so"))


def test_params_struct_matches_header_defaults():
    L = G.load()
    p = G.Params()
    L.agpu_default_params(C.byref(p))
    q = G.default_params()
    for name, _ in G.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
