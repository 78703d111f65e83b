"""The C-ABI library loads and exports every symbol include/pulser_diff_b200.h declares.

No compute calls here (no GPU in the CPU tier); the gpu tier exercises them.
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pulser_diff_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from pulser_diff_b200 import _cabi
    assert set(declared_symbols()) == set(_cabi.EXPORTS)


def test_cuda_library_exports_every_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(g.LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.pd_abi_version.restype = ctypes.c_int
    assert lib.pd_abi_version() == 1
    assert lib.pd_is_cuda() == 1
    assert lib.pd_amplitude_bytes() == 16
    # the complex64 build of the same ABI (north_star's optional tier): same symbols, 8-byte amplitudes
    lib64 = ctypes.CDLL(g.LIB_C64)
    for name in declared_symbols():
        assert hasattr(lib64, name), name
    assert lib64.pd_abi_version() == 1 and lib64.pd_is_cuda() == 1 and lib64.pd_amplitude_bytes() == 8


def test_product_refuses_cpu(emu_library):
    """No CPU fallback: the product library raises instead of computing on the host."""
    import torch
    from pulser_diff_b200 import _cabi, ops
    ops.clear_plan_cache()
    _cabi.use_library(None)
    with pytest.raises(RuntimeError):
        _cabi.Plan(2, 1, _cabi.PD_KET, torch.device("cpu"))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            _cabi.Plan(2, 1, _cabi.PD_KET, torch.device("cuda", 0))
