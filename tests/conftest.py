import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libpd_emu.so")
EMU_LIB_C64 = os.path.join(EMU_DIR, "libpd_emu_c64.so")     # same stand-in compiled with -DPD_C64


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_library():
    """Host stand-in of the C ABI (tests/emu): exercises host logic without a GPU."""
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
    return EMU_LIB


@pytest.fixture(params=[pytest.param("emu", id="emu"),
                        pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def engine_device(request, emu_library):
    """Binds the library under test and yields the torch device states live on.

    "emu"  -> tests/emu/libpd_emu.so on CPU tensors (host logic only; CPU test tier)
    "cuda" -> the product library on cuda:0 (the parity tests proper; GPU tier)
    """
    from pulser_diff_b200 import _cabi, ops
    ops.clear_plan_cache()
    if request.param == "emu":
        _cabi.use_library(emu_library, EMU_LIB_C64)
        dev = torch.device("cpu")
    else:
        if not torch.cuda.is_available():
            pytest.fail("gpu-marked test selected but no CUDA device is visible")
        _cabi.use_library(None)
        dev = torch.device("cuda", 0)
    yield dev
    ops.clear_plan_cache()
    _cabi.use_library(None)


@pytest.fixture
def cuda_device():
    from pulser_diff_b200 import _cabi, ops
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test selected but no CUDA device is visible")
    ops.clear_plan_cache()
    _cabi.use_library(None)
    yield torch.device("cuda", 0)
    ops.clear_plan_cache()
