"""pulser_diff_b200: the evolution hot path of pasqal-io/pulser-diff on NVIDIA B200.

Drop-in for ``pyqtorch.sesolve`` / ``mesolve`` behind ``TorchEmulator.run`` (reference
pulser_diff/backend.py:485-529): hand-written sm_100a CUDA kernels reached through the C ABI
of ``include/pulser_diff_b200.h``.  There is no CPU fallback.
"""
from .backend import DeviceSpec, Level60Device, MockDevice, TorchEmulator
from .derivative import deriv_param, deriv_time
from .hamiltonian import Hamiltonian, StructuredHamiltonian
from .samples import ChannelSamples, PulseBuilder, SequenceSamples
from .simconfig import SimConfig
from .simresults import CoherentResults
from .solvers import Result, SolverType, mesolve, sesolve

__all__ = [
    "TorchEmulator", "SimConfig", "SolverType", "sesolve", "mesolve", "Result", "Hamiltonian",
    "StructuredHamiltonian", "SequenceSamples", "ChannelSamples", "PulseBuilder", "DeviceSpec",
    "MockDevice", "Level60Device", "CoherentResults", "deriv_time", "deriv_param",
]
