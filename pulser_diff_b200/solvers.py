"""Drop-in ``sesolve`` / ``mesolve`` with pyqtorch's call signature, running on the B200.

These are the two functions reference ``TorchEmulator.run`` calls
(``pulser_diff/backend.py:488-494`` and ``:502-509``); argument names, the returned object's
``.states`` layout -- ``(n_t, 2^N, B)`` kets, ``(n_t, 2^N, 2^N, 1)`` density matrices,
``states[0]`` = initial state -- and ``SolverType`` members are kept.  ``H`` must be the
:class:`~pulser_diff_b200.hamiltonian.StructuredHamiltonian` produced by our ``Hamiltonian``
(the one reference-side change); an opaque ``H(t)`` closure is rejected because
there is no CPU / dense fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Optional

import torch
from torch import Tensor

from . import _cabi, ops
from .hamiltonian import CollapseOperators, StructuredHamiltonian

C128 = torch.complex128


class SolverType(str, Enum):
    """Same members as ``pyqtorch.utils.SolverType`` used at backend.py:434,483,487,495."""
    DP5_SE = "dp5_se"
    DP5_ME = "dp5_me"
    KRYLOV_SE = "krylov_se"


_SOLVER_ID = {SolverType.DP5_SE: _cabi.SOLVER_DP5_SE, SolverType.KRYLOV_SE: _cabi.SOLVER_KRYLOV_SE,
              SolverType.DP5_ME: _cabi.SOLVER_DP5_ME}


@dataclass
class Result:
    states: Tensor
    _internal: Optional[Tensor] = None

    def step_log(self) -> list[dict]:
        return ops.last_step_log(self._internal)


def _check_h(H) -> StructuredHamiltonian:
    if not isinstance(H, StructuredHamiltonian):
        raise TypeError(
            "pulser_diff_b200 solvers need a StructuredHamiltonian (see "
            "pulser_diff_b200.hamiltonian.Hamiltonian); an opaque H(t) closure cannot be "
            "evaluated on the device and there is no CPU fallback.")
    return H


def _run(H: StructuredHamiltonian, state0: Tensor, tsave: Tensor, kind: int, solver: SolverType,
         options: Optional[dict], collapse: Optional[Tensor]) -> Tensor:
    dm, dv, am, av = H.masks_and_values()
    return ops.evolve(state0, tsave, dv, av, H.pair_u, n_qubits=H.n_qubits, kind=kind, dt=H.dt,
                      det_masks=dm, amp_masks=am, collapse=collapse, solver=_SOLVER_ID[solver],
                      options=_cabi.Options.from_dict(options))


def sesolve(H, psi0: Tensor, tsave: Tensor, solver: SolverType = SolverType.DP5_SE,
            options: Optional[dict] = None) -> Result:
    """Schroedinger evolution; ``psi0`` (2^N, B), result ``.states`` (n_t, 2^N, B)."""
    H = _check_h(H)
    if solver not in (SolverType.DP5_SE, SolverType.KRYLOV_SE):
        raise ValueError(f"Solver {solver} not available.")
    if psi0.dim() != 2 or psi0.shape[0] != 2 ** H.n_qubits:
        raise ValueError(f"Incompatible shape of initial state. Expected ({2 ** H.n_qubits}, B), "
                         f"got {tuple(psi0.shape)}.")
    state0 = psi0.to(device=H.device, dtype=_cabi.Options.from_dict(options).state_dtype).transpose(0, 1).contiguous()
    internal = _run(H, state0, tsave, _cabi.PD_KET, solver, options, None)
    return Result(internal.permute(0, 2, 1), internal)


def mesolve(H, rho0: Tensor, L, tsave: Tensor, solver: SolverType = SolverType.DP5_ME,
            options: Optional[dict] = None) -> Result:
    """Lindblad evolution; ``rho0`` (2^N, 2^N, 1), ``L`` = ``Hamiltonian._collapse_ops``,
    result ``.states`` (n_t, 2^N, 2^N, 1)."""
    H = _check_h(H)
    if solver != SolverType.DP5_ME:
        raise ValueError(f"Solver {solver} not available.")
    s = 2 ** H.n_qubits
    if rho0.dim() == 2:
        rho0 = rho0.unsqueeze(-1)
    if tuple(rho0.shape[:2]) != (s, s):
        raise ValueError(f"Incompatible shape of initial density matrix. Expected ({s}, {s}, B).")
    if isinstance(L, CollapseOperators):
        collapse = L.stacked()
    elif L is None or len(L) == 0 or all(isinstance(x, Tensor) and not torch.any(x != 0) for x in L):
        collapse = None          # backend.py:496-498: a single zero operator == no dissipation
    else:
        raise TypeError(
            "mesolve needs the CollapseOperators produced by pulser_diff_b200.hamiltonian."
            "Hamiltonian (single-qubit structure); dense 2^N x 2^N jump operators are not "
            "applied on the device.")
    b = rho0.shape[2]
    state0 = rho0.to(device=H.device, dtype=_cabi.Options.from_dict(options).state_dtype).permute(2, 0, 1).reshape(b, s * s).contiguous()
    internal = _run(H, state0, tsave, _cabi.PD_DENSITY, solver, options, collapse)
    n_t = internal.shape[0]
    return Result(internal.reshape(n_t, b, s, s).permute(0, 2, 3, 1), internal)
