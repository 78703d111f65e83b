"""Restatement of the emulator front-end around the solver call.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows:

* initial state |g...g> = last basis vector, forced complex128
                                         reference backend.py:253-280
* evaluation times ("Full" / "Minimal" / list / float), always joined with
  {0, T/1000} and ``unique()``-sorted     reference backend.py:312-375
* solver dispatch, Lindblad forcing DP5_ME, rho0 = psi psi^dagger[..., None]
                                         reference backend.py:477-509
* ``expect`` for kets and density matrices   reference utils.py:68-86
"""
from __future__ import annotations

import torch
from torch import Tensor

from .ref_hamiltonian import RefHamiltonian, embed, OPS
from .ref_solvers import SolverType, mesolve, sesolve

C128 = torch.complex128


def total_magnetization(n: int) -> Tensor:
    """Dense sum_i Z_i (reference utils.py:47-65)."""
    out = torch.zeros(2 ** n, 2 ** n, dtype=C128)
    for q in range(n):
        out = out + embed(n, {q: OPS["Z"]}, dense=True)
    return out


def expect(obs: Tensor, states: Tensor) -> Tensor:
    """Dense-observable branch of reference utils.py:79-84."""
    if states.dim() == 3:
        return torch.einsum("...ij,jk,...kl->...", states.mH, obs, states)
    return torch.einsum("ij,...jik->...", obs, states)


class RefEmulator:
    def __init__(self, coords: Tensor, c6: float, samples: dict, rate: float = 1.0,
                 noise: dict | None = None, evaluation_times="Full") -> None:
        self.ham = RefHamiltonian(coords, c6, samples, rate, noise)
        self.n = self.ham.n
        self.tot_duration = self.ham.duration - 1                  # T (ns)
        self.noise = noise or {}
        self.set_evaluation_times(evaluation_times)
        psi = torch.zeros(2 ** self.n, 1, dtype=C128)
        psi[-1] = 1.0
        self.initial_state = psi

    def set_initial_state(self, state: Tensor) -> None:
        if state.shape[0] != 2 ** self.n:
            raise ValueError("Incompatible shape of initial state.")
        self.initial_state = state.to(C128)

    def set_evaluation_times(self, value) -> None:
        st = self.ham.sampling_times
        if isinstance(value, str):
            if value == "Full":
                ev = st.clone()
            elif value == "Minimal":
                ev = torch.tensor([], dtype=torch.float64)
            else:
                raise ValueError("Wrong evaluation time label.")
        elif isinstance(value, float):
            if value > 1 or value <= 0:
                raise ValueError("evaluation_times float must be between 0 and 1.")
            ind = torch.linspace(0, len(st) - 1, int(value * len(st)), dtype=torch.int)
            ev = st[ind]
        else:
            ev = torch.as_tensor(value, dtype=torch.float64)
            if ev.max() > self.tot_duration / 1000 or ev.min() < 0:
                raise ValueError("Provided evaluation-time list out of range.")
        self.evaluation_times = torch.cat(
            [ev, torch.tensor([0.0, self.tot_duration / 1000], dtype=ev.dtype)]).unique()

    def run(self, time_grad: bool = False, solver: SolverType = SolverType.DP5_SE,
            replay: list | None = None, **options):
        if time_grad:
            self.evaluation_times.requires_grad_(True)
        if any(k in self.noise for k in ("dephasing_rate", "relaxation_rate",
                                         "depolarizing_rate", "eff_noise")):
            solver = SolverType.DP5_ME
        if solver in (SolverType.DP5_SE, SolverType.KRYLOV_SE):
            return sesolve(self.ham.H, self.initial_state, self.evaluation_times, solver,
                           options, replay=replay) if solver == SolverType.DP5_SE else \
                sesolve(self.ham.H, self.initial_state, self.evaluation_times, solver, options)
        if solver == SolverType.DP5_ME:
            L = self.ham.collapse_ops or [torch.zeros(2 ** self.n, 2 ** self.n, dtype=C128)]
            rho0 = torch.matmul(self.initial_state, self.initial_state.mH).unsqueeze(-1)
            return mesolve(self.ham.H, rho0, L, self.evaluation_times, solver, options,
                           replay=replay)
        raise ValueError(f"Solver {solver} not available.")
