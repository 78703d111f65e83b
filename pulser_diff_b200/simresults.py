"""Result container with the slice of the reference API that sits next to the hot path.

Mirrors ``CoherentResults`` of reference ``pulser_diff/simresults.py`` (:81-129 ``expect``,
:398-401 ``states``) without Pulser's ``Results`` base class.  Sampling, plotting and
pseudo-density helpers are host statistics and out of scope.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor

from .utils import expect, expect_diag


class CoherentResults:
    def __init__(self, states: Tensor, size: int, basis_name: str, sim_times: Tensor,
                 meas_basis: str = "ground-rydberg") -> None:
        if basis_name not in {"ground-rydberg", "digital", "all", "XY"}:
            raise ValueError("`basis_name` must be 'ground-rydberg', 'digital', 'all' or 'XY'.")
        self._states, self._size, self._basis_name = states, size, basis_name
        self._sim_times, self._meas_basis = sim_times, meas_basis
        self._dim = 2

    def __len__(self) -> int:
        return int(self._states.shape[0])

    @property
    def states(self) -> Tensor:
        """(n_t, 2^N, B) kets or (n_t, 2^N, 2^N, 1) density matrices, on the CUDA device."""
        return self._states

    def get_final_state(self) -> Tensor:
        return self._states[-1]

    def get_state(self, t: float, t_tol: float = 1.0e-3) -> Tensor:
        idx = int(torch.argmin(torch.abs(self._sim_times.detach() - t)))
        if abs(float(self._sim_times[idx]) - t) > t_tol:
            raise IndexError(f"Given time {t} is absent from the evaluation times within {t_tol}.")
        return self._states[idx]

    def expect(self, obs_list: Sequence[Tensor]) -> list[Tensor]:
        """Expectation values of the operators in ``obs_list`` at every evaluation time.

        (S, S) tensors follow the reference (simresults.py:81-129); a 1-D tensor of length S is
        taken as the DIAGONAL of the observable and reduced by the fused device kernel.
        """
        if not isinstance(obs_list, (list, Tensor)):
            raise TypeError("`obs_list` must be a list of operators.")
        legal = (self._dim ** self._size, self._dim ** self._size)
        out = []
        for obs in obs_list:
            if not isinstance(obs, Tensor):
                raise TypeError(f"Incompatible type {type(obs)} of observable.")
            if obs.dim() == 1 and obs.shape[0] == legal[0]:
                out.append(expect_diag(obs, self._states))
                continue
            if tuple(obs.shape) != legal:
                raise ValueError(f"Incompatible shape of observable.Expected {legal}, got {obs.shape}.")
            out.append(expect(obs, self._states))
        return out
