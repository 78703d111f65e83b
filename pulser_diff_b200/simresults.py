"""Result container with the slice of the reference API that sits next to the hot path.

Mirrors ``CoherentResults`` of reference ``pulser_diff/simresults.py`` (:81-129 ``expect``,
:131-157 ``sample_state`` / ``sample_final_state``, :398-401 ``states``) without Pulser's
``Results`` base class.  Bitstring weights and the sampling itself stay on the device
(reference result.py:71-120); plotting, measurement-error flips and the
QuTiP pseudo-density are host statistics and out of scope.
"""
from __future__ import annotations

from collections import Counter
from typing import Optional, Sequence

import torch
from torch import Tensor

from .utils import expect, expect_diag


class CoherentResults:
    def __init__(self, states: Tensor, size: int, basis_name: str, sim_times: Tensor,
                 meas_basis: str = "ground-rydberg", meas_errors: Optional[dict] = None) -> None:
        if basis_name not in {"ground-rydberg", "digital", "all", "XY"}:
            raise ValueError("`basis_name` must be 'ground-rydberg', 'digital', 'all' or 'XY'.")
        self._states, self._size, self._basis_name = states, size, basis_name
        self._sim_times, self._meas_basis = sim_times, meas_basis
        self._meas_errors = meas_errors          # {"epsilon", "epsilon_prime"}: detection errors (SPAM)
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
        return self._states[self._index_of(t, t_tol)]

    def _index_of(self, t: float, t_tol: float) -> int:
        idx = int(torch.argmin(torch.abs(self._sim_times.detach() - t)))
        if abs(float(self._sim_times[idx]) - t) > t_tol:
            raise IndexError(f"Given time {t} is absent from the evaluation times within {t_tol}.")
        return idx

    def _weights(self, idx: int) -> Tensor:
        """Probabilities of the 2^N measured bitstrings at evaluation time ``idx`` (on the state's
        device).  The state vector is ordered with r first ([rr, rg, gr, gg]) and r is measured as
        1, so entry j of the result belongs to the bitstring ``binary_repr(j)``
        (reference result.py:71-87)."""
        st = self._states[idx].detach()
        if st.dim() == 3:                                  # (S, S, 1) density matrix
            probs = st[..., 0].diagonal().abs()
        else:                                              # (S, B) kets; the reference takes B = 1
            probs = (st.abs() ** 2).sum(dim=1)
        weights = probs.flip(0) if self._meas_basis == "ground-rydberg" else probs
        return weights / weights.sum()

    def sample_state(self, t: float, n_samples: int = 1000, t_tol: float = 1.0e-3) -> Counter:
        """Bitstring counts of ``n_samples`` projective measurements at time ``t``
        (reference simresults.py:131-145).  Drawn on the device by inverting the cumulative
        distribution, so the 2^N probabilities never travel to the host."""
        w = self._weights(self._index_of(t, t_tol))
        cdf = torch.cumsum(w, dim=0)
        u = torch.rand(int(n_samples), dtype=cdf.dtype, device=cdf.device) * cdf[-1]
        hits = torch.searchsorted(cdf, u, right=True).clamp_(max=w.numel() - 1)
        if self._meas_errors:
            # independent detection errors per shot and atom (reference simresults.py via Pulser's
            # SampledResult): a 0 reads as 1 with probability epsilon, a 1 as 0 with epsilon_prime
            n = self._size
            shifts = torch.arange(n - 1, -1, -1, device=hits.device)
            bits = (hits[:, None] >> shifts) & 1
            r = torch.rand(bits.shape, dtype=torch.float64, device=hits.device)
            flip = torch.where(bits == 0, r < self._meas_errors["epsilon"], r < self._meas_errors["epsilon_prime"])
            hits = ((bits ^ flip.to(bits.dtype)) << shifts).sum(dim=1)
        values, counts = torch.unique(hits, return_counts=True)
        n = self._size
        return Counter({format(int(v), f"0{n}b"): int(c) for v, c in zip(values.tolist(), counts.tolist())})

    def sample_final_state(self, N_samples: int = 1000) -> Counter:
        """Bitstring counts of the final state (reference simresults.py:147-157)."""
        return self.sample_state(float(self._sim_times[-1]), N_samples)

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


class NoisyResults:
    """Bitstring statistics accumulated over the random Hamiltonians of a noisy run (reference
    simresults.py:225-300 without the QuTiP pseudo-density): one ``Counter`` per evaluation time,
    ``n_measures = runs * samples_per_run`` shots each."""

    def __init__(self, counts: Sequence[Counter], size: int, basis_name: str, sim_times: Tensor,
                 n_measures: int) -> None:
        self._results = [Counter(c) for c in counts]
        self._size, self._basis_name, self._sim_times, self.n_measures = size, basis_name, sim_times, n_measures

    def __len__(self) -> int:
        return len(self._results)

    @property
    def results(self) -> list:
        return self._results

    def _index_of(self, t: float, t_tol: float) -> int:
        idx = int(torch.argmin(torch.abs(self._sim_times.detach() - t)))
        if abs(float(self._sim_times[idx]) - t) > t_tol:
            raise IndexError(f"Given time {t} is absent from the evaluation times within {t_tol}.")
        return idx

    def probabilities(self, t: float, t_tol: float = 1.0e-3) -> Tensor:
        """(2^N,) relative frequencies of the measured bitstrings at time ``t`` (index = bitstring value)."""
        c = self._results[self._index_of(t, t_tol)]
        out = torch.zeros(2 ** self._size, dtype=torch.float64)
        for k, v in c.items():
            out[int(k, 2)] = v
        return out / max(1, sum(c.values()))

    def get_state(self, t: float, t_tol: float = 1.0e-3) -> Tensor:
        """Diagonal pseudo-density matrix built from the frequencies (basis order r first, as the kets)."""
        return torch.diag(self.probabilities(t, t_tol).flip(0)).to(torch.complex128)

    def get_final_state(self) -> Tensor:
        return self.get_state(float(self._sim_times[-1]))

    def sample_state(self, t: float, n_samples: int = 1000, t_tol: float = 1.0e-3) -> Counter:
        p = self.probabilities(t, t_tol)
        draws = torch.multinomial(p, int(n_samples), replacement=True)
        values, counts = torch.unique(draws, return_counts=True)
        return Counter({format(int(v), f"0{self._size}b"): int(c) for v, c in zip(values.tolist(), counts.tolist())})

    def sample_final_state(self, N_samples: int = 1000) -> Counter:
        return self.sample_state(float(self._sim_times[-1]), N_samples)

    def expect(self, obs_list: Sequence[Tensor]) -> list[Tensor]:
        """Expectation values of DIAGONAL observables (1-D tensors of length 2^N, or the diagonal of a
        square one) from the frequencies, at every evaluation time."""
        out = []
        for obs in obs_list:
            d = (obs if obs.dim() == 1 else obs.diagonal()).to("cpu").real.to(torch.float64)
            out.append(torch.stack([(self.probabilities(float(t)).flip(0) * d).sum() for t in self._sim_times]))
        return out
