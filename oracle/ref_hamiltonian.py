"""Restatement of the reference's Hamiltonian assembly (sparse COO, CPU).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, function by function:

* sample sub-sampling            reference hamiltonian.py:83-91
* basis / projectors             reference hamiltonian.py:288-318, utils.py:108-133
* N-fold sparse Kronecker        reference utils.py:12-44, hamiltonian.py:221-268
* van-der-Waals term             reference hamiltonian.py:333-344, 385-404
* coefficient arrays             reference hamiltonian.py:406-454
* H(t) closure with the interpolation quirk   reference hamiltonian.py:499-548
* collapse operators             reference hamiltonian.py:98-143

Conventions: basis order (r, g) so |r> = e0, |g> = e1; qubit 0
is the most significant bit; all-ground = last basis vector.

Inputs are plain tensors instead of Pulser objects (Pulser is not installable
here): qubit coordinates, C6, and the (T+1)-long amp/det/phase sample arrays
per addressing ("Global" or per-qubit "Local").
"""
from __future__ import annotations

from math import floor
from typing import Callable

import torch
from torch import Tensor

C128 = torch.complex128

# single-qubit matrices in the (r, g) basis
_I2 = torch.eye(2, dtype=C128)
_X = torch.tensor([[0, 1], [1, 0]], dtype=C128)
_Y = torch.tensor([[0, -1j], [1j, 0]], dtype=C128)
_Z = torch.tensor([[1, 0], [0, -1]], dtype=C128)
_SIG_RR = torch.tensor([[1, 0], [0, 0]], dtype=C128)   # |r><r|
_SIG_GG = torch.tensor([[0, 0], [0, 1]], dtype=C128)   # |g><g|
_SIG_GR = torch.tensor([[0, 0], [1, 0]], dtype=C128)   # |g><r|
OPS = {"I": _I2, "X": _X, "Y": _Y, "Z": _Z, "sigma_rr": _SIG_RR, "sigma_gg": _SIG_GG,
       "sigma_gr": _SIG_GR}


def sparse_kron(mats: list[Tensor]) -> Tensor:
    """Kronecker product of sparse COO factors, right to left (utils.py:12-44)."""
    acc = mats[-1].coalesce()
    for m in reversed(mats[:-1]):
        m = m.coalesce()
        rows, cols = acc.shape
        idx, val = [], []
        for (i, j), v in zip(m.indices().T.tolist(), m.values()):
            idx.append(acc.indices() + torch.tensor([[i * rows], [j * cols]]))
            val.append(v * acc.values())
        acc = torch.sparse_coo_tensor(torch.cat(idx, 1), torch.cat(val),
                                      (m.shape[0] * rows, m.shape[1] * cols)).coalesce()
    return acc


def embed(n: int, placed: dict[int, Tensor], dense: bool = False) -> Tensor:
    """Operator acting as ``placed[q]`` on qubit q and identity elsewhere.

    Sparse when every factor is sparse, dense otherwise (utils.py:13-16); the
    reference reaches the dense branch for collapse operators because XMAT /
    ZMAT are dense (hamiltonian.py:113, 127-129, 142).
    """
    facs = [placed.get(q, _I2) for q in range(n)]
    if dense:
        out = facs[0]
        for f in facs[1:]:
            out = torch.kron(out, f)
        return out
    return sparse_kron([f.to_sparse() for f in facs])


def subsample_indices(n_full: int, rate: float) -> Tensor:
    """reference hamiltonian.py:85-90: linspace(0, len-1, int(rate*duration), dtype=int)."""
    return torch.linspace(0, n_full - 1, int(rate * n_full), dtype=torch.int)


class RefHamiltonian:
    """Term list + ``H_t`` closure, the reference way.

    Args:
        coords: (N, 2) float64 tensor (may require grad).
        c6: interaction coefficient (MockDevice: 5420158.53, level 60: 865723.02).
        samples: ``{"Global": {"amp","det","phase"}}`` and/or
            ``{"Local": {q: {"amp","det","phase"}}}`` with (T+1)-long arrays.
        rate: sampling_rate.
        noise: dict with optional ``dephasing_rate``, ``relaxation_rate``,
            ``depolarizing_rate``, ``eff_noise`` = list of (rate, 2x2 op).
    """

    def __init__(self, coords: Tensor, c6: float, samples: dict, rate: float,
                 noise: dict | None = None) -> None:
        self.coords = coords
        self.n = int(coords.shape[0])
        self.size = 2 ** self.n
        self.c6 = float(c6)
        self.rate = float(rate)
        first = (samples["Global"] if "Global" in samples and samples["Global"]
                 else next(iter(samples["Local"].values())))
        self.duration = int(first["amp"].numel())          # T + 1
        self.idx = subsample_indices(self.duration, self.rate)
        self.sampling_times = (torch.arange(self.duration, dtype=torch.double) / 1000)[self.idx]
        self.dist: dict[tuple[int, int], Tensor] = {}
        self.noise = noise or {}
        self.terms: list[list] = []
        self._build_terms(samples)
        self.collapse_ops = self._collapse_ops()
        self.H = self._closure()

    # -- reference hamiltonian.py:333-344, 385-404 ---------------------------------
    def _interaction(self) -> Tensor:
        acc = torch.sparse_coo_tensor(torch.tensor([[0], [0]]), [0], (self.size, self.size),
                                      dtype=C128)
        for i in range(self.n):
            for j in range(i + 1, self.n):
                d = torch.linalg.norm(self.coords[i] - self.coords[j])
                self.dist[(i, j)] = d
                u = (1.0 + 0.0j) * 0.5 * self.c6 / d ** 6
                acc = acc + embed(self.n, {i: _SIG_RR, j: _SIG_RR}) * u
        return acc

    # -- reference hamiltonian.py:406-454 ------------------------------------------
    def _coeff_terms(self, s: dict, qubits: list[int]) -> list[list]:
        out = []
        drive = 0.5 * s["amp"] * torch.exp(-1j * s["phase"])
        det = -0.5 * s["det"]
        for op, coeff in ((_SIG_GR, drive), (_SIG_RR, det)):
            if torch.any(coeff != 0):
                mat = embed(self.n, {qubits[0]: op})
                for q in qubits[1:]:
                    mat = mat + embed(self.n, {q: op})
                out.append([mat.coalesce(), coeff[self.idx], op is _SIG_RR, list(qubits)])
        return out

    def _build_terms(self, samples: dict) -> None:
        self.int_mat = (self._interaction() if self.n > 1
                        else torch.zeros((2, 2), dtype=C128).to_sparse())
        if samples.get("Global"):
            self.terms += self._coeff_terms(samples["Global"], list(range(self.n)))
        for q, s in (samples.get("Local") or {}).items():
            self.terms += self._coeff_terms(s, [int(q)])

    # -- reference hamiltonian.py:98-143 -------------------------------------------
    def _collapse_ops(self) -> list[Tensor]:
        local = []
        nz = self.noise
        if "dephasing_rate" in nz:
            local.append(torch.sqrt(torch.as_tensor(nz["dephasing_rate"]) / 2) * _Z)
        if "relaxation_rate" in nz:
            local.append(torch.sqrt(torch.as_tensor(nz["relaxation_rate"])) * _SIG_GR)
        if "depolarizing_rate" in nz:
            c = torch.sqrt(torch.as_tensor(nz["depolarizing_rate"]) / 4)
            local += [c * _X, c * _Y, c * _Z]
        for rate, op in nz.get("eff_noise", []):
            local.append(torch.sqrt(torch.as_tensor(rate)) * torch.as_tensor(op, dtype=C128))
        self.local_ops = local          # the single-qubit operators, rate folded in
        return [embed(self.n, {q: op}, dense=True) for op in local for q in range(self.n)]

    # -- reference hamiltonian.py:499-548 ------------------------------------------
    def _closure(self) -> Callable[[float | Tensor], Tensor]:
        det_terms = [(m, (1.0 + 0.0j) * c) for m, c, is_diag, _ in self.terms if is_diag]
        amp_terms = [(m, c) for m, c, is_diag, _ in self.terms if not is_diag]
        dt = 0.001 / self.rate
        n_samples = len(self.terms[-1][1]) if self.terms else int(self.idx.numel())
        self.dt, self.n_samples = dt, n_samples
        int_mat = self.int_mat

        def H_t(t):
            if not isinstance(t, Tensor):
                t = torch.tensor(t)
            i1 = max(int(min(floor(float(t) / dt), n_samples - 2)), 0)
            i2 = min(i1 + 1, n_samples - 2)
            ham = 2 * int_mat
            for mat, val in det_terms + amp_terms:
                c = val[i1] + (val[i2] - val[i1]) * (t - i1 * dt) / dt
                piece = mat * c
                ham = ham + (piece + piece.adjoint())
            return ham

        return H_t

    # -- structure-only view, used to feed the CUDA path in parity tests -----------
    def pair_couplings(self) -> Tensor:
        """(N, N) upper-triangular C6 / r_ij**6 (the net n_i n_j coefficient)."""
        u = torch.zeros(self.n, self.n, dtype=torch.float64)
        for (i, j), d in self.dist.items():
            u[i, j] = self.c6 / d ** 6
        return u
