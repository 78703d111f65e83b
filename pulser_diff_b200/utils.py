"""Small linear-algebra helpers with the reference's names and conventions.

Mirrors the helpers of reference ``pulser_diff/utils.py`` that sit next to the hot path
(``kron`` :12-44, ``total_magnetization`` :47-65, ``expect`` :68-86, ``trace`` :89-94,
``basis_state`` :108-133) so notebooks and tests read the same.  Dense / sparse torch code for
small registers; the device path for diagonal observables is :func:`expect_diag`.
"""
from __future__ import annotations

from functools import reduce
from math import prod

import torch
from torch import Tensor

C128 = torch.complex128
IMAT = torch.eye(2, dtype=C128)
XMAT = torch.tensor([[0, 1], [1, 0]], dtype=C128)
YMAT = torch.tensor([[0, -1j], [1j, 0]], dtype=C128)
ZMAT = torch.tensor([[1, 0], [0, -1]], dtype=C128)
HMAT = torch.tensor([[1, 1], [1, -1]], dtype=C128) / 2 ** 0.5


def kron(*args: Tensor) -> Tensor:
    """Kronecker product; sparse iff every factor is sparse (reference utils.py:12-44)."""
    if not all(t.is_sparse for t in args):
        return reduce(torch.kron, [t.to_dense() if t.is_sparse else t for t in args])
    out = args[-1].coalesce()
    for m in reversed(args[:-1]):
        m = m.coalesce()
        r, c = out.shape
        bi, bj = m.indices()
        oi, oj = out.indices()
        idx = torch.stack([(bi[:, None] * r + oi[None, :]).reshape(-1),
                           (bj[:, None] * c + oj[None, :]).reshape(-1)])
        val = (m.values()[:, None] * out.values()[None, :]).reshape(-1)
        out = torch.sparse_coo_tensor(idx, val, (m.shape[0] * r, m.shape[1] * c)).coalesce()
    return out


def basis_state(dim, number) -> Tensor:
    """(prod(dim), 1) column with a single 1 at the row-major index of ``number`` in a register of
    local dimensions ``dim`` (contract of reference utils.py:108-133)."""
    dims = [dim] if isinstance(dim, int) else list(dim)
    occ = [number] if isinstance(number, int) else list(number)
    if len(dims) != len(occ):
        raise ValueError(
            "Arguments `number` must have the same length as `dim` of length"
            f" {len(dims)}, but has length {len(occ)}.")
    strides = [prod(dims[k + 1:]) for k in range(len(dims))]
    ket = torch.zeros(prod(dims), 1)
    ket[sum(o * st for o, st in zip(occ, strides))] = 1.0
    return ket


def total_magnetization(n_qubits: int, use_sparse: bool = False) -> Tensor:
    """sum_i Z_i as a (sparse) matrix (reference utils.py:47-65)."""
    obs = None
    for i in range(n_qubits):
        facs = [IMAT.to_sparse() if use_sparse else IMAT for _ in range(n_qubits)]
        facs[i] = ZMAT.to_sparse() if use_sparse else ZMAT
        term = kron(*facs)
        obs = term if obs is None else obs + term
    return obs


def total_magnetization_diag(n_qubits: int, device=None) -> Tensor:
    """Diagonal of sum_i Z_i: (#r bits) - (#g bits) with |r> = bit 0."""
    s = torch.arange(2 ** n_qubits, device=device)
    ones = torch.zeros_like(s)
    for b in range(n_qubits):
        ones += (s >> b) & 1
    return (n_qubits - 2 * ones).to(torch.float64)


def occupation_diag(n_qubits: int, qubits, device=None) -> Tensor:
    """Diagonal of prod_{q in qubits} n_q (Rydberg occupation, bit value 0)."""
    s = torch.arange(2 ** n_qubits, device=device)
    out = torch.ones(2 ** n_qubits, dtype=torch.float64, device=device)
    for q in qubits:
        out = out * (1 - ((s >> (n_qubits - 1 - q)) & 1)).to(torch.float64)
    return out


def trace(mat: Tensor) -> Tensor:
    """Trace over the last two dims, sparse or dense (reference utils.py:89-94)."""
    if mat.is_sparse:
        mat = mat.coalesce()
        i, j = mat.indices()[-2], mat.indices()[-1]
        return mat.values()[i == j].sum()
    return torch.diagonal(mat, dim1=-2, dim2=-1).sum(-1)


def expect(obs: Tensor, states: Tensor) -> Tensor:
    """<O>(t) for kets (n_t, S, B) or density matrices (n_t, S, S, B) (reference utils.py:68-86).

    Runs on whatever device ``states`` lives on; a dense S x S observable is moved there.  For
    diagonal observables use :func:`expect_diag`, which never forms the matrix.
    """
    obs = obs.to(states.device)
    if obs.is_sparse:
        obs = obs.to_dense()
    obs = obs.to(states.dtype)
    if states.dim() == 3:
        return torch.einsum("...ij,jk,...kl->...", states.mH, obs, states)
    if states.dim() == 4:
        return torch.einsum("ij,...jik->...", obs, states)
    raise ValueError("states must be (n_t, S, B) kets or (n_t, S, S, B) density matrices")


class _DiagExpect(torch.autograd.Function):
    @staticmethod
    def forward(ctx, states: Tensor, diag: Tensor, kind: int):
        from . import ops
        ctx.kind = kind
        ctx.save_for_backward(states, diag)
        return torch.ops.pulser_diff_b200.expect_diag(states, diag, kind).to(states.device)

    @staticmethod
    def backward(ctx, g):
        states, diag = ctx.saved_tensors
        if ctx.kind == 0:
            gs = 2.0 * g.real[:, None, None] * diag[None, None, :] * states
        else:
            s = diag.numel()
            gs = torch.zeros_like(states)
            v = gs.view(states.shape[0], states.shape[1], s, s)
            torch.diagonal(v, dim1=-2, dim2=-1).copy_(g[:, None, None] * diag[None, None, :])
        return gs, None, None


def expect_diag(diag: Tensor, states: Tensor) -> Tensor:
    """Fused <diag(O)> on the device for states in the reference layout.

    ``states``: kets (n_t, S, B) or density matrices (n_t, S, S, B); returns (n_t,) complex128,
    summed over the batch columns exactly like the reference's einsum (utils.py:81-84).
    """
    if states.dim() == 3:
        internal = states.permute(0, 2, 1).contiguous()
        kind = 0
    elif states.dim() == 4:
        n_t, s = states.shape[0], states.shape[1]
        internal = states.permute(0, 3, 1, 2).reshape(n_t, states.shape[3], s * s).contiguous()
        kind = 1
    else:
        raise ValueError("states must be (n_t, S, B) kets or (n_t, S, S, B) density matrices")
    return _DiagExpect.apply(internal, diag.to(device=states.device, dtype=torch.float64), kind)


def interpolate_sine(num_values: int, duration: int) -> Tensor:
    """(duration, num_values) weights that blend consecutive control points with a sine ease
    (same matrix as reference utils.py:151-180; used by docs/state_preparation.ipynb)."""
    step = duration / (num_values + 1)
    k = torch.arange(duration, dtype=torch.float64)
    idx = torch.floor(k / step).to(torch.long)
    h = (k - idx * step) / step
    s = (1 + torch.sin(torch.pi * h - torch.pi / 2)) / 2
    mat = torch.zeros(duration, num_values, dtype=torch.float64)
    rows = torch.arange(duration)
    left = idx > 0
    mat[rows[left], idx[left] - 1] = (1 - s)[left]
    right = idx < num_values
    mat[rows[right], idx[right]] = s[right]
    return mat.to(torch.float32)


def s(t: float) -> float:
    """Sine ease between 0 and 1 for a normalised time ``t`` (reference utils.py:136-148)."""
    import math
    return 0.5 * (1.0 + math.sin(math.pi * (t - 0.5)))


def vn_entropy(rho: Tensor) -> Tensor:
    """Von Neumann entropy (base 2) of a density matrix (reference utils.py:97-105): only the
    strictly positive eigenvalues contribute."""
    ev = torch.linalg.eigvalsh(rho)
    pos = ev[ev > 0]
    return -(pos * torch.log2(pos)).sum()
