"""Restatement of ``pyqtorch.sesolve`` / ``pyqtorch.mesolve`` (torch CPU, tape autograd).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``pyqtorch`` is an UNPINNED
dependency of the reference (pyproject.toml:30), absent from /root/reference and
not installable here.  This file restates its published time-dependent solvers
from SURVEY.md Appendix A [UPSTREAM-RECALLED]: the call sites it must satisfy
are reference backend.py:488-494 (sesolve) and backend.py:502-509 (mesolve).

* driver loop                 Appendix A.1
* Dormand-Prince 5(4), FSAL   Appendix A.2
* step controller + initial step (Hairer)   Appendix A.3
* right-hand sides            Appendix A.4
* Krylov (H frozen at the interval END)      Appendix A.5
* backward = plain autograd tape             Appendix A.6

``steplog`` / ``replay`` are oracle-only additions for the shared-step-sequence
parity protocol: ``Result.steplog`` lists every attempted step
``(t, dt, accepted, error, clipped)``; ``replay`` = the ACCEPTED entries of such a
log, which are then taken verbatim (every replayed step counts as accepted).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from math import sqrt
from typing import Callable

import torch
from torch import Tensor


class SolverType(str, Enum):
    DP5_SE = "dp5_se"
    DP5_ME = "dp5_me"
    KRYLOV_SE = "krylov_se"


@dataclass
class AdaptiveOptions:
    atol: float = 1e-8
    rtol: float = 1e-6
    max_steps: int = 100_000
    safety_factor: float = 0.9
    min_factor: float = 0.2
    max_factor: float = 5.0
    use_sparse: bool = False


@dataclass
class KrylovOptions:
    max_krylov: int = 80
    exp_tolerance: float = 1e-10
    norm_tolerance: float = 1e-10
    use_sparse: bool = False


@dataclass
class Result:
    states: Tensor
    steplog: list = field(default_factory=list)


ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
B5 = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
B4 = [5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40]


def hairer_norm(x: Tensor) -> Tensor:
    """RMS over the state dimension(s), one value per batch column (last dim)."""
    dims = tuple(range(x.dim() - 1))
    n = 1
    for d in dims:
        n *= x.shape[d]
    return torch.sqrt((x.abs() ** 2).sum(dim=dims) / n)


class _DP5:
    def __init__(self, f: Callable[[Tensor, Tensor], Tensor], opt: AdaptiveOptions) -> None:
        self.f, self.opt = f, opt

    def error(self, y_err: Tensor, y0: Tensor, y1: Tensor) -> float:
        scale = self.opt.atol + self.opt.rtol * torch.max(y0.abs(), y1.abs())
        return float(hairer_norm(y_err / scale).max())

    def init_tstep(self, t0: Tensor, y0: Tensor, f0: Tensor) -> float:
        o = self.opt
        sc = o.atol + o.rtol * y0.abs()
        d0 = float(hairer_norm(y0 / sc).max())
        d1 = float(hairer_norm(f0 / sc).max())
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        y1 = y0 + h0 * f0
        f1 = self.f(t0 + h0, y1)
        d2 = float(hairer_norm((f1 - f0) / sc).max()) / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** (1.0 / 6.0)
        return min(100 * h0, h1)

    def update_tstep(self, dt: float, error: float) -> float:
        o = self.opt
        if error == 0:
            return dt * o.max_factor
        fac = o.safety_factor * error ** (-1.0 / 5.0)
        if error <= 1:
            return dt * max(1.0, min(o.max_factor, fac))
        return dt * min(0.9, max(o.min_factor, fac))

    def step(self, t: Tensor, y: Tensor, f0: Tensor, dt):
        k = [f0]
        for i in range(6):
            dy = sum(dt * b * kj for b, kj in zip(BETA[i], k) if b != 0.0)
            k.append(self.f(t + dt * ALPHA[i], y + dy))
        y1 = y + sum(dt * b * kj for b, kj in zip(B5[:6], k[:6]) if b != 0.0)
        y_err = sum(dt * (b5 - b4) * kj for b5, b4, kj in zip(B5, B4, k))
        return k[6], y1, y_err


def _integrate_adaptive(f, y0: Tensor, tsave: Tensor, opt: AdaptiveOptions,
                        replay: list | None = None):
    dp = _DP5(f, opt)
    t = tsave[0]
    y = y0
    ft = f(t, y)
    log: list = []
    saved = []
    if replay is None:
        dt, error = dp.init_tstep(t, y, ft), 1.0
    else:
        dt, error = 0.0, 1.0
        replay = list(replay)
    pos = 0
    for t_next in tsave:
        cache = (dt, error)
        steps = 0
        while float(t) < float(t_next):
            if replay is None:
                dt = dp.update_tstep(dt, error)
                clipped = float(t) + dt >= float(t_next)
            else:
                _, dt, _, _, clipped = replay[pos]
                pos += 1
            if clipped:
                cache = (dt, error)
                dt_used = t_next - t           # tensor: this is how tsave enters the tape
                dt = float(dt_used)            # upstream overwrites dt with the clipped value
            else:
                dt_used = dt
            f_new, y_new, y_err = dp.step(t, y, ft, dt_used)
            error = dp.error(y_err, y, y_new)
            accepted = True if replay is not None else error <= 1
            log.append((float(t), float(dt_used), bool(accepted), float(error), bool(clipped)))
            if accepted:
                t = t_next if clipped else t + dt_used
                y, ft = y_new, f_new
            steps += 1
            if steps >= opt.max_steps:
                raise RuntimeError("max_steps reached")
        dt, error = cache
        saved.append(y)
    return torch.stack(saved), log


def _se_rhs(H):
    def f(t, y):
        return -1j * (H(t) @ y)
    return f


def _me_rhs(H, L: list[Tensor]):
    Ls = torch.stack(L)                                  # (n_L, S, S)
    Ld = Ls.mH
    LdL = (Ld @ Ls).sum(0)

    def f(t, rho):                                       # rho (S, S, 1)
        r = rho[..., 0]
        Hm = H(t)
        out = -1j * (Hm @ r - (Hm.adjoint() @ r.mH).mH)  # sparse@dense twice: H rho - rho H
        out = out + (Ls @ r @ Ld).sum(0) - 0.5 * (LdL @ r) - 0.5 * (r @ LdL)
        return out[..., None]
    return f


# How exp(-i*delta*T) of the small tridiagonal T is evaluated.  Upstream calls
# torch.linalg.matrix_exp; on torch 2.11 that routine is only accurate to ~3e-12 for these
# matrices (measured: 3x3, norm 0.044), which over 900 intervals drifts 2e-9 from the exact
# propagator.  "eigh" evaluates the same quantity to round-off and stays differentiable, so it
# is the default for parity; "matrix_exp" reproduces upstream's call literally.
SMALL_EXPM = "eigh"


def _small_expm_col0(T: Tensor, delta) -> Tensor:
    if SMALL_EXPM == "matrix_exp":
        return torch.linalg.matrix_exp(-1j * delta * T)[:, 0]
    lam, Q = torch.linalg.eigh(T.real)
    return (Q.to(T.dtype) * torch.exp(-1j * delta * lam)[None, :]) @ Q[0, :].to(T.dtype)


def _krylov_exp(Hm, psi: Tensor, delta, opt: KrylovOptions) -> Tensor:
    """exp(-i*delta*H) psi for one column, Lanczos with full re-use (Appendix A.5)."""
    nrm = torch.linalg.norm(psi)
    v = [psi / nrm]
    a, b = [], []
    w = None
    for j in range(opt.max_krylov):
        r = Hm @ v[-1]
        a.append(torch.vdot(v[-1], r).real)
        r = r - a[-1] * v[-1] - (b[-1] * v[-2] if j > 0 else 0)
        beta = torch.linalg.norm(r)
        T = torch.diag(torch.stack(a)).to(torch.complex128)
        if b:
            off = torch.stack(b).to(torch.complex128)
            T = T + torch.diag(off, 1) + torch.diag(off, -1)
        w = _small_expm_col0(T, delta)
        if float(beta) < opt.norm_tolerance:
            break
        if j >= 1 and float((w[-1].abs() + w[-2].abs()) * beta * abs(float(delta))) < opt.exp_tolerance:
            break
        b.append(beta)
        v.append(r / beta)
    V = torch.stack(v[: w.numel()], dim=1)
    return nrm * (V @ w)


def sesolve(H, psi0: Tensor, tsave: Tensor, solver: SolverType = SolverType.DP5_SE,
            options: dict | None = None, replay: list | None = None) -> Result:
    options = dict(options or {})
    if solver == SolverType.DP5_SE:
        st, log = _integrate_adaptive(_se_rhs(H), psi0, tsave, AdaptiveOptions(**options), replay)
        return Result(st, log)
    if solver == SolverType.KRYLOV_SE:
        opt = KrylovOptions(**options)
        y, t, saved = psi0, tsave[0], []
        for t_next in tsave:
            if float(t_next) > float(t):
                Hm = H(t_next)
                y = torch.stack([_krylov_exp(Hm, y[:, c], t_next - t, opt)
                                 for c in range(y.shape[1])], dim=1)
            saved.append(y)
            t = t_next
        return Result(torch.stack(saved))
    raise ValueError(f"Solver {solver} not available.")


def mesolve(H, rho0: Tensor, L: list[Tensor], tsave: Tensor,
            solver: SolverType = SolverType.DP5_ME, options: dict | None = None,
            replay: list | None = None) -> Result:
    if solver != SolverType.DP5_ME:
        raise ValueError(f"Solver {solver} not available.")
    st, log = _integrate_adaptive(_me_rhs(H, L), rho0, tsave,
                                  AdaptiveOptions(**dict(options or {})), replay)
    return Result(st, log)
