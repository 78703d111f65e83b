"""Derivative helpers with the reference's API (reference pulser_diff/derivative.py:26-78).

They are thin ``torch.autograd.grad`` calls; they work unchanged on the B200 path because
:func:`pulser_diff_b200.ops.evolve` is a proper autograd node that survives
``retain_graph=True`` and repeated one-hot cotangents.
"""
from __future__ import annotations

import torch
from torch import Tensor


def _extrapolate_borders(deriv: Tensor, borders: list, dt: Tensor) -> Tensor:
    """Replace derivative samples at pulse boundaries by linear extrapolation from their
    neighbours (same rule as reference derivative.py:7-23)."""
    prev = 0
    with torch.no_grad():
        for idx in borders:
            if idx == 0:
                deriv[0] = deriv[2] - ((deriv[2] - deriv[1]) / dt) * 2 * dt
            elif (idx - prev) != 1 or idx + 3 >= len(deriv):
                deriv[idx - 1] = deriv[idx - 3] + ((deriv[idx - 2] - deriv[idx - 3]) / dt) * 2 * dt
                deriv[idx] = deriv[idx - 2] + ((deriv[idx - 1] - deriv[idx - 2]) / dt) * 2 * dt
            else:
                deriv[idx] = deriv[idx + 2] - ((deriv[idx + 2] - deriv[idx + 1]) / dt) * 2 * dt
            prev = idx
    return deriv


def deriv_time(f: Tensor, times: Tensor, pulse_endtimes: list | None = None) -> Tensor:
    """df/dt at every evaluation time (``times`` must have been run with ``time_grad=True``)."""
    res = torch.autograd.grad(f, times, torch.ones_like(f), retain_graph=True)[0]
    if pulse_endtimes is not None:
        res = _extrapolate_borders(res, pulse_endtimes, times[1] - times[0])
    return res


def deriv_param(f: Tensor, x: list[Tensor], times: Tensor | None = None,
                t: int | float | Tensor | None = None):
    """df(t)/dx for every tensor in ``x``; ``t`` in ns, default = final time."""
    v = torch.zeros(len(f), dtype=f.dtype if not f.dtype.is_complex else torch.float64,
                    device=f.device)
    if times is None:
        v[-1] = 1.0
    else:
        tt = float(times[-1] if t is None else float(t) / 1000)
        v[torch.abs(times.detach().to(v.device) - tt).argmin()] = 1.0
    return torch.autograd.grad(f, x, v.to(f.dtype), retain_graph=True)
