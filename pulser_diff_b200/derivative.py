"""Derivative helpers with the reference's API (reference pulser_diff/derivative.py:26-78).

They are thin ``torch.autograd.grad`` calls; they work unchanged on the B200 path because
:func:`pulser_diff_b200.ops.evolve` is a proper autograd node that survives
``retain_graph=True`` and repeated one-hot cotangents.
"""
from __future__ import annotations

import torch
from torch import Tensor


def _line_through(a: Tensor, b: Tensor) -> Tensor:
    """Value two grid steps after ``a`` on the straight line through consecutive samples a, b."""
    return a + 2 * (b - a)


def _extrapolate_borders(deriv: Tensor, borders: list, dt: Tensor) -> Tensor:
    """Derivative samples at pulse boundaries are artefacts of the piecewise pulse profile; each is
    replaced, in order, by the straight line through two neighbouring samples (rule of reference
    derivative.py:7-23; ``dt`` cancels out of it and is kept for signature compatibility).

    An isolated boundary (or one too close to the end) is a pair (idx-1, idx) rebuilt from the
    samples before it; a boundary directly following another one, and index 0, are rebuilt from
    the two samples after it."""
    n = len(deriv)
    backward = [i != 0 and (i - p != 1 or i + 3 >= n)
                for i, p in zip(borders, [0] + list(borders[:-1]))]
    with torch.no_grad():
        for i, from_left in zip(borders, backward):
            if from_left:
                deriv[i - 1] = _line_through(deriv[i - 3], deriv[i - 2])
                deriv[i] = _line_through(deriv[i - 2], deriv[i - 1])
            else:
                deriv[i] = _line_through(deriv[i + 2], deriv[i + 1])
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
