"""Replays golden vectors produced by the reference's own code (tests/golden/make_reference_fixtures.py)."""
import torch

from helpers import golden
from pulser_diff_b200.derivative import _extrapolate_borders, deriv_param, deriv_time

GOLD = golden("reference_derivative.json")


def test_border_extrapolation_matches_reference_outputs():
    for c in GOLD["fix_border_vals"]:
        got = _extrapolate_borders(torch.tensor(c["deriv"], dtype=torch.float64), c["borders"],
                                   torch.tensor(c["dt"], dtype=torch.float64))
        want = torch.tensor(c["fixed"], dtype=torch.float64)
        assert (got - want).abs().max() <= 1e-13 * want.abs().max()


def test_deriv_helpers_match_reference_outputs():
    c = GOLD["analytic"]
    t = torch.linspace(0.0, 1.0, 21, dtype=torch.float64, requires_grad=True)
    a = torch.tensor(c["a"], dtype=torch.float64, requires_grad=True)
    b = torch.tensor(c["b"], dtype=torch.float64, requires_grad=True)
    f = torch.sin(a * t) * torch.exp(b * t)
    assert (deriv_time(f, t, c["endtimes"]) - torch.tensor(c["deriv_time"], dtype=torch.float64)).abs().max() < 1e-13
    got = deriv_param(f, [a, b], t, c["t_ns"])
    assert all(abs(float(x) - y) < 1e-13 for x, y in zip(got, c["deriv_param"]))
