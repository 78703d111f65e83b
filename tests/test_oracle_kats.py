"""The oracle against every known answer the reference holds for this path.

The reference has no golden vectors; its notebooks PRINT seven numbers, one
160-point series and a handful of state entries.  Those are in tests/golden/notebook_kats.json
(made by tests/golden/make_notebook_kats.py).  The state entries include values at the 1e-20
level that depend on the exact accepted-step sequence, so they pin the restated step
controller, not just the ODE solution.
"""
import math

import pytest
import torch

from helpers import KAT_SOLVER, golden, kat_problem
from oracle.ref_emulator import expect, total_magnetization
from oracle.ref_solvers import SolverType

GOLD = golden("notebook_kats.json")


def _sum_z(em, res):
    return expect(total_magnetization(em.n), res.states).real


@pytest.mark.parametrize("name", ["K-B", "K-C", "K-D", "K-E", "K-F"])
def test_final_magnetisation(name):
    em = kat_problem(name).ref()
    res = em.run(solver=SolverType(KAT_SOLVER[name]))
    assert abs(_sum_z(em, res)[-1].item() - GOLD[name]["final_sum_z"]) < 6e-5   # 4 printed decimals


def test_ka_series_and_printed_states():
    em = kat_problem("K-A").ref()
    res = em.run(solver=SolverType.DP5_SE)
    g = GOLD["K-A"]
    assert torch.allclose(em.evaluation_times, torch.tensor(g["eval_times"], dtype=torch.float64), atol=1e-9)
    z = _sum_z(em, res)
    assert (z - torch.tensor(g["sum_z"], dtype=torch.float64)).abs().max() < 6e-5
    for k, entries in g["states_printed"].items():
        for i, (re, im) in entries.items():
            got = res.states[int(k), int(i), 0]
            for want, have in ((re, got.real.item()), (im, got.imag.item())):
                if int(k) > 0:
                    # first steps: 5 printed significant digits, down to 1e-20 magnitudes
                    assert abs(have - want) <= 6e-5 * abs(want) + 1e-30, (k, i, want, have)
                else:
                    # end of the run: agreement to the solver's own tolerance (rtol 1e-6)
                    assert abs(have - want) <= max(2e-6, 6e-5 * abs(want)), (k, i, want, have)


def test_kg_gate_infidelity():
    em = kat_problem("K-G").ref()
    em.set_initial_state(torch.eye(4))
    res = em.run(solver=SolverType.DP5_SE)
    h = torch.tensor([[1, 1], [1, -1]], dtype=torch.complex128) / math.sqrt(2)
    infid = 1 - abs(torch.trace(torch.kron(h, h).mH @ res.states[-1])) / 4
    assert abs(infid.item() - GOLD["K-G"]["infidelity"]) < 6e-7
