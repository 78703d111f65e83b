"""The oracle (and the engine) against an INDEPENDENT solution of the same physics.

``tests/golden/independent_kats.json`` comes from ``tests/golden/make_independent_kats.py``: NumPy / SciPy
only, dense matrices, DOP853 at 1e-13 / matrix exponentials, gradients by fourth-order central differences.
It shares no code with ``oracle/`` or the product, so agreement here pins what the reference's own tests
cannot (they compare with QuTiP at 1e-2...5e-3 and hold no gradient value): states and expectation values
to 1e-9, parameter gradients to 1e-7 relative -- the oracle's Dormand-Prince solver at tight tolerances
(atol 1e-15, rtol 1e-14: a tolerance of 1e-12 PER STEP adds up to 1e-9 over ~1000 steps) and its Lanczos propagator against the exact ODE / exponential solution.

Four scalar parameters reach every autograd leaf class of the path: amplitude scale (drive samples),
detuning scale (detuning samples), phase chirp (complex phase of the drive), x of atom 1 (pair couplings).
"""
import math

import pytest
import torch

from helpers import KAT_SOLVER, golden, independent_problem
from oracle.ref_emulator import expect as ref_expect, total_magnetization as ref_totmag
from oracle.ref_solvers import SolverType as RefSolver
import pulser_diff_b200 as pdb
from pulser_diff_b200.utils import total_magnetization_diag

GOLD = golden("independent_kats.json")
CASES = ["K-A", "K-B", "K-C", "K-D", "K-E", "K-F", "K-G", "C1", "C2-small"]
TIGHT = {"dp5_se": dict(atol=1e-15, rtol=1e-14), "dp5_me": dict(atol=1e-15, rtol=1e-14),
         "krylov_se": dict(exp_tolerance=1e-14, norm_tolerance=1e-14)}
ATOL_STATE, RTOL_GRAD = 1e-9, 1e-7


def _nn_diag(n: int) -> torch.Tensor:
    s = torch.arange(2 ** n)
    r = [1 - ((s >> (n - 1 - q)) & 1) for q in range(n)]
    return sum(r[q] * r[q + 1] for q in range(n - 1)).to(torch.float64)


def _loss(name: str, n: int, states: torch.Tensor, sum_z: torch.Tensor) -> torch.Tensor:
    if name == "K-G":
        h = torch.tensor([[1, 1], [1, -1]], dtype=torch.complex128) / math.sqrt(2)
        return 1 - torch.abs(torch.trace(torch.kron(h, h).mH.to(states.device) @ states[-1])) / 4
    if name == "C2-small":
        return (_nn_diag(n).to(states.device)[:, None] * states[-1].abs() ** 2).sum()
    return sum_z[-1]


def _check(name: str, tsave, sum_z, final, loss, grad):
    g = GOLD[name]
    assert (tsave.detach().cpu() - torch.tensor(g["tsave"], dtype=torch.float64)).abs().max() < 1e-15
    if name != "K-G":
        assert (sum_z.detach().cpu() - torch.tensor(g["sum_z"], dtype=torch.float64)).abs().max() < ATOL_STATE
    want = torch.complex(torch.tensor(g["final_re"], dtype=torch.float64),
                         torch.tensor(g["final_im"], dtype=torch.float64)).reshape(g["final_shape"])
    assert (final.detach().cpu().reshape(want.shape) - want).abs().max() < ATOL_STATE
    assert abs(loss.item() - g["loss"]) < ATOL_STATE
    gw = torch.tensor(g["grad"], dtype=torch.float64)
    assert (grad.cpu() - gw).abs().max() < RTOL_GRAD * gw.abs().max()


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_independent_solution(name):
    theta = torch.tensor(GOLD["theta0"], dtype=torch.float64, requires_grad=True)
    p = independent_problem(name, theta)
    ref = p.ref()
    if name == "K-G":
        ref.set_initial_state(torch.eye(4))
    solver = KAT_SOLVER[name]
    st = ref.run(solver=RefSolver(solver), **TIGHT[solver]).states
    sum_z = ref_expect(ref_totmag(p.n), st).real
    loss = _loss(name, p.n, st, sum_z)
    (grad,) = torch.autograd.grad(loss, [theta])
    final = st[-1, ..., 0] if solver == "dp5_me" else st[-1]
    _check(name, ref.evaluation_times, sum_z, final, loss, grad)


@pytest.mark.parametrize("name", CASES)
def test_engine_matches_independent_solution(engine_device, name):
    theta = torch.tensor(GOLD["theta0"], dtype=torch.float64, requires_grad=True)
    p = independent_problem(name, theta)
    em = p.emulator(engine_device)
    if name == "K-G":
        em.set_initial_state(torch.eye(4))
    solver = KAT_SOLVER[name]
    res = em.run(solver=pdb.SolverType(solver), **TIGHT[solver])
    sum_z = res.expect([total_magnetization_diag(p.n)])[0].real
    loss = _loss(name, p.n, res.states, sum_z)
    (grad,) = torch.autograd.grad(loss, [theta])
    final = res.states[-1, ..., 0] if solver == "dp5_me" else res.states[-1]
    _check(name, em.evaluation_times, sum_z, final, loss, grad)
