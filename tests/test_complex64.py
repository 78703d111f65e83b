"""complex64 mode (north_star's optional 1e-5 tier): the SAME evolution with complex64 state vectors.

The reference computes in complex128 only (backend.py:271, 280), so the yardstick of this mode is the
complex128 path of this package (itself pinned to the oracle by test_parity.py): states within 1e-5,
gradients within 1e-5 of the gradient's scale.  Which library computes what:

* kets of N >= 15 and every density matrix: the complex64 build of the library (csrc compiled with -DPD_C64;
  gather kernels for any shape, stream kernels for kets of N >= 19);
* kets of N <= 14: the complex128 small-register kernels behind a cast (nothing there is bandwidth-bound).
"""
import pytest
import torch

from helpers import random_problem
import pulser_diff_b200 as pdb
from pulser_diff_b200 import _cabi, ops
from pulser_diff_b200.utils import total_magnetization_diag

C64, C128 = torch.complex64, torch.complex128


def _loss_and_grads(p, dev, solver, **options):
    """Returns (states, loss, gradients, accepted steps).  Comparisons between the two precisions use the
    shared-step protocol of test_parity.py (``replay`` = the complex128 run's accepted steps): an adaptive
    controller re-deciding its steps on float round-off moves per-sample gradients by O(solver tolerance),
    which is not what these tests measure."""
    em = p.emulator(dev)
    res = em.run(solver=solver, **options)
    f = res.expect([total_magnetization_diag(p.n)])[0].real[-1]
    leaves = [p.channels[0].amp, p.channels[0].det, p.channels[0].phase, p.coords]
    grads = torch.autograd.grad(f, leaves)
    steps = None
    if solver != pdb.SolverType.KRYLOV_SE:
        steps = [(r["dt"], r["clipped"]) for r in em._last_result.step_log() if r["accepted"]]
    return res.states.detach(), f.detach(), [g.detach().cpu() for g in grads], steps


def _close(a, b, tol):
    scale = max(float(b.abs().max()), 1e-30)
    return float((a.to(b.dtype) - b).abs().max()) / scale <= tol


def test_library_precisions(engine_device):
    assert _cabi.lib().pd_amplitude_bytes() == 16
    assert _cabi.lib(True).pd_amplitude_bytes() == 8
    assert ops.compute_dtype(C64, 12, _cabi.PD_KET) == C128        # small registers: complex128 kernels + cast
    assert ops.compute_dtype(C64, 20, _cabi.PD_KET) == C64
    assert ops.compute_dtype(C64, 5, _cabi.PD_DENSITY) == C64
    assert ops.compute_dtype(C128, 20, _cabi.PD_KET) == C128
    with pytest.raises(TypeError):
        ops.compute_dtype(torch.float64, 4, _cabi.PD_KET)
    with pytest.raises(ValueError):
        _cabi.Options.from_dict({"dtype": "float32"})
    plan = _cabi.Plan(3, 1, _cabi.PD_KET, engine_device, C64)
    with pytest.raises(TypeError):
        plan.hpsi(0.0, torch.zeros(1, 8, dtype=C128, device=engine_device))


@pytest.mark.parametrize("solver", [pdb.SolverType.DP5_SE, pdb.SolverType.KRYLOV_SE])
def test_small_ket_in_complex64_mode(engine_device, solver):
    """N <= 14: complex64 in and out, computed by the complex128 kernels."""
    ref = _loss_and_grads(random_problem(4, seed=3, T=200, rate=0.1), engine_device, solver)
    got = _loss_and_grads(random_problem(4, seed=3, T=200, rate=0.1), engine_device, solver, dtype="complex64")
    assert got[0].dtype == C64 and ref[0].dtype == C128
    assert _close(got[0], ref[0], 1e-6)
    assert abs(float(got[1] - ref[1])) < 1e-6
    for g, r in zip(got[2], ref[2]):
        assert _close(g, r, 1e-5)


@pytest.mark.parametrize("solver", [pdb.SolverType.DP5_SE, pdb.SolverType.KRYLOV_SE])
def test_ket_computed_in_complex64(engine_device, solver, monkeypatch):
    """The complex64 library itself on a ket (threshold lowered so that N = 5 takes it): states and every
    gradient class (amplitude, detuning, phase samples, atom positions) against the complex128 path."""
    opt = {} if solver == pdb.SolverType.DP5_SE else {"exp_tolerance": 1e-6, "norm_tolerance": 1e-6}
    ref = _loss_and_grads(random_problem(5, seed=5, T=200, rate=0.1), engine_device, solver)
    monkeypatch.setattr(ops, "C64_MIN_KET_QUBITS", 1)
    if ref[3] is not None:
        opt["replay"] = ref[3]
    got = _loss_and_grads(random_problem(5, seed=5, T=200, rate=0.1), engine_device, solver, dtype="complex64", **opt)
    assert got[0].dtype == C64
    assert _close(got[0], ref[0], 1e-5)
    assert abs(float(got[1] - ref[1])) < 1e-5
    for g, r in zip(got[2], ref[2]):
        assert _close(g, r, 2e-5)


def test_density_computed_in_complex64(engine_device):
    """Lindblad path: vec(rho) in complex64 (the complex64 library's gather kernels)."""
    noise = {"dephasing_rate": 0.5, "relaxation_rate": 0.1}
    ref = _loss_and_grads(random_problem(3, seed=7, T=200, rate=0.1, noise=noise), engine_device, pdb.SolverType.DP5_ME)
    got = _loss_and_grads(random_problem(3, seed=7, T=200, rate=0.1, noise=noise), engine_device, pdb.SolverType.DP5_ME,
                          dtype="complex64", replay=ref[3])
    assert got[0].dtype == C64
    assert _close(got[0], ref[0], 1e-5)
    for g, r in zip(got[2], ref[2]):
        assert _close(g, r, 2e-5)


def _chain_plan(n, dev, dtype, path=0, batch=1):
    """A plan of the N-atom chain with a global drive (phase != 0: the general flip arithmetic) and one
    local detuning, as the large-register tests of test_gpu_scale.py build it."""
    plan = _cabi.Plan(n, batch, _cabi.PD_KET, dev, dtype)
    x = torch.arange(n, dtype=torch.float64) * 6.0
    r = (x[:, None] - x[None, :]).abs() + torch.eye(n, dtype=torch.float64)
    u = torch.triu(865723.02 / r ** 6, diagonal=1)
    plan.set_interaction(u)
    ns = 8
    k = torch.arange(ns, dtype=torch.float64)
    full = (1 << n) - 1
    det = torch.stack([-0.5 * (1.0 + 0.3 * k), -0.5 * (0.4 - 0.1 * k)])
    amp = torch.stack([0.5 * (2.0 + 0.2 * k) * torch.exp(-1j * (0.3 + 0.05 * k)), 0.5 * (0.7 + 0.0 * k) + 0j])
    plan.set_terms(0.05, [full, 1 << (n // 2)], det, [full, 1 << 1], amp.to(C128))
    plan.set_path(path)
    return plan


@pytest.mark.gpu
@pytest.mark.parametrize("n,path,batch", [(15, 0, 1), (17, 1, 3), (17, 4, 1), (19, 0, 2), (21, 0, 1)])
def test_hpsi_complex64_families(cuda_device, n, path, batch):
    """One H(t) psi per kernel family of the complex64 library (gather: N = 15, 17; stream: N = 17 on request,
    19 and 21 by default; batches of columns on both) against the complex128 library on the same vectors."""
    g = torch.Generator(device="cpu").manual_seed(n)
    psi = torch.randn(batch, 2 ** n, dtype=C128, generator=g)
    psi = (psi / psi.norm()).to(cuda_device)
    ref = _chain_plan(n, cuda_device, C128, 0, batch).hpsi(0.11, psi)
    got = _chain_plan(n, cuda_device, C64, path, batch).hpsi(0.11, psi.to(C64))
    assert got.dtype == C64
    assert _close(got, ref, 2e-6)           # one application: float round-off of ~N terms


@pytest.mark.gpu
@pytest.mark.parametrize("n", [16, 19, 20])
def test_evolution_and_gradient_complex64(cuda_device, n):
    """DP5 forward + adjoint through the fused step kernels of each family in complex64 (gather at N = 16,
    stream at N = 19, 20) against complex128 on the complex128 run's accepted steps, plus one free-running
    complex64 evolution whose own controller must land within the solver tolerance."""
    def run(dtype, opt):
        plan = _chain_plan(n, cuda_device, dtype)
        psi0 = torch.zeros(1, 2 ** n, dtype=dtype, device=cuda_device)
        psi0[0, -1] = 1.0
        ts = torch.tensor([0.0, 0.12, 0.3], dtype=torch.float64)
        states, tape = plan.evolve_forward(_cabi.SOLVER_DP5_SE, opt, psi0, ts, want_tape=True)
        obs = (torch.arange(2 ** n, device=cuda_device) & 1).to(torch.float64)      # n of the last atom
        gs = torch.zeros_like(states)
        gs[-1] = 2.0 * obs * states[-1]
        g_det, g_amp, _, g_ts, g_s0 = plan.evolve_backward(tape, states, gs, True, True, False, True, True)
        return states, g_det, g_amp, g_ts, g_s0, tape.records()

    ref = run(C128, _cabi.Options())
    frozen = _cabi.Options(replay=[(r["dt"], r["clipped"]) for r in ref[5] if r["accepted"]])
    got = run(C64, frozen)
    assert got[0].dtype == C64 and got[4].dtype == C64
    assert _close(got[0], ref[0], 1e-5)
    for k in (1, 2, 3):
        assert _close(got[k], ref[k], 2e-5), k
    assert _close(got[4], ref[4], 1e-5)
    free = run(C64, _cabi.Options())
    assert _close(free[0], ref[0], 1e-4)      # its own step sequence: solver tolerance (rtol 1e-6 per step)
    assert abs(float(free[0][-1].norm()) - 1.0) < 1e-4
