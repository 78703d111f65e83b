"""GPU-only checks at sizes the oracle cannot reach: size-independent properties.

* unitarity (norm conservation) and linearity of the evolution,
* H Hermitian: <a|H b> == conj(<b|H a>),
* gather kernels vs tiled kernels produce the same vectors,
* forward + adjoint gradient against a central finite difference of the same engine.
"""
import pytest
import torch

import pulser_diff_b200 as pdb
from pulser_diff_b200 import _cabi, ops

pytestmark = pytest.mark.gpu


def _program(n, T=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    dv = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4
    av = torch.complex(torch.rand(1, T, dtype=torch.float64, generator=g) * 3,
                       torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5)
    x = torch.arange(n, dtype=torch.float64) * 7.0
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = 865723.02 / float(abs(x[i] - x[j])) ** 6
    full = (1 << n) - 1
    return dict(dt=0.002, det_masks=[full], det_values=dv, amp_masks=[full], amp_values=av, pair_u=u)


@pytest.mark.parametrize("n", [10, 16, 20])
def test_hermitian_and_linear(cuda_device, n):
    pr = _program(n)
    dev = cuda_device
    g = torch.Generator().manual_seed(n)
    a = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=g).to(dev)
    b = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=g).to(dev)
    args = (0.0371, pr["det_values"], pr["amp_values"], pr["pair_u"], pr["det_masks"], pr["amp_masks"], pr["dt"])
    Ha = torch.ops.pulser_diff_b200.hpsi(a, *args)
    Hb = torch.ops.pulser_diff_b200.hpsi(b, *args)
    lhs = torch.vdot(a.flatten(), Hb.flatten())
    rhs = torch.vdot(b.flatten(), Ha.flatten()).conj()
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)
    Hab = torch.ops.pulser_diff_b200.hpsi(a + 2j * b, *args)
    assert (Hab - (Ha + 2j * Hb)).abs().max() < 1e-10 * Hab.abs().max()


@pytest.mark.parametrize("n", [14, 20])
def test_norm_conservation_and_fd_gradient(cuda_device, n):
    """Adjoint gradient vs central finite differences of the SAME engine on a frozen step
    sequence (replay), so the differenced function is smooth (an adaptive controller re-deciding
    its steps under the perturbation adds ~tolerance/eps noise to a finite difference)."""
    pr = _program(n, T=16)
    dev = cuda_device
    psi0 = torch.zeros(1, 2 ** n, dtype=torch.complex128, device=dev)
    psi0[0, -1] = 1.0
    tsave = torch.tensor([0.0, 0.012, 0.024], dtype=torch.float64)
    av = pr["amp_values"].clone().requires_grad_(True)
    obs = torch.arange(2 ** n, device=dev).remainder(7).to(torch.float64)

    def f(av_, opt):
        st = ops.evolve(psi0, tsave, pr["det_values"], av_, pr["pair_u"], n_qubits=n,
                        kind=_cabi.PD_KET, dt=pr["dt"], det_masks=pr["det_masks"],
                        amp_masks=pr["amp_masks"], options=opt)
        return st, (obs * st[-1, 0].abs() ** 2).sum()

    st, val = f(av, _cabi.Options(atol=1e-12, rtol=1e-10))
    assert (st.detach().norm(dim=-1) - 1).abs().max() < 1e-8
    log = ops.last_step_log(st)
    frozen = _cabi.Options(replay=[(r["dt"], r["clipped"]) for r in log if r["accepted"]])
    st2, val2 = f(av, frozen)
    assert (st2.detach() - st.detach()).abs().max() < 1e-13
    (g,) = torch.autograd.grad(val2, [av])
    eps = 1e-4
    for idx in (3, 9):
        d = torch.zeros_like(av.detach())
        d[0, idx] = eps
        fd = (f(av.detach() + d, frozen)[1] - f(av.detach() - d, frozen)[1]) / (2 * eps)
        assert abs(fd.item() - g[0, idx].real.item()) < 1e-6 * abs(fd.item()) + 1e-9


def _vary_drive(pr, n, drive):
    """"complex": one global drive with a phase (general flip arithmetic); "real": phase-free global
    drive (the two-FMA flip path); "local": global terms plus per-qubit drives / detunings on a few
    qubits spread over the low, middle and high bit groups (non-uniform coefficients)."""
    if drive == "real":
        pr["amp_values"] = torch.complex(pr["amp_values"].real, torch.zeros_like(pr["amp_values"].real))
    elif drive == "local":
        g = torch.Generator().manual_seed(9)
        T = pr["det_values"].shape[1]
        qs = [0, n // 2, n - 1]
        pr["det_masks"] = pr["det_masks"] + [1 << q for q in qs]
        pr["amp_masks"] = pr["amp_masks"] + [1 << q for q in qs]
        pr["det_values"] = torch.cat([pr["det_values"], torch.rand(3, T, dtype=torch.float64, generator=g)])
        pr["amp_values"] = torch.cat([pr["amp_values"],
                                      torch.complex(torch.rand(3, T, dtype=torch.float64, generator=g),
                                                    torch.rand(3, T, dtype=torch.float64, generator=g) - 0.5)])
    return pr


@pytest.mark.parametrize("n,other,drive", [(16, 2, "complex"), (18, 2, "complex"), (21, 2, "complex"),
                                           (22, 2, "complex"), (23, 2, "complex"),
                                           (16, 4, "complex"), (19, 4, "complex"), (21, 4, "complex"),
                                           (24, 4, "complex"), (26, 4, "complex"),
                                           (16, 4, "real"), (17, 4, "local"), (20, 4, "real"), (21, 4, "local"),
                                           (22, 4, "real"), (24, 4, "local"), (25, 4, "real")])
def test_tiled_equals_gather(cuda_device, n, other, drive):
    """The tiled kernels (path 2: fused DP5 step, two-launch stage, adjoint sweep) and the stream
    kernels (path 4: one bit-group of H per launch, A + first group as one L2-blocked dataflow launch)
    against the gather kernels (path 1) on the same inputs: states, H.psi and gradients."""
    pr = _vary_drive(_program(n, T=16), n, drive)
    dev = cuda_device
    nb = 2 if n <= 23 else 1
    psi0 = torch.randn(nb, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(5)).to(dev)
    psi0 /= psi0.norm(dim=1, keepdim=True)
    tsave = torch.tensor([0.0, 0.004, 0.008], dtype=torch.float64)
    outs, grads, hp = [], [], []
    for path in (1, other):
        av = pr["amp_values"].clone().requires_grad_(True)
        dv = pr["det_values"].clone().requires_grad_(True)
        pu = pr["pair_u"].clone().requires_grad_(True)
        st = ops.evolve(psi0, tsave, dv, av, pu, n_qubits=n, kind=_cabi.PD_KET,
                        dt=pr["dt"], det_masks=pr["det_masks"], amp_masks=pr["amp_masks"],
                        options=_cabi.Options(path=path))
        outs.append(st.detach())
        w = torch.arange(2 ** n, device=dev).remainder(5).to(torch.float64)
        val = (w * st[-1].abs() ** 2).sum() + (w * st[1].abs() ** 2).sum()
        grads.append(torch.autograd.grad(val, [av, dv, pu]))
        plan = ops.get_plan(n, nb, _cabi.PD_KET, dev)
        plan.set_path(path)
        hp.append(plan.hpsi(0.0051, psi0))
        plan.set_path(0)
    assert (outs[0] - outs[1]).abs().max() < 1e-12
    assert (hp[0] - hp[1]).abs().max() < 1e-12 * hp[0].abs().max()
    for a, b in zip(grads[0], grads[1]):
        assert (a - b).abs().max() < 1e-9 * max(1e-30, b.abs().max().item())


@pytest.mark.parametrize("n,batch,local", [(3, 1, False), (6, 4, True), (10, 1, False), (12, 1, True),
                                           (13, 1, False)])
def test_small_register_kernels_equal_gather(cuda_device, n, batch, local):
    """One-launch cooperative kernels (csrc/small_ket.cuh, path 3) against the stage-by-stage gather
    kernels (path 1): states, attempted-step log, and every gradient (samples, times, pair
    couplings, initial state)."""
    pr = _program(n, T=16, seed=3)
    dev = cuda_device
    if local:   # per-qubit drive and detuning terms on two qubits + the global terms
        g = torch.Generator().manual_seed(9)
        pr["det_masks"] = pr["det_masks"] + [1 << 0, 1 << (n - 1)]
        pr["amp_masks"] = pr["amp_masks"] + [1 << 1]
        pr["det_values"] = torch.cat([pr["det_values"], torch.rand(2, 16, dtype=torch.float64, generator=g)])
        pr["amp_values"] = torch.cat([pr["amp_values"],
                                      torch.complex(torch.rand(1, 16, dtype=torch.float64, generator=g),
                                                    torch.rand(1, 16, dtype=torch.float64, generator=g))])
    psi0 = torch.randn(batch, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(5)).to(dev)
    psi0 /= psi0.norm(dim=1, keepdim=True)
    tsave0 = torch.tensor([0.0, 0.004, 0.0095, 0.013], dtype=torch.float64)
    def run(path, replay=None):
        av = pr["amp_values"].clone().requires_grad_(True)
        dv = pr["det_values"].clone().requires_grad_(True)
        pu = pr["pair_u"].clone().requires_grad_(True)
        ts = tsave0.clone().requires_grad_(True)
        p0 = psi0.clone().requires_grad_(True)
        st = ops.evolve(p0, ts, dv, av, pu, n_qubits=n, kind=_cabi.PD_KET, dt=pr["dt"],
                        det_masks=pr["det_masks"], amp_masks=pr["amp_masks"],
                        options=_cabi.Options(path=path, replay=replay))
        w = torch.arange(2 ** n, device=dev).remainder(5).to(torch.float64)
        val = (w * st[-1].abs() ** 2).sum() + (w * st[1].abs() ** 2).sum() + (w * st[2].real).sum()
        return st.detach(), ops.last_step_log(st), torch.autograd.grad(val, [av, dv, pu, ts, p0])

    # free-running controllers: same decisions, step sizes equal up to the rounding of the
    # error norm (a cancellation-prone quantity entering as error^(-1/5))
    st_g, log_g, _ = run(1)
    st_s, log_s, _ = run(3)
    assert len(log_g) == len(log_s) and len(log_g) > 3
    for a, b in zip(log_g, log_s):
        assert a["accepted"] == b["accepted"] and a["clipped"] == b["clipped"]
        assert abs(a["dt"] - b["dt"]) <= 1e-6 * abs(b["dt"])
    assert (st_g - st_s).abs().max() < 1e-9
    # shared step sequence: round-off agreement of states and of every gradient
    frozen = [(r["dt"], r["clipped"]) for r in log_g if r["accepted"]]
    st_g, _, g_g = run(1, frozen)
    st_s, _, g_s = run(3, frozen)
    assert (st_g - st_s).abs().max() < 1e-12
    for a, b in zip(g_g, g_s):
        assert (a - b).abs().max() < 1e-9 * max(1e-30, b.abs().max().item())


def test_c2_workload_vs_oracle_first_intervals(cuda_device):
    """BASELINE configs[1] at FULL size (12-atom chain, the bench workload): the first two tsave
    intervals against the oracle on the same seeded inputs -- states to 1e-10 and the gradient
    w.r.t. the 60 pulse parameters to 1e-8 relative (shared accepted-step sequence), plus the
    free-running controller taking the same decisions."""
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench as B
    from helpers import Channel, Problem
    from oracle.ref_solvers import SolverType as RefSolver, sesolve as ref_sesolve
    from pulser_diff_b200.utils import interpolate_sine

    dev = cuda_device
    interp = interpolate_sine(B.N_PARAM, B.DURATION).to(torch.float64)
    ta, td = B.workload_params(0)
    amp, det, ph = B.pulse_samples(ta, td, interp)
    p = Problem(B.chain_coords(B.N_QUBITS), B.C6, [Channel(amp, det, ph)], rate=B.RATE)
    ref = p.ref()
    ts = ref.evaluation_times[:3].clone()
    r = ref_sesolve(ref.ham.H, ref.initial_state, ts, RefSolver.DP5_SE, {})
    d = B.loss_diag(B.N_QUBITS, "cpu")
    loss_r = (d[:, None] * r.states[-1].abs() ** 2).sum()
    g_ref = torch.autograd.grad(loss_r, [ta, td], retain_graph=True)

    em = p.emulator(dev)
    H = em._hamiltonian._hamiltonian
    accepted = [rec for rec in r.steplog if rec[2]]
    res = pdb.sesolve(H, em.initial_state, ts, pdb.SolverType.DP5_SE,
                      options={"replay": [(dt, clipped) for (_, dt, _, _, clipped) in accepted]})
    assert (res.states.detach().cpu() - r.states.detach()).abs().max() < 1e-10
    loss = (d.to(dev)[:, None] * res.states[-1].abs() ** 2).sum()
    assert abs(loss.item() - loss_r.item()) < 1e-10
    g = torch.autograd.grad(loss, [ta, td], retain_graph=True)
    for a, b in zip(g, g_ref):
        assert (a.cpu() - b).abs().max() < 1e-8 * b.abs().max()
    # free-running controller: same accept/reject pattern as the oracle's
    res2 = pdb.sesolve(H, em.initial_state, ts, pdb.SolverType.DP5_SE)
    log = res2.step_log()
    assert [bool(a["accepted"]) for a in log] == [bool(b[2]) for b in r.steplog]
    assert (res2.states.detach().cpu() - r.states.detach()).abs().max() < 1e-7


@pytest.mark.parametrize("n", [16, 18])
def test_large_register_families_against_oracle_sparse_h(cuda_device, n):
    """Every kernel family of the large registers DIRECTLY against the oracle's sparse-COO H(t)
    (reference hamiltonian.py:526-546 restated), not only against each other: H(t) psi at three times for
    the gather (path 1), tiled (path 2) and stream (path 4) kernels, with a global drive that carries a
    phase plus local drives / detunings on qubits of the low, middle and top bit groups; and, at N = 16, a
    short DP5 evolution on the oracle's step sequence."""
    from helpers import C6_60, Channel, Problem, chain
    dev = cuda_device
    T = 40
    g = torch.Generator().manual_seed(n)
    r = lambda: torch.rand(T, dtype=torch.float64, generator=g)
    chans = [Channel(3 * r(), r() - 0.5, r())] + [Channel(r(), 2 * r() - 1, r(), "Local", q) for q in (1, n // 2, n - 2)]
    p = Problem(chain(n), C6_60, chans, rate=1.0)
    ref = p.ref()
    em = p.emulator(dev)
    H = em._hamiltonian._hamiltonian
    dm, dv, am, av = H.masks_and_values()
    psi = torch.randn(2 ** n, 1, dtype=torch.complex128, generator=g)
    psi_d = psi.T.contiguous().to(dev)
    plan = ops.get_plan(n, 1, _cabi.PD_KET, dev)
    for t in (0.0, 0.0123, 0.0377):
        want = (ref.ham.H(torch.tensor(t, dtype=torch.float64)) @ psi).T
        for path in (1, 2, 4):
            plan.set_path(path)
            got = torch.ops.pulser_diff_b200.hpsi(psi_d, t, dv, av, H.pair_u.detach(), dm, am, H.dt).cpu()
            assert (got - want).abs().max() < 1e-12 * want.abs().max(), (path, t)
    plan.set_path(0)
    if n > 16:
        return
    from oracle.ref_solvers import SolverType as RefSolver, sesolve as ref_sesolve
    psi0 = psi / psi.norm()
    ts = torch.tensor([0.0, 0.004, 0.009], dtype=torch.float64)
    rr = ref_sesolve(ref.ham.H, psi0, ts, RefSolver.DP5_SE, {})
    replay = [(dt, clipped) for (_, dt, acc, _, clipped) in rr.steplog if acc]
    for path in (2, 4):
        res = pdb.sesolve(H, psi0, ts, pdb.SolverType.DP5_SE, options={"replay": replay, "path": path})
        assert (res.states.detach().cpu() - rr.states).abs().max() < 1e-12, path


def test_c2_full_workload_against_oracle_fixture(cuda_device):
    """ALL 55 tsave intervals of the C2 workload (12 atoms, 406 attempted / 223 accepted DP5 steps) against
    the oracle's full run, stored by tests/golden/make_c2_full.py (the oracle's tape gradient alone takes
    five minutes on the CPU): loss at every evaluation time and the final state on the shared step
    sequence (1e-10), the gradient w.r.t. the 60 pulse parameters (1e-8 relative), and the free-running
    controller's accept/reject log."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench as B
    from helpers import Channel, Problem, golden
    from pulser_diff_b200.utils import interpolate_sine

    gold = golden("c2_full.json")
    dev = cuda_device
    interp = interpolate_sine(B.N_PARAM, B.DURATION).to(torch.float64)
    ta, td = B.workload_params(0)
    amp, det, ph = B.pulse_samples(ta, td, interp)
    p = Problem(B.chain_coords(B.N_QUBITS), B.C6, [Channel(amp, det, ph)], rate=B.RATE)
    em = p.emulator(dev)
    assert (em.evaluation_times.detach().cpu() - torch.tensor(gold["tsave"], dtype=torch.float64)).abs().max() < 1e-15
    H = em._hamiltonian._hamiltonian
    accepted = [rec for rec in gold["steplog"] if rec[2]]
    res = pdb.sesolve(H, em.initial_state, em.evaluation_times, pdb.SolverType.DP5_SE,
                      options={"replay": [(dt, clipped) for (_, dt, _, _, clipped) in accepted]})
    d = B.loss_diag(B.N_QUBITS, dev)
    series = (d[None, :, None] * res.states.abs() ** 2).sum(dim=(1, 2))
    assert (series.detach().cpu() - torch.tensor(gold["loss_series"], dtype=torch.float64)).abs().max() < 1e-10
    final = torch.complex(torch.tensor(gold["final_re"], dtype=torch.float64),
                          torch.tensor(gold["final_im"], dtype=torch.float64))
    assert (res.states[-1].detach().cpu().reshape(-1) - final).abs().max() < 1e-10
    ga, gd = torch.autograd.grad(series[-1], [ta, td])
    for got, want in ((ga, gold["grad_ta"]), (gd, gold["grad_td"])):
        want = torch.tensor(want, dtype=torch.float64)
        assert (got.cpu() - want).abs().max() < 1e-8 * want.abs().max()
    res2 = pdb.sesolve(H, em.initial_state, em.evaluation_times, pdb.SolverType.DP5_SE)
    log = res2.step_log()
    assert [bool(a["accepted"]) for a in log] == [bool(b[2]) for b in gold["steplog"]]
    assert (res2.states[-1].detach().cpu().reshape(-1) - final).abs().max() < 1e-7


def test_small_tape_overwritten_falls_back(cuda_device):
    """Two evolutions on the same plan, then the gradient of the FIRST: its device-side stage tape
    has been overwritten by the second, so the adjoint must fall back to the recomputing sweep and
    still return the same gradient as an undisturbed run."""
    n = 10
    pr = _program(n, T=16, seed=5)
    dev = cuda_device
    psi0 = torch.zeros(1, 2 ** n, dtype=torch.complex128, device=dev)
    psi0[0, -1] = 1.0
    tsave = torch.tensor([0.0, 0.005, 0.011], dtype=torch.float64)
    w = torch.arange(2 ** n, device=dev).remainder(3).to(torch.float64)

    def run(av):
        st = ops.evolve(psi0, tsave, pr["det_values"], av, pr["pair_u"], n_qubits=n, kind=_cabi.PD_KET,
                        dt=pr["dt"], det_masks=pr["det_masks"], amp_masks=pr["amp_masks"])
        return st, (w * st[-1, 0].abs() ** 2).sum()

    av1 = pr["amp_values"].clone().requires_grad_(True)
    _, val = run(av1)
    (g_clean,) = torch.autograd.grad(val, [av1])
    av2 = pr["amp_values"].clone().requires_grad_(True)
    _, val2 = run(av2)
    av3 = (pr["amp_values"] * 1.3).clone().requires_grad_(True)
    run(av3)                                   # overwrites the plan's stage tape
    (g_late,) = torch.autograd.grad(val2, [av2])
    assert (g_late - g_clean).abs().max() < 1e-9 * g_clean.abs().max()


@pytest.mark.parametrize("n,batch", [(15, 1), (16, 1), (17, 2)])
def test_family_boundaries_norm_and_replay(cuda_device, n, batch):
    """Register sizes between the kernel families (small-register <= 14 < gather < tiled >= 18):
    unitarity and agreement of the automatic choice with the forced gather kernels."""
    pr = _program(n, T=16)
    dev = cuda_device
    psi0 = torch.randn(batch, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(1)).to(dev)
    psi0 /= psi0.norm(dim=1, keepdim=True)
    tsave = torch.tensor([0.0, 0.004, 0.009], dtype=torch.float64)
    outs = []
    for path in (0, 1):
        st = ops.evolve(psi0, tsave, pr["det_values"], pr["amp_values"], pr["pair_u"], n_qubits=n,
                        kind=_cabi.PD_KET, dt=pr["dt"], det_masks=pr["det_masks"], amp_masks=pr["amp_masks"],
                        options=_cabi.Options(path=path))
        outs.append(st)
    assert (outs[0].norm(dim=-1) - 1).abs().max() < 1e-6
    assert (outs[0] - outs[1]).abs().max() < 1e-9


@pytest.mark.parametrize("n,other", [(18, 2), (20, 4)])
def test_krylov_on_tiled_and_stream_kernels(cuda_device, n, other):
    """KRYLOV_SE (Lanczos propagation, adjoint by Krylov-space quadrature) runs its H.psi through the
    same kernel families: tiled (path 2) and stream (path 4) against gather (path 1)."""
    pr = _program(n, T=16)
    dev = cuda_device
    psi0 = torch.zeros(1, 2 ** n, dtype=torch.complex128, device=dev)
    psi0[0, -1] = 1.0
    tsave = torch.tensor([0.0, 0.003, 0.006], dtype=torch.float64)
    w = torch.arange(2 ** n, device=dev).remainder(5).to(torch.float64)
    outs, grads = [], []
    for path in (1, other):
        av = pr["amp_values"].clone().requires_grad_(True)
        dv = pr["det_values"].clone().requires_grad_(True)
        st = ops.evolve(psi0, tsave, dv, av, pr["pair_u"], n_qubits=n, kind=_cabi.PD_KET, dt=pr["dt"],
                        det_masks=pr["det_masks"], amp_masks=pr["amp_masks"], solver=_cabi.SOLVER_KRYLOV_SE,
                        options=_cabi.Options(path=path))
        outs.append(st.detach())
        val = (w * st[-1, 0].abs() ** 2).sum()
        grads.append(torch.autograd.grad(val, [av, dv]))
    assert (outs[0].norm(dim=-1) - 1).abs().max() < 1e-9
    assert (outs[0] - outs[1]).abs().max() < 1e-10
    for a, b in zip(grads[0], grads[1]):
        assert (a - b).abs().max() < 1e-8 * max(1e-30, b.abs().max().item())


def test_lindblad_midsize_invariants_and_fd_gradient(cuda_device):
    """DP5_ME at N = 7 (4^7 entries, beyond what the dense-operator oracle does in seconds): trace and
    hermiticity are conserved, the purity decays, and the adjoint gradient w.r.t. a pulse sample equals
    a central finite difference on the frozen step sequence."""
    from pulser_diff_b200.utils import occupation_diag
    n, T = 7, 24
    dev = cuda_device
    pr = _program(n, T=T, seed=8)
    col = torch.tensor([[[0.5, 0], [0, -0.5]], [[0, 0], [0.3, 0]]], dtype=torch.complex128)   # sqrt(g/2) Z, sqrt(g) |g><r|
    rho0 = torch.zeros(1, 4 ** n, dtype=torch.complex128, device=dev)
    rho0[0, -1] = 1.0
    tsave = torch.tensor([0.0, 0.01, 0.02], dtype=torch.float64)
    obs = torch.zeros(2 ** n, dtype=torch.float64, device=dev)
    for i in range(n):
        obs = obs + occupation_diag(n, [i], dev)

    def f(av_, opt):
        st = ops.evolve(rho0, tsave, pr["det_values"], av_, pr["pair_u"], n_qubits=n, kind=_cabi.PD_DENSITY,
                        dt=pr["dt"], det_masks=pr["det_masks"], amp_masks=pr["amp_masks"], collapse=col,
                        solver=_cabi.SOLVER_DP5_ME, options=opt)
        rho = st[-1, 0].reshape(2 ** n, 2 ** n)
        return st, (obs * rho.diagonal().real).sum()

    av = pr["amp_values"].clone().requires_grad_(True)
    st, val = f(av, _cabi.Options(atol=1e-11, rtol=1e-9))
    rho = st.detach()[-1, 0].reshape(2 ** n, 2 ** n)
    assert abs(torch.trace(rho).item() - 1) < 1e-9
    assert (rho - rho.mH).abs().max() < 1e-12
    assert torch.trace(rho @ rho).real.item() < 1 - 1e-6
    log = ops.last_step_log(st)
    frozen = _cabi.Options(replay=[(r["dt"], r["clipped"]) for r in log if r["accepted"]])
    _, val2 = f(av, frozen)
    (g,) = torch.autograd.grad(val2, [av])
    eps = 1e-4
    for idx in (2, 7):
        d = torch.zeros_like(av.detach())
        d[0, idx] = eps
        fd = (f(av.detach() + d, frozen)[1] - f(av.detach() - d, frozen)[1]) / (2 * eps)
        assert abs(fd.item() - g[0, idx].real.item()) < 1e-6 * abs(fd.item()) + 1e-9


def test_c4_full_size_lindblad_invariants_and_fd_gradient(cuda_device):
    """BASELINE configs[3] at its full size (N = 12: rho has 4^12 entries, 256 MiB per vector; the
    dense-operator oracle cannot run this).  Size-independent properties of the Lindblad flow with
    dephasing + amplitude damping: trace and hermiticity conserved, purity decays from 1, total
    Rydberg population stays in [0, N]; and the adjoint gradient w.r.t. one drive sample and one
    detuning sample equals a central difference on the frozen step sequence (1e-6 relative)."""
    from pulser_diff_b200.utils import occupation_diag
    n, T = 12, 64
    dev = cuda_device
    pr = _program(n, T=T, seed=12)
    g_deph, g_damp = 0.5, 0.1
    col = torch.tensor([[[(g_deph / 2) ** 0.5, 0], [0, -(g_deph / 2) ** 0.5]],
                        [[0, 0], [g_damp ** 0.5, 0]]], dtype=torch.complex128)
    rho0 = torch.zeros(1, 4 ** n, dtype=torch.complex128, device=dev)
    rho0[0, -1] = 1.0
    tsave = torch.tensor([0.0, 0.05], dtype=torch.float64)
    obs = torch.zeros(2 ** n, dtype=torch.float64, device=dev)
    for i in range(n):
        obs = obs + occupation_diag(n, [i], dev)

    def f(av_, dv_, opt):
        st = ops.evolve(rho0, tsave, dv_, av_, pr["pair_u"], n_qubits=n, kind=_cabi.PD_DENSITY,
                        dt=pr["dt"], det_masks=pr["det_masks"], amp_masks=pr["amp_masks"], collapse=col,
                        solver=_cabi.SOLVER_DP5_ME, options=opt)
        rho = st[-1, 0].reshape(2 ** n, 2 ** n)
        return st, (obs * rho.diagonal().real).sum()

    av = pr["amp_values"].clone().requires_grad_(True)
    dv = pr["det_values"].clone().requires_grad_(True)
    st, val = f(av, dv, _cabi.Options(atol=1e-10, rtol=1e-8))
    rho = st.detach()[-1, 0].reshape(2 ** n, 2 ** n)
    assert abs(rho.diagonal().sum().item() - 1) < 1e-9
    assert (rho - rho.mH).abs().max() < 1e-12
    purity = (rho.abs() ** 2).sum().item()
    assert 0.0 < purity < 1 - 1e-4
    assert 0.0 < val.item() < n
    log = [r for r in ops.last_step_log(st) if r["accepted"]]
    assert len(log) >= 2
    frozen = _cabi.Options(replay=[(r["dt"], r["clipped"]) for r in log])
    del st, rho
    _, val2 = f(av, dv, frozen)
    g_av, g_dv = torch.autograd.grad(val2, [av, dv])
    eps = 1e-4
    d = torch.zeros_like(av.detach()); d[0, 9] = eps
    fd = (f(av.detach() + d, dv.detach(), frozen)[1] - f(av.detach() - d, dv.detach(), frozen)[1]) / (2 * eps)
    assert abs(fd.item() - g_av[0, 9].real.item()) < 1e-6 * abs(fd.item()) + 1e-9
    d = torch.zeros_like(dv.detach()); d[0, 14] = eps
    fd = (f(av.detach(), dv.detach() + d, frozen)[1] - f(av.detach(), dv.detach() - d, frozen)[1]) / (2 * eps)
    assert abs(fd.item() - g_dv[0, 14].item()) < 1e-6 * abs(fd.item()) + 1e-9


@pytest.mark.parametrize("n,noise", [(8, "both"), (9, "both"), (10, "dephasing"), (10, "both"), (11, "both")])
def test_density_tiles_equal_gather(cuda_device, n, noise):
    """The opt-in density tile kernels (path 5: both bits of up to six sites closed per launch, two of them in
    registers; dens_tile.cu) against the gather kernel (path 1) on the same inputs: DP5_ME states and gradients with a
    phase-carrying drive.  "dephasing": diagonal collapse operator only (no double flips); "both": dephasing +
    relaxation (every entry of the 4x4 site super-operators in use)."""
    dev = cuda_device
    pr = _program(n, T=16, seed=3)
    col = torch.tensor([[[0.5, 0], [0, -0.5]], [[0, 0], [0.3, 0]]], dtype=torch.complex128)
    if noise == "dephasing":
        col = col[:1]
    g = torch.Generator().manual_seed(n)
    a = torch.randn(2 ** n, 2 ** n, dtype=torch.complex128, generator=g)
    rho = a @ a.mH
    rho = (rho / torch.trace(rho)).reshape(1, 4 ** n).to(dev)
    tsave = torch.tensor([0.0, 0.003, 0.006], dtype=torch.float64)
    outs, grads = [], []
    for path in (1, 5):
        av = pr["amp_values"].clone().requires_grad_(True)
        dv = pr["det_values"].clone().requires_grad_(True)
        st = ops.evolve(rho, tsave, dv, av, pr["pair_u"], n_qubits=n, kind=_cabi.PD_DENSITY, dt=pr["dt"],
                        det_masks=pr["det_masks"], amp_masks=pr["amp_masks"], collapse=col,
                        solver=_cabi.SOLVER_DP5_ME, options=_cabi.Options(path=path))
        outs.append(st.detach())
        w = torch.arange(4 ** n, device=dev).remainder(7).to(torch.float64)
        val = (w * st[-1].real).sum() + (w * st[1].imag).sum()
        grads.append(torch.autograd.grad(val, [av, dv]))
    assert (outs[0] - outs[1]).abs().max() < 1e-13
    for x, y in zip(grads[0], grads[1]):
        assert (x - y).abs().max() < 1e-9 * max(1e-30, y.abs().max().item())
