"""world_size > 1 paths on CPU: gloo backend, host stand-in library (tests/emu).

* axis 1 (independent parameter sets): sharding is a partition, results gather in unit order,
  and the sharded evaluation equals the serial one.
* axis 2 (one register sharded by its top qubits): ShardedKet.hpsi == single-process H.psi.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libpd_emu.so")
EMU_C64 = os.path.join(ROOT, "tests", "emu", "libpd_emu_c64.so")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _program(n, T=24, seed=0):
    g = torch.Generator().manual_seed(seed)
    full = (1 << n) - 1
    dv = torch.stack([(torch.rand(T, dtype=torch.float64, generator=g) - 0.5) * 4,
                      torch.rand(T, dtype=torch.float64, generator=g)])
    av = torch.stack([torch.complex(torch.rand(T, dtype=torch.float64, generator=g) * 3,
                                    torch.rand(T, dtype=torch.float64, generator=g) - 0.5),
                      torch.complex(torch.rand(T, dtype=torch.float64, generator=g),
                                    torch.rand(T, dtype=torch.float64, generator=g))])
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = 40.0 / (j - i) ** 6
    # one global term + one local term on qubit 0 (a GLOBAL/shard qubit) for det and amp
    return dict(dt=0.002, det_masks=[full, 1], det_values=dv, amp_masks=[full, 1 << 1], amp_values=av,
                pair_u=u)


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    from pulser_diff_b200 import _cabi, ops, parallel
    _cabi.use_library(EMU)
    dev = torch.device("cpu")
    pr = _program(n)
    # ---- axis 2 -------------------------------------------------------------------------
    psi = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(3))
    plan = ops.get_plan(n, 1, _cabi.PD_KET, dev)
    ops.configure(plan, ops.make_program(n, _cabi.PD_KET, pr["dt"], pr["det_masks"], pr["det_values"],
                                         pr["amp_masks"], pr["amp_values"], pr["pair_u"], None))
    sk = parallel.ShardedKet(n, pr["pair_u"], pr["dt"], pr["det_masks"], pr["det_values"],
                             pr["amp_masks"], pr["amp_values"], dev)
    errs = []
    for t in (0.0, 0.0131, 0.0377):
        full = plan.hpsi(t, psi)
        mine = sk.hpsi(t, sk.local_slice(psi))
        errs.append((mine - sk.local_slice(full)).abs().max().item() / full.abs().max().item())
    # ---- axis 1 -------------------------------------------------------------------------
    n_units = 7
    mine = parallel.shard_units(n_units)

    def unit(u):
        v = torch.full((1, 2 ** n), 1.0 + u, dtype=torch.complex128)
        return plan.hpsi(0.001 * u, v).abs().sum().reshape(1)

    local = parallel.run_units(n_units, unit)
    gathered = parallel.gather_results(local, n_units)
    serial = torch.stack([unit(u) for u in range(n_units)])
    torch.save({"errs": errs, "units": mine, "gather_err": (gathered - serial).abs().max().item()},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_paths_gloo(world, emu_library, tmp_path):
    n = 6
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    seen = []
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert max(res["errs"]) < 1e-13
        assert res["gather_err"] == 0.0
        seen += res["units"]
    assert sorted(seen) == list(range(7))


def _evolve_worker(rank, world, port, n, out_dir, c64=False):
    """configs[4] on the CPU tier: sharded DP5 evolution + discrete adjoint vs the single-process
    engine on the full register (same library, so the comparison isolates the sharding)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    from pulser_diff_b200 import _cabi, ops, parallel
    _cabi.use_library(EMU, EMU_C64)
    dev = torch.device("cpu")
    cd = torch.complex64 if c64 else torch.complex128
    pr = _program(n, T=40)
    gen = torch.Generator().manual_seed(9)
    psi0 = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=gen)
    psi0 = psi0 / psi0.norm()
    tsave = torch.tensor([0.0, 0.011, 0.03, 0.0525], dtype=torch.float64)
    w = torch.rand(len(tsave), 1, 2 ** n, dtype=torch.float64, generator=gen)
    v = torch.randn(len(tsave), 1, 2 ** n, dtype=torch.complex128, generator=gen)

    def loss_of(st):
        return (w * st.abs() ** 2).sum() + (v.conj() * st).real.sum()

    # single-process run of the full register
    leaves = [x.clone().requires_grad_(True) for x in (psi0, pr["det_values"], pr["amp_values"], pr["pair_u"])]
    st_full = ops.evolve(leaves[0], tsave, leaves[1], leaves[2], leaves[3], n_qubits=n, kind=_cabi.PD_KET,
                         dt=pr["dt"], det_masks=pr["det_masks"], amp_masks=pr["amp_masks"])
    g_full = torch.autograd.grad(loss_of(st_full), leaves)
    log = [r for r in ops.last_step_log(st_full) if r["accepted"]]

    sk = parallel.ShardedKet(n, pr["pair_u"], pr["dt"], pr["det_masks"], pr["det_values"],
                             pr["amp_masks"], pr["amp_values"], dev, dtype=cd)
    n_loc = 2 ** sk.nl
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    res = {}
    psi0 = psi0.to(cd)
    # free-running controller: the same first steps (the all-reduced error norm differs from the
    # engine's by rounding, which the controller amplifies until an accept/reject decision
    # flips -- hence the shared-step protocol below for the tight comparison)
    st_free, steps_free = sk.evolve(sk.local_slice(psi0), tsave.tolist())
    res["n_steps"] = (len(steps_free), len(log))
    res["dt_err"] = max(abs(a[1] - b["dt"]) / b["dt"] for a, b in zip(steps_free[:3], log[:3]))
    res["free_err"] = (st_free - st_full.detach()[:, :, sl]).abs().max().item()
    # shared-step protocol for the tight comparison of states and gradients
    replay = [(r["t"], r["dt"], r["interval"], bool(r["clipped"])) for r in log]
    st, steps = sk.evolve(sk.local_slice(psi0), tsave.tolist(), replay=replay)
    res["state_err"] = (st - st_full.detach()[:, :, sl]).abs().max().item()
    res["state_dtype"] = str(st.dtype)
    st_leaf = st.clone().requires_grad_(True)
    # this rank's share of the loss: its slices only
    l_loc = (w[:, :, sl] * st_leaf.abs() ** 2).sum() + (v[:, :, sl].conj() * st_leaf).real.sum()
    (g_st,) = torch.autograd.grad(l_loc, st_leaf)
    out = sk.evolve_backward(st, g_st, steps)
    relerr = lambda a, b: ((a.to(b.dtype) - b).abs().max() / b.abs().max()).item()
    res["g_det"] = relerr(out["det"], g_full[1])
    res["g_amp"] = relerr(out["amp"], g_full[2])
    res["g_pair"] = relerr(out["pair"], g_full[3])
    res["g_psi0"] = relerr(out["state0"], g_full[0][:, sl])
    torch.save(res, os.path.join(out_dir, f"e{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_evolution_and_gradient_gloo(world, emu_library, tmp_path):
    """Sharded register, full pulse sequence + gradient (BASELINE configs[4]): states to 1e-10,
    gradients to 1e-8 relative (north_star's tolerances) against the unsharded engine."""
    n = 6
    mp.spawn(_evolve_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"e{r}.pt"))
        assert abs(res["n_steps"][0] - res["n_steps"][1]) <= 0.1 * res["n_steps"][1], res
        assert res["dt_err"] < 1e-9 and res["free_err"] < 1e-3, res
        assert res["state_err"] < 1e-12, res
        for key in ("g_det", "g_amp", "g_pair", "g_psi0"):
            assert res[key] < 1e-8, (key, res)


def test_sharded_complex64_gloo(emu_library, tmp_path):
    """The sharded register in the complex64 tier (slices, exchange and stage vectors in complex64) against
    the unsharded complex128 engine on its accepted steps: states 1e-5, gradients 2e-5 of their scale."""
    n, world = 6, 2
    mp.spawn(_evolve_worker, args=(world, _free_port(), n, str(tmp_path), True), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"e{r}.pt"))
        assert res["state_dtype"] == "torch.complex64"
        assert res["state_err"] < 1e-5, res
        assert max(res["g_det"], res["g_amp"], res["g_pair"], res["g_psi0"]) < 2e-5, res
