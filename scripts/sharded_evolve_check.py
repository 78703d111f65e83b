"""Sharded register on real GPUs (torchrun, one rank per GPU): full pulse sequence + gradient.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/sharded_evolve_check.py

1. parity (N = PD_PARITY_QUBITS, default 16): ShardedKet.evolve / evolve_backward over NVLink peer
   memory vs the single-GPU engine on the full register, shared step sequence.
2. timing: DP5 steps of the sharded register at 2^PD_LOCAL_QUBITS amplitudes per GPU.
"""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pulser_diff_b200 import _cabi, ops, parallel
from test_parallel_gloo import _program

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = world.bit_length() - 1
n = int(os.environ.get("PD_PARITY_QUBITS", "16"))
pr = _program(n, T=40)
gen = torch.Generator().manual_seed(9)
psi0 = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=gen)
psi0 = (psi0 / psi0.norm()).to(dev)
tsave = torch.tensor([0.0, 0.004, 0.009], dtype=torch.float64)
w = torch.rand(len(tsave), 1, 2 ** n, dtype=torch.float64, generator=gen).to(dev)
v = torch.randn(len(tsave), 1, 2 ** n, dtype=torch.complex128, generator=gen).to(dev)
leaves = [x.clone().requires_grad_(True) for x in (psi0, pr["det_values"], pr["amp_values"], pr["pair_u"])]
st_full = ops.evolve(leaves[0], tsave, leaves[1], leaves[2], leaves[3], n_qubits=n, kind=_cabi.PD_KET,
                     dt=pr["dt"], det_masks=pr["det_masks"], amp_masks=pr["amp_masks"])
loss = (w * st_full.abs() ** 2).sum() + (v.conj() * st_full).real.sum()
g_full = torch.autograd.grad(loss, leaves)
log = [r for r in ops.last_step_log(st_full) if r["accepted"]]
replay = [(r["t"], r["dt"], r["interval"], bool(r["clipped"])) for r in log]

sk = parallel.ShardedKet(n, pr["pair_u"], pr["dt"], pr["det_masks"], pr["det_values"],
                         pr["amp_masks"], pr["amp_values"], dev, peer_memory=True)
n_loc = 2 ** sk.nl
sl = slice(rank * n_loc, (rank + 1) * n_loc)
st, steps = sk.evolve(sk.local_slice(psi0), tsave.tolist(), replay=replay)
state_err = (st - st_full.detach()[:, :, sl]).abs().max().item()
st_leaf = st.clone().requires_grad_(True)
l_loc = (w[:, :, sl] * st_leaf.abs() ** 2).sum() + (v[:, :, sl].conj() * st_leaf).real.sum()
(g_st,) = torch.autograd.grad(l_loc, st_leaf)
out = sk.evolve_backward(st, g_st, steps)
rel = lambda a, b: ((a.cpu() - b.cpu()).abs().max() / b.abs().max().cpu()).item()
errs = torch.tensor([state_err, rel(out["det"], g_full[1]), rel(out["amp"], g_full[2]),
                     rel(out["pair"], g_full[3]), rel(out["state0"], g_full[0][:, sl])],
                    dtype=torch.float64, device=dev)
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
st_free, steps_free = sk.evolve(sk.local_slice(psi0), tsave.tolist())
del sk, st, st_full, st_leaf, out, w, v, leaves, g_full, st_free
ops.clear_plan_cache(); torch.cuda.empty_cache()

# ---- timing at scale ------------------------------------------------------------------------
nl = int(os.environ.get("PD_LOCAL_QUBITS", "24"))
n2 = nl + g
pr2 = _program(n2)
sk2 = parallel.ShardedKet(n2, pr2["pair_u"], pr2["dt"], pr2["det_masks"], pr2["det_values"],
                          pr2["amp_masks"], pr2["amp_values"], dev, peer_memory=True)
loc = torch.zeros(1, 2 ** nl, dtype=torch.complex128, device=dev)
if rank == world - 1:
    loc[0, -1] = 1.0                                   # all-ground register
n_steps = int(os.environ.get("PD_STEPS", "6"))
fixed = [(i * 1e-4, 1e-4, 1, i == n_steps - 1) for i in range(n_steps)]
sk2.evolve(loc, [0.0, n_steps * 1e-4], replay=fixed[:2] + [(2e-4, 1e-4, 1, True)] if n_steps > 3 else fixed)
dist.barrier(); torch.cuda.synchronize(dev)
t0 = time.perf_counter()
st2, steps2 = sk2.evolve(loc, [0.0, n_steps * 1e-4], replay=fixed)
torch.cuda.synchronize(dev); dist.barrier()
t_fwd = time.perf_counter() - t0
norm = torch.tensor([(st2[-1].abs() ** 2).sum().item()], dtype=torch.float64, device=dev)
dist.all_reduce(norm)
gs = torch.zeros_like(st2); gs[-1] = st2[-1]
torch.cuda.synchronize(dev); dist.barrier()
t0 = time.perf_counter()
out2 = sk2.evolve_backward(st2, gs, steps2)
torch.cuda.synchronize(dev); dist.barrier()
t_bwd = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"world": world, "parity_n": n, "accepted_steps": len(replay), "free_steps": len(steps_free),
                      "state_err": errs[0].item(), "g_det_rel": errs[1].item(), "g_amp_rel": errs[2].item(),
                      "g_pair_rel": errs[3].item(), "g_psi0_rel": errs[4].item(),
                      "timed_n": n2, "local_qubits": nl, "dp5_steps": n_steps,
                      "s_per_step_forward": t_fwd / n_steps, "s_per_step_adjoint": t_bwd / n_steps,
                      "norm_after": norm.item()}))
assert errs[0].item() < 1e-10 and errs[1:].max().item() < 1e-8
dist.destroy_process_group()
