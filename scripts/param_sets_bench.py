"""configs[2]: gate-pulse optimisation over a batch of pulse-parameter sets (docs/gate_optimization.ipynb
scale-up): N=2 atoms 6.5 um apart (C6 of Rydberg level 60), 8 constant pulses x 131 ns with parameters
(amp, det, phase) x 8 per set, psi0 = eye(4), rate 0.05, loss = Hadamard^(x2) gate infidelity, gradient
w.r.t. all 24 parameters of every set.  One rank evolves its share of the sets in ONE call
(ops.evolve_units: one CTA per set); under torchrun the sets are dealt to the ranks with no data-path
collective (parallel.shard_units) and only the losses are gathered.

    python scripts/param_sets_bench.py [n_sets=4096] [check=8]
"""
import json, math, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulser_diff_b200 import _cabi, ops, parallel

n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_check = int(sys.argv[2]) if len(sys.argv) > 2 else 8
world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)

N, PULSES, DUR, RATE, C6, SPACING = 2, 8, 131, 0.05, 865723.02, 6.5
D = PULSES * DUR + 1                       # one extra sample (reference backend.py:114-115)
n_keep = int(RATE * D)
idx = torch.linspace(0, D - 1, n_keep, dtype=torch.int).long()   # hamiltonian.py:83-91
dt = 0.001 / RATE
tsave = torch.cat([(torch.arange(D, dtype=torch.float64) / 1000)[idx],
                   torch.tensor([0.0, (D - 1) / 1000], dtype=torch.float64)]).unique()
pair_u = torch.zeros(N, N, dtype=torch.float64); pair_u[0, 1] = C6 / SPACING ** 6
full = (1 << N) - 1
had = torch.tensor([[1, 1], [1, -1]], dtype=torch.complex128) / math.sqrt(2)
target = torch.kron(had, had).to(dev)
mine = parallel.shard_units(n_sets, rank, world) if world > 1 else list(range(n_sets))

def tables(params):
    """params (U, 3, PULSES) -> reference coefficient arrays: 0.5*amp*exp(-i phase), -0.5*det, sub-sampled."""
    U = params.shape[0]
    z = torch.zeros(U, 1, dtype=torch.float64)
    full_s = [torch.cat([params[:, k].repeat_interleave(DUR, dim=1), z], dim=1) for k in range(3)]
    amp, det, ph = (f[:, idx] for f in full_s)
    av = 0.5 * amp * torch.exp(-1j * ph)
    dv = -0.5 * det
    return dv[:, None, :], av[:, None, :]

g = torch.Generator().manual_seed(0)
params_all = torch.rand(n_sets, 3, PULSES, dtype=torch.float64, generator=g) * 4 * math.pi
params = params_all[mine].clone().requires_grad_(True)
U = len(mine)
psi0 = torch.eye(4, dtype=torch.complex128, device=dev).repeat(U, 1, 1)     # rows = initial states

phase_ms = {}


def run():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dv, av = tables(params)
    t1 = time.perf_counter()
    st = ops.evolve_units(psi0, tsave, dv, av, pair_u, n_qubits=N, dt=dt, det_masks=[full], amp_masks=[full])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    Uf = st[:, -1].transpose(1, 2)                       # column b = evolved basis state b
    fid = (target.conj().T @ Uf).diagonal(dim1=1, dim2=2).sum(-1).abs() / 4
    loss = 1 - fid
    (gp,) = torch.autograd.grad(loss.sum(), [params])
    torch.cuda.synchronize(); t3 = time.perf_counter()
    phase_ms.update(tables=(t1 - t0) * 1e3, forward=(t2 - t1) * 1e3, loss_and_backward=(t3 - t2) * 1e3)
    return loss.detach(), gp

for _ in range(2):
    loss, gp = run()
torch.cuda.synchronize(); t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    loss, gp = run()
torch.cuda.synchronize(); dt_s = (time.perf_counter() - t0) / reps
plan = ops.get_plan(N, 4, _cabi.PD_KET, dev)
# parity on a few sets: the same set on its own through ops.evolve
errs = []
for u in range(min(n_check, U)):
    p1 = params.detach()[u:u + 1].clone().requires_grad_(True)
    dv, av = tables(p1)
    st = ops.evolve(psi0[0], tsave, dv[0], av[0], pair_u, n_qubits=N, kind=_cabi.PD_KET, dt=dt,
                    det_masks=[full], amp_masks=[full])
    Uf = st[-1].transpose(0, 1)
    l1 = 1 - (target.conj().T @ Uf).diagonal().sum().abs() / 4
    (g1,) = torch.autograd.grad(l1, [p1])
    errs.append(max(abs(l1.item() - loss[u].item()), (g1[0] - gp[u]).abs().max().item()))
if world > 1:
    tt = torch.tensor([dt_s], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_s = tt.item()
    losses = parallel.gather_results({u: loss[i].reshape(1) for i, u in enumerate(mine)}, n_sets, dev)
else:
    losses = loss
if rank == 0:
    print(json.dumps({"config": "C3 parameter-set batch", "n_sets": n_sets, "n_gpus": world, "n_qubits": N,
                      "columns": 4, "n_t": int(tsave.numel()), "s_per_sweep": dt_s,
                      "sets_per_s": n_sets / dt_s, "mean_loss": losses.mean().item(),
                      "first_loss": losses[0].item(), "max_abs_diff_vs_single": max(errs) if errs else None,
                      "phase_ms_rank0": {k: round(v, 2) for k, v in phase_ms.items()}}))
if world > 1:
    dist.destroy_process_group()
