"""NCCL check of the top-qubit sharded H.psi on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/sharded_check.py

1. parity: ShardedKet.hpsi on world ranks == the single-GPU H.psi of the full register (N = 20),
2. timing: sharded H.psi at 2^26 amplitudes per GPU (CUDA events, max over ranks).
"""
import json, os, sys
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
n = 20
pr = _program(n)
psi = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(3)).to(dev)
plan = _cabi.Plan(n, 1, _cabi.PD_KET, dev)
ops.configure(plan, ops.make_program(n, _cabi.PD_KET, pr["dt"], pr["det_masks"], pr["det_values"],
                                     pr["amp_masks"], pr["amp_values"], pr["pair_u"], None))
sk = parallel.ShardedKet(n, pr["pair_u"], pr["dt"], pr["det_masks"], pr["det_values"],
                         pr["amp_masks"], pr["amp_values"], dev)
skp = parallel.ShardedKet(n, pr["pair_u"], pr["dt"], pr["det_masks"], pr["det_values"],
                          pr["amp_masks"], pr["amp_values"], dev, peer_memory=True)
errs, errs_p = [], []
for t in (0.0, 0.0131, 0.0377):
    full = plan.hpsi(t, psi)
    mine = sk.hpsi(t, sk.local_slice(psi))
    errs.append((mine - sk.local_slice(full)).abs().max().item() / full.abs().max().item())
    mine = skp.hpsi(t, skp.local_slice(psi))
    errs_p.append((mine - skp.local_slice(full)).abs().max().item() / full.abs().max().item())
err = torch.tensor([max(errs), max(errs_p)], dtype=torch.float64, device=dev)
dist.all_reduce(err, op=dist.ReduceOp.MAX)
del plan, sk, skp, psi, full, mine
ops.clear_plan_cache(); torch.cuda.empty_cache()

nl = int(os.environ.get("PD_LOCAL_QUBITS", "26"))
n2 = nl + g
pr2 = _program(n2)
sk2 = parallel.ShardedKet(n2, pr2["pair_u"], pr2["dt"], pr2["det_masks"], pr2["det_values"],
                          pr2["amp_masks"], pr2["amp_values"], dev)
loc = torch.randn(1, 2 ** nl, dtype=torch.float64, device=dev).to(torch.complex128)
for _ in range(3):
    sk2.hpsi(0.01, loc)
dist.barrier(); torch.cuda.synchronize(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    out = sk2.hpsi(0.01, loc)
e1.record()
dist.barrier(); torch.cuda.synchronize(dev)
ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# breakdown: local kernels alone, exchange alone
def timed(fn, reps=5):
    fn(); dist.barrier(); torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize(dev)
    v = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return v.item()
ms_local = timed(lambda: sk2.plan.hpsi(0.01, loc))
# peer-memory variant: state kept in the peer-visible buffer, partner slices read in place
sk3 = parallel.ShardedKet(n2, pr2["pair_u"], pr2["dt"], pr2["det_masks"], pr2["det_values"],
                          pr2["amp_masks"], pr2["amp_values"], dev, peer_memory=os.environ.get("PD_PEER_MODE", "read"))
sbuf = sk3.state_buffer(); sbuf.copy_(loc)
ms_peer = timed(lambda: sk3.hpsi(0.01, sbuf))
ms_peer_staged = timed(lambda: sk3.hpsi(0.01, loc))
chk = (sk3.hpsi(0.01, sbuf) - sk2.hpsi(0.01, loc)).abs().max().item()
ptrs = [sk3._peer_ptrs[rank ^ (1 << k)] for k in range(g)]
ms_acc = timed(lambda: sk3.plan.sharded_accumulate(out, sbuf, 0.1, ptrs, [0.3 + 0.1j] * g))
ms_acc1 = timed(lambda: sk3.plan.sharded_accumulate(out, sbuf, 0.1, ptrs[:1], [0.3 + 0.1j]))
recv = torch.empty_like(loc)
def xchg():
    peer = rank ^ 1
    for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, loc, peer), dist.P2POp(dist.irecv, recv, peer)]):
        r.wait()
ms_x = timed(xchg)
ms_axpy = timed(lambda: out.add_(recv, alpha=0.3 + 0.1j))
if rank == 0:
    print(json.dumps({"ms_local_hpsi": ms_local, "ms_exchange_1GiB": ms_x, "exchange_GBs": 2 ** nl * 16 / ms_x / 1e6,
                      "ms_axpy": ms_axpy, "ms_peer_hpsi": ms_peer, "ms_peer_hpsi_staged": ms_peer_staged,
                      "peer_vs_exchange_maxabs": chk, "ms_peer_accumulate_all": ms_acc,
                      "ms_peer_accumulate_1": ms_acc1,
                      "peer_read_GBs_1": 2 ** nl * 16 / ms_acc1 / 1e6,
                      "peer_read_GBs_all": g * 2 ** nl * 16 / ms_acc / 1e6}))
    amps = 2 ** nl
    t = ms.item() * 1e-3
    print(json.dumps({"world": world, "parity_n": n, "max_rel_err": err[0].item(), "max_rel_err_peer": err[1].item(), "sharded_n": n2,
                      "local_qubits": nl, "ms_per_hpsi": ms.item(),
                      "hbm_alg_GBs_per_gpu": 40.0 * amps / t / 1e9,
                      "nvlink_GBs_per_dir_per_gpu": g * 16.0 * amps / t / 1e9}))
assert err.max().item() < 1e-12
dist.destroy_process_group()
