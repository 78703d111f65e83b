"""hpsi of the stream family (pipelined dataflow launch) against the gather kernels: where do they differ?"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pulser_diff_b200 import _cabi
dev = torch.device("cuda", 0)
T = 16
for n, batch, phase in [(22, 1, 0.0), (22, 2, 0.0), (22, 1, 0.4), (22, 2, 0.4), (23, 1, 0.0), (23, 2, 0.0), (24, 1, 0.0), (24, 1, 0.3)]:
    g = torch.Generator().manual_seed(0)
    dv = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4
    av = torch.polar(torch.rand(1, T, dtype=torch.float64, generator=g) * 3, torch.full((1, T), phase, dtype=torch.float64))
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = 865723.02 / (7.0 * (j - i)) ** 6
    plan = _cabi.Plan(n, batch, _cabi.PD_KET, dev)
    plan.set_interaction(u)
    plan.set_terms(0.02, [(1 << n) - 1], dv, [(1 << n) - 1], av)
    psi = torch.randn(batch, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(1)).to(dev)
    res = {}
    for path in (1, 4):
        plan.set_path(path)
        res[path] = plan.hpsi(0.0051, psi).clone()
    d = (res[1] - res[4]).abs()
    bad = (d > 1e-9).nonzero()
    info = {"n": n, "batch": batch, "phase": phase, "max": d.max().item(), "n_bad": int(bad.shape[0])}
    if bad.shape[0]:
        idx = bad[:, 1]
        info["first_bad"] = [int(bad[0, 0]), int(bad[0, 1])]
        info["last_bad"] = [int(bad[-1, 0]), int(bad[-1, 1])]
        # which bits vary among the bad indices
        info["or_bits"] = bin(int(torch.bitwise_or(idx.max(), idx.min()))) if False else None
        info["bad_per_col"] = [int((bad[:, 0] == c).sum()) for c in range(batch)]
        info["tile_ids_sample"] = sorted(set((idx[:2000] >> 12).tolist()))[:12]
    print(json.dumps(info), flush=True)
    del plan
    torch.cuda.empty_cache()
