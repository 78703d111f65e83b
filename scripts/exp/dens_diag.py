import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pulser_diff_b200 import _cabi
dev = torch.device("cuda", 0)
T = 16
def run(n, amp, det, inter, col, phase=0.3):
    g = torch.Generator().manual_seed(0)
    dv = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4 * det
    av = torch.polar(torch.rand(1, T, dtype=torch.float64, generator=g) * 3 * amp, torch.full((1, T), phase, dtype=torch.float64))
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = inter * 865723.02 / (7.0 * (j - i)) ** 6
    plan = _cabi.Plan(n, 1, _cabi.PD_DENSITY, dev)
    plan.set_interaction(u)
    plan.set_terms(0.02, [(1 << n) - 1], dv, [(1 << n) - 1], av)
    plan.set_collapse(col)
    rho = torch.randn(1, 4 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(1)).to(dev)
    res = {}
    for path in (1, 0):
        plan.set_path(path)
        res[path] = plan.hpsi(0.0051, rho, rhs=True).clone()
    d = (res[1] - res[0]).abs()[0]
    bad = (d > 1e-9 * res[1].abs().max()).nonzero().flatten()
    info = {"n": n, "amp": amp, "det": det, "inter": inter, "ncol": 0 if col is None else int(col.shape[0]), "max": d.max().item(), "ref_max": res[1].abs().max().item(), "n_bad": int(bad.numel())}
    if bad.numel():
        orb = 0; andb = (1 << (2 * n)) - 1
        bl = bad.tolist()
        for b in bl[:100000]:
            orb |= b; andb &= b
        info["or_bits"] = bin(orb); info["and_bits"] = bin(andb); info["first"] = bl[:6]
    print(json.dumps(info), flush=True)
    del plan
Z = torch.tensor([[[0.5, 0], [0, -0.5]]], dtype=torch.complex128)
R = torch.tensor([[[0, 0], [0.3, 0]]], dtype=torch.complex128)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
run(n, 0, 1, 0, None)
run(n, 0, 0, 1, None)
run(n, 1, 0, 0, None, phase=0.0)
run(n, 1, 0, 0, None)
run(n, 0, 0, 0, Z)
run(n, 0, 0, 0, R)
run(n, 1, 1, 1, torch.cat([Z, R]))
