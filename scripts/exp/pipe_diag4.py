import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pulser_diff_b200 import _cabi, ops
from test_gpu_scale import _program, _vary_drive
dev = torch.device("cuda", 0)
n, nb = 22, 2
pr = _vary_drive(_program(n, T=16), n, "real")
plan = _cabi.Plan(n, nb, _cabi.PD_KET, dev)
plan.set_interaction(pr["pair_u"])
plan.set_terms(pr["dt"], pr["det_masks"], pr["det_values"], pr["amp_masks"], pr["amp_values"])
psi = torch.randn(nb, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(5)).to(dev)
plan.set_path(1)
ref = plan.hpsi(0.0051, psi).clone()
plan.set_path(4)
nfail = 0
for rep in range(int(os.environ.get("REPS", "60"))):
    out = plan.hpsi(0.0051, psi)
    d = (out - ref).abs()
    bad = (d > 1e-9).nonzero()
    if bad.shape[0] == 0:
        continue
    nfail += 1
    if nfail > 4:
        continue
    col, idx = bad[:, 0], bad[:, 1]
    info = {"rep": rep, "n_bad": int(bad.shape[0]), "cols": sorted(set(col.tolist())),
            "A_tiles(idx>>12)": sorted(set((idx >> 12).tolist()))[:40], "n_A_tiles": len(set((idx >> 12).tolist())),
            "uh(idx>>17)": sorted(set((idx >> 17).tolist())), "ul((idx>>7)&31)": sorted(set(((idx >> 7) & 31).tolist())),
            "rows((idx>>12)&31)": sorted(set(((idx >> 12) & 31).tolist())),
            "low7": [int(idx.bitwise_and(127).min()), int(idx.bitwise_and(127).max())]}
    # compare with the A-only result: is the wrong value = missing group contribution, or stale?
    print(json.dumps(info), flush=True)
print("failures", nfail)
