import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pulser_diff_b200 import _cabi, ops
if os.environ.get("PD_LIB"): _cabi.use_library(os.path.join(ROOT, os.environ["PD_LIB"]))
from test_gpu_scale import _program, _vary_drive
dev = torch.device("cuda", 0)
for n, nb, drive in [(22, 1, "real"), (22, 2, "real"), (22, 2, "complex"), (22, 1, "complex"), (24, 1, "real"), (24, 1, "complex")]:
    pr = _vary_drive(_program(n, T=16), n, drive)
    psi0 = torch.randn(nb, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(5)).to(dev)
    psi0 /= psi0.norm(dim=1, keepdim=True)
    tsave = torch.tensor([0.0, 0.004, 0.008], dtype=torch.float64)
    outs = []
    for path in (1, 4):
        st = ops.evolve(psi0, tsave, pr["det_values"], pr["amp_values"], pr["pair_u"], n_qubits=n, kind=_cabi.PD_KET,
                        dt=pr["dt"], det_masks=pr["det_masks"], amp_masks=pr["amp_masks"], options=_cabi.Options(path=path))
        outs.append(st.detach().clone())
    d = (outs[0] - outs[1]).abs()
    info = {"n": n, "batch": nb, "drive": drive, "max_per_save": [d[k].max().item() for k in range(d.shape[0])]}
    bad = (d[1] > 1e-10).nonzero()
    info["n_bad"] = int(bad.shape[0])
    if bad.shape[0]:
        idx = bad[:, -1]
        info["bad_cols"] = sorted(set(bad[:, 0].tolist()))
        info["tiles"] = sorted(set((idx >> 12).tolist()))[:16]
        info["n_tiles_bad"] = len(set((idx >> 12).tolist()))
        info["low12_sample"] = sorted(set((idx & 4095).tolist()))[:16]
    print(json.dumps(info), flush=True)
    ops._PLAN_CACHE.clear() if hasattr(ops, "_PLAN_CACHE") else None
    torch.cuda.empty_cache()
