import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pulser_diff_b200 import _cabi, ops
from test_gpu_scale import _program, _vary_drive
dev = torch.device("cuda", 0)
n, nb, drive = 22, 2, sys.argv[1] if len(sys.argv) > 1 else "real"
pr = _vary_drive(_program(n, T=16), n, drive)
psi0 = torch.randn(nb, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(5)).to(dev)
psi0 /= psi0.norm(dim=1, keepdim=True)
tsave = torch.tensor([0.0, 0.004, 0.008], dtype=torch.float64)
ref = None
for rep in range(int(os.environ.get("REPS", "6"))):
    path = 1 if rep == 0 else 4
    av = pr["amp_values"].clone().requires_grad_(True)
    dv = pr["det_values"].clone().requires_grad_(True)
    pu = pr["pair_u"].clone().requires_grad_(True)
    st = ops.evolve(psi0, tsave, dv, av, pu, n_qubits=n, kind=_cabi.PD_KET, dt=pr["dt"], det_masks=pr["det_masks"],
                    amp_masks=pr["amp_masks"], options=_cabi.Options(path=path))
    w = torch.arange(2 ** n, device=dev).remainder(5).to(torch.float64)
    val = (w * st[-1].abs() ** 2).sum() + (w * st[1].abs() ** 2).sum()
    gr = torch.autograd.grad(val, [av, dv, pu])
    plan = ops.get_plan(n, nb, _cabi.PD_KET, dev)
    plan.set_path(path)
    hp = plan.hpsi(0.0051, psi0).clone()
    nsteps = getattr(plan, "last_steps", None)
    plan.set_path(0)
    cur = (st.detach().clone(), hp, [g.clone() for g in gr])
    if ref is None:
        ref = cur
        continue
    d = (cur[0] - ref[0]).abs()
    print(json.dumps({"rep": rep, "states": [d[k].max().item() for k in range(3)], "hpsi": (cur[1] - ref[1]).abs().max().item(),
                      "grads": [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(cur[2], ref[2])],
                      "n_bad1": int((d[1] > 1e-10).sum()), "n_bad2": int((d[2] > 1e-10).sum())}), flush=True)
