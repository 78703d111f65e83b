import torch, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_diff_b200 import _cabi, ops
dev = torch.device("cuda", 0)
n, T = 12, 16
g = torch.Generator().manual_seed(0)
full = (1 << n) - 1
dv = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4
av = torch.complex(torch.rand(1, T, dtype=torch.float64, generator=g) * 3, torch.zeros(1, T, dtype=torch.float64))
u = torch.zeros(n, n, dtype=torch.float64)
for i in range(n):
    for j in range(i + 1, n):
        u[i, j] = 865723.02 / (7.0 * (j - i)) ** 6
col = torch.tensor([[[0.5, 0], [0, -0.5]], [[0, 0], [0.3162, 0]]], dtype=torch.complex128)
plan = _cabi.Plan(n, 1, _cabi.PD_DENSITY, dev)
ops.configure(plan, ops.make_program(n, _cabi.PD_DENSITY, 0.02, [full], dv, [full], av, u, col))
y = torch.zeros(1, 4 ** n, dtype=torch.complex128, device=dev); y[0, -1] = 1.0
for path in (1, 0):
    plan.set_path(path)
    ms = plan.bench_dp5_steps(0.3, 1e-3, 3, y.clone())
    print(json.dumps({"n": n, "path": path, "ms_dp5_me_step": ms, "alg_GBs": 528.0 * 4 ** n / ms / 1e6, "frac": 528.0 * 4 ** n / ms / 1e6 / 6548.2}))
