"""Column-tile density kernels A/B (PD_DENSITY_CT=0/1): one apply, one DP5_ME step, complex128 and complex64."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pulser_diff_b200 import _cabi, ops
dev = torch.device("cuda", 0)
n, T = int(sys.argv[1]) if len(sys.argv) > 1 else 12, 16
g = torch.Generator().manual_seed(0)
full = (1 << n) - 1
dv = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4
av = torch.complex(torch.rand(1, T, dtype=torch.float64, generator=g) * 3, torch.zeros(1, T, dtype=torch.float64))
u = torch.zeros(n, n, dtype=torch.float64)
for i in range(n):
    for j in range(i + 1, n):
        u[i, j] = 865723.02 / (7.0 * (j - i)) ** 6
col = torch.tensor([[[0.5, 0], [0, -0.5]], [[0, 0], [0.3162, 0]]], dtype=torch.complex128)
for cd in (torch.complex128, torch.complex64):
    plan = _cabi.Plan(n, 1, _cabi.PD_DENSITY, dev, cd)
    ops.configure(plan, ops.make_program(n, _cabi.PD_DENSITY, 0.02, [full], dv, [full], av, u, col))
    y = torch.zeros(1, 4 ** n, dtype=cd, device=dev); y[0, -1] = 1.0
    psi = torch.randn(1, 4 ** n, dtype=torch.float64, device=dev).to(cd)
    out = plan.hpsi(0.3, psi, rhs=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.hpsi(0.3, psi, rhs=True, out=out)
    e1.record(); torch.cuda.synchronize()
    ms_apply = e0.elapsed_time(e1) / 10
    ms = plan.bench_dp5_steps(0.3, 1e-3, 3, y.clone())
    ab = 16 if cd == torch.complex128 else 8
    alg = (33 * ab) * 4 ** n
    print(json.dumps({"n": n, "dtype": str(cd), "ct": os.environ.get("PD_DENSITY_CT", "0"), "form": os.environ.get("PD_LINDBLAD_FORM", "1"), "ms_apply": ms_apply,
                      "ms_dp5_me_step": ms, "frac": alg / ms / 1e6 / 6548.2, "checksum": float(out.abs().sum())}))
    del plan
