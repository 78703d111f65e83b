"""Forward + adjoint-gradient wall time of a short DP5_SE evolution at N qubits (BASELINE metric
"fwd+grad wall time at N qubits"): fixed step sequence (replay) so that the work is n_steps steps.

    python scripts/grad_bench.py 20 24 26
"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulser_diff_b200 import _cabi, ops

dev = torch.device("cuda", 0)
T = 64
g = torch.Generator().manual_seed(0)
dv0 = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4
av0 = torch.complex(torch.rand(1, T, dtype=torch.float64, generator=g) * 3, torch.zeros(1, T, dtype=torch.float64))
n_steps = 4
for n in [int(a) for a in (sys.argv[1:] or ["20"])]:
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = 865723.02 / (7.0 * (j - i)) ** 6
    full = (1 << n) - 1
    psi0 = torch.zeros(1, 2 ** n, dtype=torch.complex128, device=dev); psi0[0, -1] = 1.0
    tsave = torch.tensor([0.0, 0.004], dtype=torch.float64)
    replay = [(0.001, False)] * (n_steps - 1) + [(0.001, True)]
    w = torch.arange(2 ** n, device=dev).remainder(7).to(torch.float64)
    res = {}
    for rep in range(3):      # the third sweep runs with the slope cache set up by the second
        dv = dv0.clone().requires_grad_(True); av = av0.clone().requires_grad_(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = ops.evolve(psi0, tsave, dv, av, u, n_qubits=n, kind=_cabi.PD_KET, dt=0.02, det_masks=[full],
                        amp_masks=[full], options=_cabi.Options(replay=replay))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        val = (w * st[-1, 0].abs() ** 2).sum()
        torch.autograd.grad(val, [dv, av])
        torch.cuda.synchronize(); t2 = time.perf_counter()
        res = {"n": n, "dp5_steps": n_steps, "fwd_ms_per_step": (t1 - t0) * 1e3 / n_steps,
               "bwd_ms_per_step": (t2 - t1) * 1e3 / n_steps, "norm": float(st[-1].norm())}
        del st, val
    print(json.dumps(res), flush=True)
    ops.clear_plan_cache(); torch.cuda.empty_cache()
