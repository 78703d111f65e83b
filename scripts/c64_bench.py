"""complex64 vs complex128: fused DP5 step and H.psi at N qubits (device time from the C ABI's CUDA events),
plus the adjoint per step.  Usage: python scripts/c64_bench.py [N ...] [c64only] [--lib=PATH_TO_A_C64_BUILD]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulser_diff_b200 import _cabi

dev = torch.device("cuda", 0)
libs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--lib=")]
if libs:
    _cabi.use_library(None, os.path.abspath(libs[0]))      # A/B builds of the complex64 library
only64 = "c64only" in sys.argv


def plan_for(n, cd):
    plan = _cabi.Plan(n, 1, _cabi.PD_KET, dev, cd)
    x = torch.arange(n, dtype=torch.float64) * 7.0
    r = (x[:, None] - x[None, :]).abs() + torch.eye(n, dtype=torch.float64)
    plan.set_interaction(torch.triu(865723.02 / r ** 6, diagonal=1))
    k = torch.arange(8, dtype=torch.float64)
    full = (1 << n) - 1
    plan.set_terms(0.05, [full], (-0.5 * (1.0 + 0.3 * k))[None], [full], (0.5 * (2.0 + 0.2 * k) + 0j)[None].to(torch.complex128))
    return plan


for n in [int(a) for a in sys.argv[1:] if a.isdigit()] or [26]:
    for cd in ((torch.complex64,) if only64 else (torch.complex128, torch.complex64)):
        plan = plan_for(n, cd)
        y = torch.zeros(1, 2 ** n, dtype=cd, device=dev)
        y[0, -1] = 1.0
        plan.bench_dp5_steps(0.3, 1e-3, 2, y)
        ms = plan.bench_dp5_steps(0.3, 1e-3, 5, y)
        psi = torch.randn(1, 2 ** n, dtype=torch.float64, device=dev).to(cd)
        plan.bench_hpsi(0.3, psi, 3)
        mh = plan.bench_hpsi(0.3, psi, 12)
        ab = 16.0 if cd == torch.complex128 else 8.0
        amps = 2.0 ** n
        print(f"N={n} {str(cd):18s} dp5 step {ms:8.3f} ms ({36 * ab * amps / ms / 1e6:7.0f} GB/s algorithmic)   "
              f"hpsi {mh:7.3f} ms ({(2 * ab + 8) * amps / mh / 1e6:7.0f} GB/s)  |y|={float(y.norm()):.6f}", flush=True)
        del plan, y, psi
        torch.cuda.empty_cache()
