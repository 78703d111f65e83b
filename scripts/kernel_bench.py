"""Kernel-level timing of the ket path: one H.psi and one fused DP5 step per register size.

    python scripts/kernel_bench.py --n 20 22 24 26 --path 0 1 2

Timing is CUDA events inside the C ABI (pd_bench_hpsi / pd_bench_dp5_steps).  Prints one JSON
line per (n, path) with achieved algorithmic GB/s (40 B and 576 B per amplitude).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pulser_diff_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[22])
    ap.add_argument("--path", type=int, nargs="+", default=[0])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    T = 64
    g = torch.Generator().manual_seed(0)
    dv = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4
    av = torch.complex(torch.rand(1, T, dtype=torch.float64, generator=g) * 3,
                       torch.zeros(1, T, dtype=torch.float64))
    for n in args.n:
        u = torch.zeros(n, n, dtype=torch.float64)
        for i in range(n):
            for j in range(i + 1, n):
                u[i, j] = 865723.02 / (7.0 * (j - i)) ** 6
        plan = _cabi.Plan(n, args.batch, _cabi.PD_KET, dev)
        plan.set_interaction(u)
        plan.set_terms(0.02, [(1 << n) - 1], dv, [(1 << n) - 1], av)
        y = torch.zeros(args.batch, 2 ** n, dtype=torch.complex128, device=dev)
        y[:, -1] = 1.0
        psi = torch.randn(args.batch, 2 ** n, dtype=torch.float64, device=dev).to(torch.complex128)
        for path in args.path:
            plan.set_path(path)
            ms_step = plan.bench_dp5_steps(0.3, 1e-3, args.steps, y)
            ms_h = plan.bench_hpsi(0.3, psi, args.steps * 3)
            amps = args.batch * 2 ** n
            print(json.dumps({"n": n, "batch": args.batch, "path": path, "ms_dp5_step": ms_step,
                              "dp5_GBs": 576 * amps / ms_step / 1e6, "dp5_frac": 576 * amps / ms_step / 1e6 / peak,
                              "ms_hpsi": ms_h, "hpsi_GBs": 40 * amps / ms_h / 1e6,
                              "hpsi_frac": 40 * amps / ms_h / 1e6 / peak}), flush=True)
        del plan, y, psi
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
