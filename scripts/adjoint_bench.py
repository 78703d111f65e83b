"""Forward + adjoint per DP5 step at N qubits (bench.fwd_grad_large), for A/B runs of the PD_CORR_* switches.
Usage: python scripts/adjoint_bench.py [N] [c64]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
dev = torch.device("cuda", 0)
peak, _ = B.measured_peak()
os.environ.setdefault("PD_TIMING", "1")
out = B.fwd_grad_large(dev, n, 4, peak, torch.complex64 if "c64" in sys.argv else torch.complex128)
print(json.dumps({k: out[k] for k in ("fwd_ms_per_step", "adjoint_ms_per_step", "adjoint_over_forward")}),
      {k: os.environ[k] for k in os.environ if k.startswith("PD_CORR")})
