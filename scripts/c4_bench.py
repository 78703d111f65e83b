"""configs[3] leg of bench.py on its own (forward + adjoint seconds of the 4^12 Lindblad run)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
dev = torch.device("cuda", 0)
peak, _ = B.measured_peak()
out = B.c4_lindblad(dev, peak)
print(json.dumps({k: out[k] for k in ("fwd_s", "adjoint_s", "adjoint_over_forward", "steps_per_s_forward", "loss")}),
      {k: v for k, v in os.environ.items() if k.startswith("PD_")})
