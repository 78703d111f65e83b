"""Golden fixture for the FULL C2 workload (12-atom chain, 1100 ns, rate 0.05, all 55 tsave intervals):
the oracle's DP5_SE forward run with pyqtorch's default controller (atol 1e-8, rtol 1e-6) and the tape
gradient of the bench loss w.r.t. the 60 pulse parameters.  The oracle needs ~10 minutes for this on 8 cores,
too long for a test, so its outputs are stored here: the accepted step sequence (replayed by the CUDA path
for round-off parity), the attempted-step log, the loss at every evaluation time, the final
state and the gradient.  Run:  python tests/golden/make_c2_full.py   (writes tests/golden/c2_full.json)
"""
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B  # noqa: E402
from helpers import Channel, Problem  # noqa: E402
from oracle.ref_solvers import SolverType as RefSolver, sesolve as ref_sesolve  # noqa: E402
from pulser_diff_b200.utils import interpolate_sine  # noqa: E402

torch.set_num_threads(os.cpu_count())
interp = interpolate_sine(B.N_PARAM, B.DURATION).to(torch.float64)
ta, td = B.workload_params(0)
amp, det, ph = B.pulse_samples(ta, td, interp)
p = Problem(B.chain_coords(B.N_QUBITS), B.C6, [Channel(amp, det, ph)], rate=B.RATE)
ref = p.ref()
t0 = time.time()
r = ref_sesolve(ref.ham.H, ref.initial_state, ref.evaluation_times, RefSolver.DP5_SE, {})
t1 = time.time()
d = B.loss_diag(B.N_QUBITS, "cpu")
series = (d[None, :, None] * r.states.abs() ** 2).sum(dim=(1, 2))
loss = series[-1]
g_ta, g_td = torch.autograd.grad(loss, [ta, td])
t2 = time.time()
final = r.states[-1].detach().reshape(-1)
out = {"_doc": "tests/golden/make_c2_full.py", "oracle_forward_s": t1 - t0, "oracle_backward_s": t2 - t1,
       "threads": torch.get_num_threads(), "tsave": ref.evaluation_times.tolist(),
       "steplog": [[float(a), float(b), bool(c), float(e), bool(f)] for a, b, c, e, f in r.steplog],
       "loss_series": series.detach().tolist(), "grad_ta": g_ta.tolist(), "grad_td": g_td.tolist(),
       "final_re": final.real.tolist(), "final_im": final.imag.tolist()}
json.dump(out, open(os.path.join(HERE, "c2_full.json"), "w"))
print("forward", t1 - t0, "s, backward", t2 - t1, "s, steps", len(r.steplog))
