"""Timing of the small-register path on the C2-like workload: forward and backward separately."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
import pulser_diff_b200 as pdb
from pulser_diff_b200 import _cabi, ops
from pulser_diff_b200.samples import ChannelSamples, SequenceSamples
from pulser_diff_b200.utils import interpolate_sine, expect_diag

if os.environ.get("PD_LIB"):
    _cabi.use_library(os.environ["PD_LIB"])
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B.N_QUBITS = n
interp = interpolate_sine(B.N_PARAM, B.DURATION).to(torch.float64)
coords = B.chain_coords(n)
register = {f"q{i}": coords[i] for i in range(n)}
ta, td = B.workload_params(0)
amp, det, ph = B.pulse_samples(ta, td, interp)
em = pdb.TorchEmulator(SequenceSamples([ChannelSamples(amp, det, ph)]), register, pdb.DeviceSpec(B.C6),
                       sampling_rate=B.RATE, torch_device=dev)
H = em._hamiltonian._hamiltonian
dm, dv, am, av = H.masks_and_values()
dv_d = dv.detach().clone().requires_grad_(True); av_d = av.detach().clone().requires_grad_(True)
psi0 = em.initial_state.to(dev).transpose(0, 1).contiguous()
tsave = em.evaluation_times.detach()
diag = B.loss_diag(n, dev)
solver = _cabi.SOLVER_KRYLOV_SE if os.environ.get("PD_SOLVER", "dp5") == "krylov" else _cabi.SOLVER_DP5_SE
for path in [int(p) for p in (sys.argv[2:] or ["0", "1"])]:
    for rep in range(int(os.environ.get("PD_REPS", "3"))):
        print("rep", rep, file=sys.stderr)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = ops.evolve(psi0, tsave, dv_d, av_d, H.pair_u.detach(), n_qubits=n, kind=_cabi.PD_KET, dt=H.dt,
                        det_masks=dm, amp_masks=am, solver=solver, options=_cabi.Options(path=path))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        val = expect_diag(diag, st.permute(0, 2, 1)).real[-1]
        torch.autograd.grad(val, [dv_d, av_d])
        torch.cuda.synchronize(); t2 = time.perf_counter()
    log = ops.last_step_log(st) if solver == _cabi.SOLVER_DP5_SE else []
    print(json.dumps({"n": n, "path": path, "solver": "krylov_se" if solver == _cabi.SOLVER_KRYLOV_SE else "dp5_se", "fwd_ms": (t1 - t0) * 1e3, "bwd_ms": (t2 - t1) * 1e3,
                      "attempts": len(log), "accepted": sum(1 for r in log if r["accepted"]),
                      "E": os.environ.get("PD_SMALL_E", "auto")}), flush=True)
