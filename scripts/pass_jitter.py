"""Per-pass wall times of the C2 resident pass (jitter hunting)."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
import pulser_diff_b200 as pdb
from pulser_diff_b200 import _cabi, ops
from pulser_diff_b200.samples import ChannelSamples, SequenceSamples
from pulser_diff_b200.utils import interpolate_sine, expect_diag
dev = torch.device("cuda", 0)
n = B.N_QUBITS
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
ts = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 80):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = ops.evolve(psi0, tsave, dv_d, av_d, H.pair_u.detach(), n_qubits=n, kind=_cabi.PD_KET, dt=H.dt,
                    det_masks=dm, amp_masks=am, solver=_cabi.SOLVER_DP5_SE)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    val = expect_diag(diag, st.permute(0, 2, 1)).real[-1]
    torch.autograd.grad(val, [dv_d, av_d])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((round((t1 - t0) * 1e3, 1), round((t2 - t1) * 1e3, 1)))
print(ts)
