"""cProfile of the end-to-end C2 pass (host-side overhead hunting)."""
import cProfile, pstats, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
import pulser_diff_b200 as pdb
from pulser_diff_b200 import _cabi, ops
from pulser_diff_b200.samples import ChannelSamples, SequenceSamples
from pulser_diff_b200.utils import interpolate_sine

dev = torch.device("cuda", 0)
n = B.N_QUBITS
interp = interpolate_sine(B.N_PARAM, B.DURATION).to(torch.float64)
coords = B.chain_coords(n)
register = {f"q{i}": coords[i] for i in range(n)}
ta, td = B.workload_params(0)
diag = B.loss_diag(n, dev)
spec = pdb.DeviceSpec(B.C6)

def e2e():
    amp, det, ph = B.pulse_samples(ta, td, interp)
    em = pdb.TorchEmulator(SequenceSamples([ChannelSamples(amp, det, ph)]), register, spec,
                           sampling_rate=B.RATE, torch_device=dev)
    res = em.run(solver=pdb.SolverType.DP5_SE)
    loss = res.expect([diag])[0].real[-1]
    return torch.autograd.grad(loss, [ta, td])

for _ in range(3): e2e()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): e2e()
torch.cuda.synchronize(); print("ms per pass", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): e2e()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
