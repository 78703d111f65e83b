"""Where does the host time of one end-to-end C2 pass go?  cProfile over 20 passes of bench.e2e_pass's body."""
import cProfile, pstats, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
import pulser_diff_b200 as pdb
from pulser_diff_b200.samples import ChannelSamples, SequenceSamples
from pulser_diff_b200.utils import interpolate_sine

dev = torch.device("cuda", 0)
interp = interpolate_sine(B.N_PARAM, B.DURATION).to(torch.float64)
coords = B.chain_coords(B.N_QUBITS)
register = {f"q{i}": coords[i] for i in range(B.N_QUBITS)}
spec = pdb.DeviceSpec(B.C6)
diag = B.loss_diag(B.N_QUBITS, dev)
ta, td = B.workload_params(0)

def one():
    amp, det, ph = B.pulse_samples(ta, td, interp)
    em = pdb.TorchEmulator(SequenceSamples([ChannelSamples(amp, det, ph)]), register, spec, sampling_rate=B.RATE, torch_device=dev)
    res = em.run(solver=pdb.SolverType.DP5_SE)
    loss = res.expect([diag])[0].real[-1]
    ga, gd = torch.autograd.grad(loss, [ta, td])
    return float(loss), ga, gd

for _ in range(5):
    one()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    one()
torch.cuda.synchronize()
print("ms per pass", (time.perf_counter() - t0) * 1e3 / 20)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    one()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)
