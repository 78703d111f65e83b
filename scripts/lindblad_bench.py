"""configs[3]: noisy Lindblad evolution (DP5_ME) of an N-atom chain with dephasing + relaxation,
density matrix 4^N, forward + adjoint gradient.  Prints steps/s and the achieved algorithmic GB/s
(528 B per density-matrix entry per DP5 step).

    python scripts/lindblad_bench.py 10 12
"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench as B
import pulser_diff_b200 as pdb
from pulser_diff_b200.samples import ChannelSamples, SequenceSamples
from pulser_diff_b200.utils import interpolate_sine, occupation_diag

dev = torch.device("cuda", 0)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
for n in [int(a) for a in (sys.argv[1:] or ["10"])]:
    interp = interpolate_sine(B.N_PARAM, B.DURATION).to(torch.float64)
    coords = B.chain_coords(n)
    register = {f"q{i}": coords[i] for i in range(n)}
    ta, td = B.workload_params(0)
    amp, det, ph = B.pulse_samples(ta, td, interp)
    cfg = pdb.SimConfig(noise=("dephasing", "relaxation"), dephasing_rate=0.5, relaxation_rate=0.1)
    em = pdb.TorchEmulator(SequenceSamples([ChannelSamples(amp, det, ph)]), register, pdb.DeviceSpec(B.C6),
                           sampling_rate=B.RATE, config=cfg, torch_device=dev)
    em.set_evaluation_times("Minimal")
    obs = torch.zeros(2 ** n, dtype=torch.float64, device=dev)
    for i in range(n):
        obs = obs + occupation_diag(n, [i], dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = em.run()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss = res.expect([obs])[0].real[-1]
    ga, gd = torch.autograd.grad(loss, [ta, td])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    log = em._last_result.step_log()
    acc = sum(1 for r in log if r["accepted"])
    entries = 4 ** n
    rho = res.states.detach()[-1, :, :, 0]
    print(json.dumps({"config": "C4 lindblad", "n": n, "attempts": len(log), "accepted": acc,
                      "fwd_s": t1 - t0, "bwd_s": t2 - t1, "fwd_steps_per_s": len(log) / (t1 - t0),
                      "fwd_alg_GBs": 528.0 * entries * len(log) / (t1 - t0) / 1e9,
                      "fwd_frac_of_peak": 528.0 * entries * len(log) / (t1 - t0) / 1e9 / peak,
                      "trace": torch.trace(rho).real.item(), "loss": loss.item(),
                      "grad_norm": float(torch.cat([ga, gd]).norm())}), flush=True)
    del res, em, rho
    torch.cuda.empty_cache()
