#!/usr/bin/env python
"""Benchmark of the pulser-diff evolution hot path on B200 (driver contract: see README/DESIGN).

    python bench.py --gpus N --steps K --warmup W            # B200 arm
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path (oracle port)

Workload (BASELINE.json configs[1]): 12-atom 1-D chain, antiferromagnetic state preparation with
parametrised Omega(t)/delta(t) (docs/state_preparation.ipynb parametrisation: 30+30 control
points through interpolate_sine, 1100 ns, sampling_rate 0.05, DP5_SE), loss built from
<n_i n_{i+1}> and <n_i>, gradient w.r.t. the 60 parameters.  One bench "step" = one forward +
gradient pass.  Metric = DP5 evolution steps per second over forward+gradient.

* ``value``  : passes with the register state and pulse coefficients already on the device.
* ``e2e``    : the same pass through ``TorchEmulator.run`` from HOST tensors, loss and gradients
               read back to the host, every step.
* ``roofline``: the HBM-bound regime of the same kernels -- one fused DP5 step at north_star's N = 26
               (1 GiB vectors; stream family: A tiles + first bit group as one L2-blocked dataflow launch,
               last bit group as a second launch) with a full-register global drive, timed with CUDA
               events inside the C ABI on the launching stream; algorithmic bytes 576 B x 2^N per step,
               40 B x 2^N per H.psi (``roofline_hpsi``), SURVEY.md 8d.  ``roofline_n23`` is the same
               family at N = 23 (128 MiB vectors); ``fwd_grad_n26`` is forward + adjoint gradient wall time per DP5 step at
               N = 26 (BASELINE metric "fwd+grad wall time at N qubits") with ``roofline_adjoint``.
* ``c3_batch``: BASELINE configs[2] -- 4096 pulse-parameter sets of the 2-atom gate workload, forward +
               gradient through ``ops.evolve_units``, sets dealt over the ranks (strong scaling, no
               data-path collective); coefficient tables are built on the device.
* ``c4_lindblad``: BASELINE configs[3] -- 12-atom Lindblad run (dephasing + relaxation, 4^12 density
               matrix), forward + adjoint gradient seconds and the DP5_ME step roofline (528 B/entry).
* ``cpu_baseline``: the oracle (restated reference CPU path: sparse-COO H(t) re-assembly, DP5,
               tape autograd) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_QUBITS, DURATION, N_PARAM, GAMMA, RATE = 12, 1100, 30, 0.02, 0.05
MAX_AMP, MAX_DET, SPACING, C6 = 12, 6, 7.0, 865723.02


# ------------------------------------------------------------------------------------------------
def workload_params(seed: int):
    g = torch.Generator().manual_seed(seed)
    ta = ((2 * torch.rand(N_PARAM, dtype=torch.float64, generator=g) - 1) * 50).requires_grad_(True)
    td = ((2 * torch.rand(N_PARAM, dtype=torch.float64, generator=g) - 1) * 50).requires_grad_(True)
    return ta, td


def pulse_samples(ta, td, interp):
    amp = interp @ (MAX_AMP * torch.sigmoid(GAMMA * ta))
    det = interp @ (MAX_DET * torch.tanh(GAMMA * td))
    return amp, det, torch.zeros(DURATION, dtype=torch.float64)


def chain_coords(n):
    return torch.stack([torch.arange(n, dtype=torch.float64) * SPACING,
                        torch.zeros(n, dtype=torch.float64)], dim=1)


def loss_diag(n, device):
    from pulser_diff_b200.utils import occupation_diag
    d = torch.zeros(2 ** n, dtype=torch.float64, device=device)
    for i in range(n - 1):
        d = d + occupation_diag(n, [i, i + 1], device)
    for i in range(0, n, 2):
        d = d - occupation_diag(n, [i], device)
    return d


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions: one in-process NVML query after every timed
    pass, from the timing thread itself.  (A background sampler -- `nvidia-smi` every 200 ms, or NVML from
    a second thread -- contends for driver locks with the cooperative launches and stalled single passes
    of this 10-ms workload by 20-200 ms; the query below costs ~0.1 ms and is inside the timed region.)"""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.sm, self.reasons, self.max = index, [], set(), 0.0
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def start(self):
        self.sample()

    def sample(self):
        try:
            if self._nvml is not None:
                nv = self._nvml
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.reasons |= {n for n in self.NAMES if mask & self.BITS[n]}
            else:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                r = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(r[0])); self.max = max(self.max, float(r[1]))
                self.reasons |= {n for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")}
        except Exception:
            pass

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max or None,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml (one query per timed pass)" if self._nvml is not None else "nvidia-smi"}


def measured_traffic(key: str):
    """dram bytes read+written per DP5 step from the committed ncu capture (profiles/r02_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[key]
    except Exception:
        return None


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def fwd_grad_large(dev, n, n_steps, peak, cdtype=torch.complex128):
    """Forward + adjoint-gradient wall time of n_steps fixed DP5 steps at n qubits (BASELINE metric
    "fwd+grad wall time at N qubits"), full-register global drive.  Algorithmic bytes of the adjoint of one
    step: 576 B/amplitude to recompute the stage inputs + 576 B for the six
    transposed generator applications with their stage combinations + 6 x 32 B for the site correlations
    (each reads the stage adjoint and the stage input once) = 1344 B/amplitude."""
    from pulser_diff_b200 import _cabi, ops
    ops.clear_plan_cache()
    torch.cuda.empty_cache()
    T = 64
    g = torch.Generator().manual_seed(0)
    dv0 = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4
    av0 = torch.complex(torch.rand(1, T, dtype=torch.float64, generator=g) * 3, torch.zeros(1, T, dtype=torch.float64))
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = C6 / (SPACING * (j - i)) ** 6
    full = (1 << n) - 1
    psi0 = torch.zeros(1, 2 ** n, dtype=cdtype, device=dev)
    psi0[0, -1] = 1.0
    tsave = torch.tensor([0.0, 0.001 * n_steps], dtype=torch.float64)
    replay = [(0.001, False)] * (n_steps - 1) + [(0.001, True)]
    w = torch.arange(2 ** n, device=dev).remainder(7).to(torch.float64)
    out = None
    for _ in range(3):      # the last sweep runs warm (slope cache, allocations)
        dv = dv0.clone().requires_grad_(True)
        av = av0.clone().requires_grad_(True)
        torch.cuda.synchronize(dev)
        torch.cuda.nvtx.range_push("fwd_n%d" % n)
        t0 = time.perf_counter()
        st = ops.evolve(psi0, tsave, dv, av, u, n_qubits=n, kind=_cabi.PD_KET, dt=0.02, det_masks=[full],
                        amp_masks=[full], options=_cabi.Options(replay=replay))
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        torch.cuda.nvtx.range_pop()
        val = (w * st[-1, 0].abs() ** 2).sum()
        torch.cuda.nvtx.range_push("adjoint_n%d" % n)
        torch.autograd.grad(val, [dv, av])
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        torch.cuda.nvtx.range_pop()
        fwd, bwd = (t1 - t0) * 1e3 / n_steps, (t2 - t1) * 1e3 / n_steps
        amps = 2.0 ** n * (1.0 if cdtype == torch.complex128 else 0.5)     # bytes scale with the amplitude size
        out = {"workload": f"chain_n{n}: {n_steps} fixed DP5 steps forward + adjoint gradient w.r.t. all samples",
               "state_dtype": str(cdtype).replace("torch.", ""),
               "fwd_ms_per_step": fwd, "adjoint_ms_per_step": bwd, "fwd_grad_ms_per_step": fwd + bwd,
               "adjoint_over_forward": bwd / fwd,
               "roofline_adjoint": {"bound": "hbm", "achieved": 1344.0 * amps / (bwd * 1e-3) / 1e9, "peak": peak,
                                    "unit": "GB/s", "frac": 1344.0 * amps / (bwd * 1e-3) / 1e9 / peak,
                                    "algorithmic_bytes": 1344.0 * amps, "traffic": None}}
        del st, val
    ops.clear_plan_cache()
    torch.cuda.empty_cache()
    return out


C3 = dict(n=2, pulses=8, dur=131, rate=0.05, c6=865723.02, spacing=6.5, n_sets=4096)


def c64_hpsi_check(dev, H, dv, av, n):
    """max |H psi (complex64 library) - H psi (complex128 library)| / max |H psi| on one random vector."""
    from pulser_diff_b200 import _cabi
    outs = []
    psi = torch.randn(1, 2 ** n, dtype=torch.complex128, device=dev)
    psi /= psi.norm()
    for cd in (torch.complex128, torch.complex64):
        plan = _cabi.Plan(n, 1, _cabi.PD_KET, dev, cd)
        cu = torch.zeros(n, n, dtype=torch.float64)
        for i in range(n):
            for j in range(i + 1, n):
                cu[i, j] = C6 / (SPACING * (j - i)) ** 6
        plan.set_interaction(cu)
        plan.set_terms(H.dt, [(1 << n) - 1], dv, [(1 << n) - 1], av)
        outs.append(plan.hpsi(0.3, psi.to(cd)).to(torch.complex128))
        del plan
    return float((outs[1] - outs[0]).abs().max() / outs[0].abs().max())


def c3_tables(params, idx):
    """params (U, 3, pulses) on any device -> the reference's coefficient arrays 0.5*amp*exp(-i phase) and
    -0.5*det, sub-sampled (hamiltonian.py:83-91, 419-423), as (U, 1, n_samples)."""
    U = params.shape[0]
    z = torch.zeros(U, 1, dtype=torch.float64, device=params.device)
    amp, det, ph = (torch.cat([params[:, k].repeat_interleave(C3["dur"], dim=1), z], dim=1)[:, idx] for k in range(3))
    return (-0.5 * det)[:, None, :], (0.5 * amp * torch.exp(-1j * ph))[:, None, :]


def c3_setup(dev):
    import math
    D = C3["pulses"] * C3["dur"] + 1                     # one extra sample (reference backend.py:114-115)
    idx = torch.linspace(0, D - 1, int(C3["rate"] * D), dtype=torch.int).long()
    tsave = torch.cat([(torch.arange(D, dtype=torch.float64) / 1000)[idx],
                       torch.tensor([0.0, (D - 1) / 1000], dtype=torch.float64)]).unique()
    pair_u = torch.zeros(2, 2, dtype=torch.float64)
    pair_u[0, 1] = C3["c6"] / C3["spacing"] ** 6
    had = torch.tensor([[1, 1], [1, -1]], dtype=torch.complex128) / math.sqrt(2)
    return idx, tsave, pair_u, torch.kron(had, had)


def c3_batch(dev, rank, world, reps, barrier, weak=False):
    """BASELINE configs[2]: 4096 parameter sets of the 2-atom gate workload (8 constant pulses x 131 ns,
    psi0 = eye(4), rate 0.05, Hadamard x Hadamard infidelity, gradient w.r.t. the 24 parameters of every
    set), dealt over the ranks; tables, evolution, loss and gradients stay on the device."""
    import math
    from pulser_diff_b200 import ops, parallel
    idx, tsave, pair_u, target = c3_setup(dev)
    target = target.to(dev)
    g = torch.Generator().manual_seed(0)
    # weak: every rank brings a full batch of its own (n_sets x world sets in total)
    n_total = C3["n_sets"] * (world if weak else 1)
    params_all = torch.rand(n_total, 3, C3["pulses"], dtype=torch.float64, generator=g) * 4 * math.pi
    mine = parallel.shard_units(n_total, rank, world) if world > 1 else list(range(n_total))
    params = params_all[mine].to(dev).requires_grad_(True)
    idx_d = idx.to(dev)
    psi0 = torch.eye(4, dtype=torch.complex128, device=dev).repeat(len(mine), 1, 1)
    dt = 0.001 / C3["rate"]

    def sweep():
        dv, av = c3_tables(params, idx_d)
        st = ops.evolve_units(psi0, tsave, dv, av, pair_u, n_qubits=2, dt=dt, det_masks=[3], amp_masks=[3])
        Uf = st[:, -1].transpose(1, 2)
        loss = 1 - (target.conj().T @ Uf).diagonal(dim1=1, dim2=2).sum(-1).abs() / 4
        (gp,) = torch.autograd.grad(loss.sum(), [params])
        return loss.detach(), gp

    for _ in range(2):
        loss, gp = sweep()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        loss, gp = sweep()
    barrier()
    secs = (time.perf_counter() - t0) / reps
    return secs, float(loss.sum()), float(gp.abs().sum()), len(mine)


def c4_lindblad(dev, peak):
    """BASELINE configs[3]: 12-atom chain, C2 pulses, dephasing 0.5 + relaxation 0.1, rho0 = |g..g><g..g|,
    tsave Minimal, loss tr(rho sum n_i); forward + adjoint gradient through TorchEmulator, and one DP5_ME
    step against the HBM roofline (528 B per density-matrix entry)."""
    import pulser_diff_b200 as pdb
    from pulser_diff_b200 import ops
    from pulser_diff_b200.samples import ChannelSamples, SequenceSamples
    from pulser_diff_b200.utils import interpolate_sine, occupation_diag
    ops.clear_plan_cache()
    torch.cuda.empty_cache()
    n = N_QUBITS
    interp = interpolate_sine(N_PARAM, DURATION).to(torch.float64)
    coords = chain_coords(n)
    register = {f"q{i}": coords[i] for i in range(n)}
    obs = torch.zeros(2 ** n, dtype=torch.float64, device=dev)
    for i in range(n):
        obs = obs + occupation_diag(n, [i], dev)
    cfg = pdb.SimConfig(noise=("dephasing", "relaxation"), dephasing_rate=0.5, relaxation_rate=0.1)
    out = None
    for _ in range(2):       # second pass warm (NVRTC compiles torch's complex element-wise kernels once)
        ta, td = workload_params(0)
        amp, det, ph = pulse_samples(ta, td, interp)
        em = pdb.TorchEmulator(SequenceSamples([ChannelSamples(amp, det, ph)]), register, pdb.DeviceSpec(C6),
                               sampling_rate=RATE, config=cfg, torch_device=dev)
        em.set_evaluation_times("Minimal")
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res = em.run()
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        loss = res.expect([obs])[0].real[-1]
        ga, gd = torch.autograd.grad(loss, [ta, td])
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        log = em._last_result.step_log()
        entries = 4.0 ** n
        ach = 528.0 * entries * len(log) / (t1 - t0) / 1e9
        out = {"workload": "C4: 12-atom chain, dephasing 0.5 + relaxation 0.1, DP5_ME, 4^12 density matrix",
               "attempted_steps": len(log), "accepted_steps": sum(1 for r in log if r["accepted"]),
               "fwd_s": t1 - t0, "adjoint_s": t2 - t1, "fwd_grad_s": t2 - t0, "adjoint_over_forward": (t2 - t1) / (t1 - t0),
               "steps_per_s_forward": len(log) / (t1 - t0), "loss": float(loss),
               "trace": float(torch.trace(res.states.detach()[-1, :, :, 0]).real),
               "roofline_c4": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                               "algorithmic_bytes": 528.0 * entries, "traffic": None,
                               "kernel": "DP5_ME step (6 Lindblad applications on vec(rho) + stage combines + error norm), "
                                         "whole forward run incl. the controller"}}
        del res, em
    ops.clear_plan_cache()
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import pulser_diff_b200 as pdb
    from pulser_diff_b200 import _cabi, ops
    from pulser_diff_b200.samples import ChannelSamples, SequenceSamples
    from pulser_diff_b200.utils import interpolate_sine

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    interp = interpolate_sine(N_PARAM, DURATION).to(torch.float64)
    coords = chain_coords(N_QUBITS)
    register = {f"q{i}": coords[i] for i in range(N_QUBITS)}
    device_spec = pdb.DeviceSpec(C6)
    diag_dev = loss_diag(N_QUBITS, dev)
    ta, td = workload_params(args.seed + rank)           # one parameter set per rank (weak scaling)
    plan = ops.get_plan(N_QUBITS, 1, _cabi.PD_KET, dev)

    def e2e_pass(count_steps=False):
        """Public API from host tensors: samples -> TorchEmulator.run -> loss/grad on the host."""
        amp, det, ph = pulse_samples(ta, td, interp)
        em = pdb.TorchEmulator(SequenceSamples([ChannelSamples(amp, det, ph)]), register, device_spec,
                               sampling_rate=RATE, torch_device=dev)
        res = em.run(solver=pdb.SolverType.DP5_SE)
        loss = res.expect([diag_dev])[0].real[-1]
        ga, gd = torch.autograd.grad(loss, [ta, td])
        # the accepted-step count is a property of the workload: read it once, outside the timed region
        steps = sum(1 for r in em._last_result.step_log() if r["accepted"]) if count_steps else None
        return float(loss), ga, gd, steps, em

    # device-resident variant: Hamiltonian structure + psi0 prepared once, outside the timed region
    _, _, _, n_steps, em0 = e2e_pass(count_steps=True)
    H = em0._hamiltonian._hamiltonian
    dm, dv, am, av = H.masks_and_values()
    dv_d = dv.detach().clone().requires_grad_(True)
    av_d = av.detach().clone().requires_grad_(True)
    pair_u = H.pair_u.detach()
    psi0_dev = em0.initial_state.to(dev).transpose(0, 1).contiguous()
    tsave = em0.evaluation_times.detach()

    def resident_pass():
        st = ops.evolve(psi0_dev, tsave, dv_d, av_d, pair_u, n_qubits=N_QUBITS, kind=_cabi.PD_KET,
                        dt=H.dt, det_masks=dm, amp_masks=am, solver=_cabi.SOLVER_DP5_SE)
        loss = torch.ops.pulser_diff_b200.expect_diag  # noqa: F841 (op exists)
        from pulser_diff_b200.utils import expect_diag
        val = expect_diag(diag_dev, st.permute(0, 2, 1)).real[-1]
        torch.autograd.grad(val, [dv_d, av_d])
        return val

    for _ in range(args.warmup):
        resident_pass()
        e2e_pass()

    sampler = ClockSampler(local)
    sampler.start()
    # ---- timed region 1: device-resident passes ----
    l0 = plan.launch_count
    barrier()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        resident_pass()
        sampler.sample()
    ev1.record()
    barrier()
    t_res = time.perf_counter() - t0
    launches = plan.launch_count - l0
    # ---- timed region 2: end-to-end passes from host tensors ----
    barrier()
    _cabi.transfer_counters(reset=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss_val, ga, gd, _, _ = e2e_pass()
        sampler.sample()
    barrier()
    t_e2e = time.perf_counter() - t0
    lib_h2d, lib_d2h = _cabi.transfer_counters()

    # ---- roofline: the HBM-bound regime, N = roofline_n, CUDA events inside the C ABI ----
    roof = roof_h = roof23 = roof_c64 = None
    peak, peak_src = measured_peak()
    if rank == 0 and args.roofline_n > 0:
        def measure(nr, cdtype=torch.complex128):
            # algorithmic bytes per amplitude: DP5 step = (26 reads + 7 writes) amplitudes + 6
            # diagonal reals = 576 (complex128) / 288 (complex64); H.psi = read + write + one diagonal real = 40 / 20
            ab = 16.0 if cdtype == torch.complex128 else 8.0
            b_step, b_h = 33.0 * ab + 3.0 * ab, 2.0 * ab + 0.5 * ab
            ops.clear_plan_cache()
            big = _cabi.Plan(nr, 1, _cabi.PD_KET, dev, cdtype)
            cu = torch.zeros(nr, nr, dtype=torch.float64)
            for i in range(nr):
                for j in range(i + 1, nr):
                    cu[i, j] = C6 / (SPACING * (j - i)) ** 6
            big.set_interaction(cu)
            full_mask = [(1 << nr) - 1]           # the global channel drives EVERY atom of the big register
            big.set_terms(H.dt, full_mask, dv, full_mask, av)
            y = torch.zeros(1, 2 ** nr, dtype=cdtype, device=dev)
            y[0, -1] = 1.0
            ms_step = big.bench_dp5_steps(0.3, 1e-3, args.roofline_steps, y)
            psi = torch.randn(1, 2 ** nr, dtype=torch.float64, device=dev).to(cdtype)
            ms_h = big.bench_hpsi(0.3, psi, max(4, args.roofline_steps * 3))
            s_amp = 2 ** nr
            ach = b_step * s_amp / (ms_step * 1e-3) / 1e9
            ach_h = b_h * s_amp / (ms_h * 1e-3) / 1e9
            family = ("stream (one bit-group of H per tile type, >= 256 B pieces, A + first group as one "
                      "L2-blocked dataflow launch)" if nr >= 19 else "gather")
            r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                 "traffic": measured_traffic(f"dp5_step_n{nr}") if cdtype == torch.complex128 else None,
                 "peak_source": peak_src, "state_dtype": str(cdtype).replace("torch.", ""),
                 "kernel": "DP5 step kernel sequence (6 generator applications + stage combines + error norm)",
                 "kernel_family": family, "workload": f"chain_n{nr}_dp5_step", "ms_per_launch": ms_step,
                 "algorithmic_bytes": b_step * s_amp, "steps_per_s": 1e3 / ms_step}
            rh = {"bound": "hbm", "achieved": ach_h, "peak": peak, "unit": "GB/s", "frac": ach_h / peak,
                  "traffic": None, "kernel": "H(t) psi", "workload": f"chain_n{nr}_hpsi",
                  "ms_per_launch": ms_h, "algorithmic_bytes": b_h * s_amp}
            del big, y, psi
            torch.cuda.empty_cache()
            return r, rh
        roof, roof_h = measure(args.roofline_n)
        if args.roofline_n != 23 and not args.skip_n23:
            roof23, _ = measure(23)
        if not args.quick:
            # north_star's optional complex64 tier, reported BESIDE the complex128 numbers (never instead of):
            # same kernels compiled for complex64 state vectors (libpulser_diff_b200_c64.so), half the bytes per pass
            try:
                r64, r64h = measure(args.roofline_n, torch.complex64)
                roof_c64 = {"dp5_step": r64, "hpsi": r64h, "speedup_vs_c128_step": roof["ms_per_launch"] / r64["ms_per_launch"],
                            "hpsi_rel_diff_vs_c128_n22": c64_hpsi_check(dev, H, dv, av, 22)}
            except Exception as exc:
                roof_c64 = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    clocks = sampler.summary()
    fwd_grad = c4 = None
    if rank == 0 and world == 1 and args.roofline_n > 0 and not args.quick:
        try:
            fwd_grad = fwd_grad_large(dev, args.roofline_n, 4, peak)
        except Exception as exc:           # the headline must survive an out-of-memory here
            fwd_grad = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            c4 = c4_lindblad(dev, peak)
        except Exception as exc:
            c4 = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    c3 = None
    if not args.quick:
        try:
            c3_s, c3_loss, c3_gsum, c3_mine = c3_batch(dev, rank, world, 3, barrier)
            tt = torch.tensor([c3_s], dtype=torch.float64, device=dev)
            chk = torch.tensor([c3_loss, c3_gsum], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dist.all_reduce(chk, op=dist.ReduceOp.SUM)
            c3 = {"workload": "C3: 4096 pulse-parameter sets, 2-atom gate (8 x 131 ns constant pulses, psi0 = eye(4), "
                              "rate 0.05), forward + gradient w.r.t. 24 parameters per set",
                  "n_sets": C3["n_sets"], "sets_per_rank": c3_mine, "scaling": "strong",
                  "seconds_per_sweep": tt.item(), "sets_per_s": C3["n_sets"] / tt.item(),
                  "sum_loss": chk[0].item(), "sum_abs_grad": chk[1].item(),
                  "tables": "built on the device from the parameters (no host tables cross the ABI)"}
            if world > 1:
                # one sweep of 4096/world sets is a fraction of ONE wave of per-set latency chains (16 one-warp
                # units per SM), so the strong-scaling time is bounded below by a single set's forward + adjoint
                # chain; with a full batch per GPU the same kernels scale with the device count:
                w_s, _, _, w_mine = c3_batch(dev, rank, world, 3, barrier, weak=True)
                wt = torch.tensor([w_s], dtype=torch.float64, device=dev)
                dist.all_reduce(wt, op=dist.ReduceOp.MAX)
                c3["weak"] = {"n_sets": C3["n_sets"] * world, "sets_per_rank": w_mine, "seconds_per_sweep": wt.item(),
                              "sets_per_s": C3["n_sets"] * world / wt.item()}
        except Exception as exc:
            c3 = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # max over ranks
    t = torch.tensor([t_res, t_e2e], dtype=torch.float64, device=dev)
    steps_total = torch.tensor([float(n_steps)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(steps_total, op=dist.ReduceOp.SUM)
    t_res, t_e2e = t.tolist()
    total_steps = steps_total.item() * args.steps
    value = total_steps / t_res
    e2e_value = total_steps / t_e2e

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(args.seed, budget_s=args.cpu_budget)
        if not args.quick:
            try:
                cpu["extra"] = cpu_extras()
            except Exception as exc:
                cpu["extra"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    if rank == 0:
        s12 = 2 ** N_QUBITS
        n_s, n_tp = int(dv.shape[-1]), len(tsave)
        # bytes that cross PCIe in one e2e pass: every cudaMemcpyAsync call site of the library is counted
        # (pd_transfer_counters, read around the timed region); torch adds the initial state (host -> device)
        # and the loss scalar (device -> host)
        h2d = int(lib_h2d // args.steps + s12 * 16)
        d2h = int(lib_d2h // args.steps + 16)
        line = {
            "metric": "evolution steps/sec (DP5 steps, forward+gradient)", "value": value,
            "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
            "config": {"workload": "C2: 12-atom chain state preparation, DP5_SE fwd + gradient of "
                                   "<n_i n_j> loss w.r.t. 60 pulse parameters",
                       "n_qubits": N_QUBITS, "duration_ns": DURATION, "sampling_rate": RATE,
                       "n_params": 2 * N_PARAM, "dp5_steps_per_pass": n_steps,
                       "parameter_sets": world, "parallelism": f"independent parameter sets x{world}",
                       "l2": "state (64 KiB) is smaller than L2 by construction of the workload; "
                             "the roofline blocks use N=26 / N=23 (1 GiB / 128 MiB vectors, >> 126 MB L2)"},
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_pass": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof, "roofline_hpsi": roof_h, "roofline_n23": roof23, "roofline_c64": roof_c64,
            "fwd_grad_n26": fwd_grad, "c3_batch": c3, "c4_lindblad": c4,
            "roofline_workload": {"bound": "latency", "achieved": 576.0 * s12 * total_steps / world / t_res / 1e9 * 2,
                                  "peak": peak, "unit": "GB/s",
                                  "note": "N=12: 64 KiB vectors live in registers/L2 (one cooperative kernel per sweep); "
                                          "2x = forward + adjoint"},
            "cpu_baseline": cpu,
            "sharded": None,
            "loss": loss_val,
        }
    # configs[4] leg (N > 1 only): one register sharded over all ranks, reported beside the
    # headline.  It must never cost the headline line: a watchdog prints the line without it.
    if world > 1 and world & (world - 1) == 0 and args.sharded_local_qubits > 0:
        import threading
        done = threading.Event()

        def give_up():
            if done.is_set():
                return
            if rank == 0:
                line["sharded"] = {"error": "sharded leg did not finish within 300 s"}
                print(json.dumps(line), flush=True)
            os._exit(0)

        timer = threading.Timer(420.0, give_up)
        timer.daemon = True
        timer.start()
        try:
            sharded = sharded_measure(dev, rank, world, args.sharded_local_qubits, 6, 1)
        except Exception as exc:
            sharded = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        # the same leg one qubit larger in the complex64 tier: 2^30 amplitudes per GPU = the N = 33 register of
        # configs[4] on eight GPUs, in the bytes the complex128 leg needs for N = 32
        if "error" not in sharded and not args.no_sharded_c64:
            try:
                sharded["complex64"] = sharded_measure(dev, rank, world, args.sharded_local_qubits + 1, 4, 1,
                                                       torch.complex64, parity=False)
            except Exception as exc:
                sharded["complex64"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        done.set()
        timer.cancel()
        if rank == 0:
            line["sharded"] = sharded
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_sample(seed: int, budget_s: float = 20.0):
    """The oracle on a bounded sample of the workload: forward + tape gradient over the first
    tsave intervals, all host threads."""
    from helpers import Channel, Problem
    from oracle.ref_solvers import SolverType, sesolve
    from pulser_diff_b200.utils import interpolate_sine
    torch.set_num_threads(os.cpu_count() or 1)
    interp = interpolate_sine(N_PARAM, DURATION).to(torch.float64)
    ta, td = workload_params(seed)
    amp, det, ph = pulse_samples(ta, td, interp)
    p = Problem(chain_coords(N_QUBITS), C6, [Channel(amp, det, ph)], rate=RATE)
    ref = p.ref()
    n_int = 3 if budget_s >= 10.0 else 1       # ~10-12 s of CPU work on 16 host threads
    t0 = time.perf_counter()
    ts = ref.evaluation_times[: n_int + 1].clone()
    res = sesolve(ref.ham.H, ref.initial_state, ts, SolverType.DP5_SE, {})
    d = loss_diag(N_QUBITS, "cpu")
    loss = (d[:, None] * res.states[-1].abs() ** 2).sum()
    torch.autograd.grad(loss, [ta, td])
    dt = time.perf_counter() - t0
    steps = sum(1 for r in res.steplog if r[2])
    return {"value": steps / dt, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {n_int} of 55 tsave intervals of the same N=12 workload "
                      f"({steps} DP5 steps forward + tape gradient, {dt:.1f} s)",
            "seconds": dt, "dp5_steps": steps}


def cpu_extras():
    """the CPU path on C1 (2 atoms, forward + gradient) and the cost of one
    H(t) re-assembly + sparse H.psi per register size (what every Runge-Kutta stage pays on the CPU)."""
    import math
    from helpers import Channel, Problem
    from oracle.ref_solvers import SolverType
    torch.set_num_threads(os.cpu_count() or 1)
    amp = torch.full((1000,), 5.0, dtype=torch.float64, requires_grad=True)
    z = torch.zeros(1000, dtype=torch.float64)
    p = Problem(torch.tensor([[0.0, 0.0], [8.0, 0.0]], dtype=torch.float64), 5420158.53, [Channel(amp, z, z)],
                rate=1.0, evaluation_times="Minimal")
    ref = p.ref()
    t0 = time.perf_counter()
    res = ref.run(solver=SolverType.DP5_SE)
    loss = (res.states[-1].abs() ** 2)[0].sum()
    torch.autograd.grad(loss, [amp])
    c1_s = time.perf_counter() - t0
    c1_steps = sum(1 for r in res.steplog if r[2])
    sweep = []
    for n in (8, 10, 12):
        zz = torch.zeros(201, dtype=torch.float64)
        pr = Problem(chain_coords(n), C6, [Channel(zz + 3.0, zz - 1.0, zz)], rate=0.05).ref()
        psi = torch.randn(2 ** n, 1, dtype=torch.complex128)
        pr.ham.H(0.05) @ psi
        reps = 5 if n < 12 else 2
        t0 = time.perf_counter()
        for _ in range(reps):
            Hm = pr.ham.H(0.05)
        t1 = time.perf_counter()
        for _ in range(reps):
            Hm @ psi
        t2 = time.perf_counter()
        sweep.append({"n": n, "assemble_ms": (t1 - t0) * 1e3 / reps, "spmm_ms": (t2 - t1) * 1e3 / reps})
    return {"c1": {"workload": "C1: 2 atoms, constant pulse 1000 ns, rate 1.0, DP5_SE forward + gradient",
                   "seconds": c1_s, "dp5_steps": c1_steps, "steps_per_s": c1_steps / c1_s},
            "hpsi_per_call_ms": sweep,
            "infeasible": "N >= 18 on the CPU path: COO H(t) has ~2^N (N+1) entries of 32 B re-assembled at every "
                          "stage (N=26: 58 GB per assembly) plus a tape of 7 stage vectors per step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_sample(args.seed, budget_s=args.cpu_budget)
        if i >= args.warmup:
            vals.append(last)
    tot_steps = sum(v["dp5_steps"] for v in vals)
    tot_s = sum(v["seconds"] for v in vals)
    value = tot_steps / tot_s
    last = dict(last, value=value)
    print(json.dumps({
        "impl": "reference", "metric": "evolution steps/sec (DP5 steps, forward+gradient)",
        "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
        "config": {"workload": "C2: 12-atom chain state preparation, DP5_SE fwd + gradient "
                               "(bounded sample: first tsave interval per step)",
                   "n_qubits": N_QUBITS, "duration_ns": DURATION, "sampling_rate": RATE},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference CPU path = oracle port (pyqtorch/pulser not installable: SURVEY.md 8c)",
    }))


NVLINK_PEER_GBS = 770.0     # fallback only: the peer-copy rate is measured in the same run (nvlink_measured)


def sharded_parity(dev, rank, world, n=20):
    """ShardedKet.evolve / evolve_backward over NVLink peer memory against the single-GPU engine on the
    full register (N = 20, shared step sequence): max state difference and relative gradient differences,
    max over ranks.  The driver's pytest box has one GPU, so this check lives in the bench."""
    import torch.distributed as dist
    from pulser_diff_b200 import _cabi, ops, parallel
    gen = torch.Generator().manual_seed(9)
    T = 40
    dv = (torch.rand(2, T, dtype=torch.float64, generator=gen) - 0.5) * 4
    av = torch.complex(torch.rand(2, T, dtype=torch.float64, generator=gen) * 3,
                       torch.rand(2, T, dtype=torch.float64, generator=gen) - 0.5)
    full = (1 << n) - 1
    masks = [full, 0b101 | (1 << (n - 1))]            # a global term + one touching global AND local qubits
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = C6 / (SPACING * (j - i)) ** 6
    psi0 = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=gen)
    psi0 = (psi0 / psi0.norm()).to(dev)
    tsave = torch.tensor([0.0, 0.004, 0.009], dtype=torch.float64)
    w = torch.rand(len(tsave), 1, 2 ** n, dtype=torch.float64, generator=gen).to(dev)
    leaves = [x.clone().requires_grad_(True) for x in (psi0, dv, av, u)]
    st_full = ops.evolve(leaves[0], tsave, leaves[1], leaves[2], leaves[3], n_qubits=n, kind=_cabi.PD_KET,
                         dt=0.002, det_masks=masks, amp_masks=masks)
    loss = (w * st_full.abs() ** 2).sum()
    g_full = torch.autograd.grad(loss, leaves)
    replay = [(r["t"], r["dt"], r["interval"], bool(r["clipped"])) for r in ops.last_step_log(st_full) if r["accepted"]]
    sk = parallel.ShardedKet(n, u, 0.002, masks, dv, masks, av, dev, peer_memory=True)
    n_loc = 2 ** sk.nl
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    st, steps = sk.evolve(sk.local_slice(psi0), tsave.tolist(), replay=replay)
    state_err = (st - st_full.detach()[:, :, sl]).abs().max().item()
    st_leaf = st.clone().requires_grad_(True)
    (g_st,) = torch.autograd.grad((w[:, :, sl] * st_leaf.abs() ** 2).sum(), st_leaf)
    out = sk.evolve_backward(st, g_st, steps)
    rel = lambda a, b: ((a.cpu() - b.cpu()).abs().max() / b.abs().max().cpu()).item()
    errs = torch.tensor([state_err, rel(out["det"], g_full[1]), rel(out["amp"], g_full[2]), rel(out["pair"], g_full[3]),
                         rel(out["state0"], g_full[0][:, sl])], dtype=torch.float64, device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    del sk, st, st_full, st_leaf, out, w, leaves, g_full
    ops.clear_plan_cache()
    torch.cuda.empty_cache()
    e = errs.tolist()
    return {"n_qubits": n, "accepted_steps": len(replay), "state_max_abs_diff": e[0], "grad_det_rel": e[1],
            "grad_amp_rel": e[2], "grad_pair_rel": e[3], "grad_psi0_rel": e[4],
            "ok": bool(e[0] < 1e-10 and max(e[1:]) < 1e-8)}


def sharded_measure(dev, rank, world, local_qubits, steps, warmup, cdtype=torch.complex128, parity=True):
    """configs[4]-style leg: ONE register of local_qubits + log2(world) atoms sharded over the
    ranks by its top qubits (NVLink peer memory).  Times H.psi and fixed-size DP5 steps with CUDA
    events (max over ranks) and sets them against max(HBM, NVLink) rooflines; the NVLink rate is the
    peer copy timed in the same run with every rank pulling at once."""
    import torch.distributed as dist
    from pulser_diff_b200 import parallel
    g = world.bit_length() - 1
    parity = sharded_parity(dev, rank, world) if parity else None
    ab = 16.0 if cdtype == torch.complex128 else 8.0          # bytes per amplitude
    # largest slice that fits: the forward evolution holds y, 7 slopes, the next state, the peer-visible
    # buffer, g receive buffers, 2 saved states, the bench's own initial state and the plan's two scratch
    # vectors (one amplitude each) + the 8 B diagonal
    free = torch.tensor([torch.cuda.mem_get_info(dev)[0]], dtype=torch.float64, device=dev)
    dist.all_reduce(free, op=dist.ReduceOp.MIN)
    need = lambda nl: ((15 + g) * ab + 8) * 2.0 ** nl * 1.05
    while local_qubits > 20 and need(local_qubits) > free.item():
        local_qubits -= 1
    n = local_qubits + g
    T = 64
    gen = torch.Generator().manual_seed(0)
    dv = (torch.rand(1, T, dtype=torch.float64, generator=gen) - 0.5) * 4
    av = torch.complex(torch.rand(1, T, dtype=torch.float64, generator=gen) * 3, torch.zeros(1, T, dtype=torch.float64))
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = C6 / (SPACING * (j - i)) ** 6
    full = (1 << n) - 1
    sk = parallel.ShardedKet(n, u, 0.02, [full], dv, [full], av, dev, peer_memory=True, dtype=cdtype)
    psi = sk.state_buffer()
    psi.zero_()
    if rank == world - 1:
        psi[0, -1] = 1.0

    def timed(fn, reps):
        dist.barrier(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(warmup):
        sk.hpsi(0.3, psi)
    # NVLink rate of this box, this run: every rank pulls its g partner slices at once (what H.psi does)
    def pull():
        for k in range(g):
            sk._recv[k].copy_(sk._peer_bufs[rank ^ (1 << k)])
    pull()
    ms_pull = timed(pull, 3)
    nvlink = g * ab * 2 ** local_qubits / (ms_pull * 1e-3) / 1e9
    ms_h = timed(lambda: sk.hpsi(0.3, psi), max(steps, 3))
    y0 = torch.zeros(1, 2 ** local_qubits, dtype=cdtype, device=dev)
    if rank == world - 1:
        y0[0, -1] = 1.0                                  # all-ground register
    h = 1e-3
    fixed = [(0.3 + i * h, h, 1, i == steps - 1) for i in range(steps)]
    sk.evolve(y0, [0.3, 0.3 + steps * h], replay=[(0.3, h, 1, False), (0.3 + h, (steps - 1) * h, 1, True)])
    ms_e = timed(lambda: sk.evolve(y0, [0.3, 0.3 + steps * h], replay=fixed), 1) / steps
    peak, _ = measured_peak()
    amps = 2 ** local_qubits

    def roof(alg_bytes_hbm, link_bytes, ms):
        hbm_t, link_t = alg_bytes_hbm / (peak * 1e9), link_bytes / (nvlink * 1e9)
        return {"bound": "nvlink" if link_t > hbm_t else "hbm", "hbm_s": hbm_t, "nvlink_s": link_t,
                "frac": max(hbm_t, link_t) / (ms * 1e-3)}

    del sk
    from pulser_diff_b200 import ops as _ops
    _ops.clear_plan_cache()
    torch.cuda.empty_cache()
    return {"workload": f"single register N={n} sharded by its {g} top qubits (2^{local_qubits} amplitudes per GPU, "
                        f"{str(cdtype).replace('torch.', '')})",
            "exchange": "partner slices pulled by the copy engines from NVLink peer memory beside the local kernels; "
                        "one accumulate kernel (pd_sharded_accumulate)",
            "local_qubits": local_qubits, "bytes_per_vector_per_gpu": ab * amps,
            "ms_per_hpsi": ms_h, "hpsi_per_s": 1e3 / ms_h,
            "roofline_hpsi": roof(2.5 * ab * amps, g * ab * amps, ms_h),
            "ms_per_dp5_step": ms_e, "dp5_steps_per_s": 1e3 / ms_e,
            "roofline_dp5_step": roof(36 * ab * amps, 6 * g * ab * amps, ms_e),
            "peak_hbm_GBs": peak,
            "nvlink_measured": {"GBs_per_direction_per_gpu": nvlink, "how": f"{g} concurrent peer copies of one "
                                f"{ab * amps / 2 ** 30:.2f} GiB slice per rank, CUDA events, max over ranks",
                                "ms": ms_pull},
            "parity_vs_single_gpu": parity}


def run_sharded(args):
    """configs[4]-style run on its own: reports sharded DP5 steps per second."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = sharded_measure(dev, rank, world, args.local_qubits, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps({
            "metric": "sharded-register DP5 steps/sec", "value": res["dp5_steps_per_s"], "unit": "steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_dp5_step"],
            "higher_is_better": True, "scaling": "weak", "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
            "config": {"workload": res["workload"], "exchange": res["exchange"]},
            "roofline": res["roofline_dp5_step"], "sharded": res}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--roofline-n", type=int, default=26)
    ap.add_argument("--roofline-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-n23", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + roofline only (no fwd+grad N=26, C3, C4 legs)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--workload", default="c2", choices=["c2", "sharded"])
    ap.add_argument("--local-qubits", type=int, default=26)
    ap.add_argument("--no-sharded-c64", action="store_true", help="N > 1: skip the complex64 sharded leg")
    ap.add_argument("--sharded-local-qubits", type=int, default=29,
                    help="N > 1: also time one register of this many + log2(N) qubits sharded over the ranks (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "sharded":
        run_sharded(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
