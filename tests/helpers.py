"""Shared problem builders: the SAME spec feeds the oracle and the product."""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Optional

import torch

from oracle import ref_pulses as RP
from oracle.ref_emulator import RefEmulator
from oracle.ref_solvers import SolverType as RefSolver
import pulser_diff_b200 as pdb
from pulser_diff_b200.samples import ChannelSamples, SequenceSamples

C6_70, C6_60 = 5420158.53, 865723.02
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name: str):
    return json.load(open(os.path.join(GOLDEN, name)))


@dataclass
class Channel:
    amp: torch.Tensor
    det: torch.Tensor
    phase: torch.Tensor
    addressing: str = "Global"
    target: Optional[int] = None


@dataclass
class Problem:
    coords: torch.Tensor
    c6: float
    channels: list
    rate: float = 1.0
    noise: dict = field(default_factory=dict)
    evaluation_times: object = "Full"

    @property
    def n(self) -> int:
        return int(self.coords.shape[0])

    # ---- oracle side ---------------------------------------------------------------------
    def ref(self) -> RefEmulator:
        z = torch.zeros(1, dtype=torch.float64)

        def ext(c: Channel) -> dict:
            return {"amp": torch.cat([c.amp, z]), "det": torch.cat([c.det, z]),
                    "phase": torch.cat([c.phase, c.phase[-1:].detach()])}

        # Global and Local channels are SEPARATE terms (reference hamiltonian.py:177, 478-481)
        samples: dict = {}
        per: dict = {}
        for c in self.channels:
            e = ext(c)
            if c.addressing == "Global":
                samples["Global"] = e if "Global" not in samples else \
                    {k: samples["Global"][k] + e[k] for k in e}
            else:
                per[c.target] = e if c.target not in per else {k: per[c.target][k] + e[k] for k in e}
        if per:
            samples["Local"] = per
        return RefEmulator(self.coords, self.c6, samples, rate=self.rate, noise=self.noise,
                           evaluation_times=self.evaluation_times)

    # ---- product side -----------------------------------------------------------------------
    def emulator(self, device: torch.device) -> pdb.TorchEmulator:
        chans = [ChannelSamples(c.amp, c.det, c.phase, c.addressing,
                                None if c.addressing == "Global" else [f"q{c.target}"])
                 for c in self.channels]
        reg = {f"q{i}": self.coords[i] for i in range(self.n)}
        nz = self.noise
        kinds = []
        kw = {}
        if "dephasing_rate" in nz:
            kinds.append("dephasing"); kw["dephasing_rate"] = nz["dephasing_rate"]
        if "relaxation_rate" in nz:
            kinds.append("relaxation"); kw["relaxation_rate"] = nz["relaxation_rate"]
        if "depolarizing_rate" in nz:
            kinds.append("depolarizing"); kw["depolarizing_rate"] = nz["depolarizing_rate"]
        if "eff_noise" in nz:
            kinds.append("eff_noise")
            kw["eff_noise_rates"] = [r for r, _ in nz["eff_noise"]]
            kw["eff_noise_opers"] = [o for _, o in nz["eff_noise"]]
        cfg = pdb.SimConfig(noise=tuple(kinds), **kw)
        return pdb.TorchEmulator(SequenceSamples(chans), reg, pdb.DeviceSpec(self.c6), self.rate,
                                 cfg, self.evaluation_times, torch_device=device)


def chain(n: int, spacing: float = 7.0) -> torch.Tensor:
    return torch.stack([torch.arange(n, dtype=torch.float64) * spacing,
                        torch.zeros(n, dtype=torch.float64)], dim=1)


def global_channel(*pulses) -> Channel:
    """pulses: (amp, det, phase) triples, phase scalar."""
    amp = torch.cat([p[0] for p in pulses])
    det = torch.cat([p[1] for p in pulses])
    ph = torch.cat([torch.ones(p[0].numel(), dtype=torch.float64) * torch.as_tensor(p[2], dtype=torch.float64)
                    for p in pulses])
    return Channel(amp, det, ph)


def kat_problem(name: str) -> Problem:
    """The notebook setups of SURVEY.md 8c (see tests/golden/notebook_kats.json)."""
    two = torch.tensor([[-4.0, 0.0], [4.0, 0.0]], dtype=torch.float64)
    kb = [(RP.constant(1000, 5.0), RP.constant(1000, 0.0), 0.0),
          (RP.blackman(800, math.pi), RP.ramp(800, 5.0, 0.0), 0.0)]
    if name == "K-A":
        sq = torch.tensor([[0., 0.], [0., 8.], [8., 0.], [8., 8.]], dtype=torch.float64)
        return Problem(sq, C6_70, [global_channel(
            (RP.blackman(800, math.pi), RP.ramp(800, -5.0, 0.0), 0.0),
            (RP.constant(800, 5.0), RP.constant(800, 0.0), 0.0))], rate=0.1)
    if name == "K-B":
        return Problem(two, C6_70, [global_channel(*kb)], rate=0.5)
    if name == "K-C":
        kc = [kb[0], (RP.blackman(800, 3.14), RP.ramp(800, 5.0, 0.0), 0.0)]
        return Problem(torch.tensor([[0.5, 0.4], [8.3, 0.1]], dtype=torch.float64), C6_70,
                       [global_channel(*kc)], rate=0.5)
    if name == "K-D":
        s = RP.duration_mode_samples([0.4, 0.4, 0.2], [2.0, 5.0, 3.0], [0.5, 0.0, 1.0], [0.0, 0.0, 0.0])
        return Problem(two, C6_70, [Channel(s["amp"], s["det"], s["phase"])], rate=0.5)
    if name == "K-E":
        x = torch.arange(300, dtype=torch.float64) / 300
        cust = (6.0 * torch.sin(math.pi * x) * torch.exp(-2.0 * x), RP.constant(300, 1.5), 0.0)
        return Problem(two, C6_70, [global_channel(*kb, cust)], rate=0.5)
    if name == "K-F":
        return Problem(two, C6_70, [global_channel(*kb)], rate=0.5, noise={"dephasing_rate": 2.0})
    if name == "K-G":
        p = [(RP.constant(131, 5.0), RP.constant(131, 5.0), 5.0)] * 8
        return Problem(torch.tensor([[-3.25, 0.0], [3.25, 0.0]], dtype=torch.float64), C6_60,
                       [global_channel(*p)], rate=0.05)
    raise KeyError(name)


def independent_problem(name: str, theta: torch.Tensor) -> Problem:
    """The cases of tests/golden/make_independent_kats.py, parametrised by the four scalars ``theta`` =
    (amplitude scale, detuning scale, phase chirp, x shift of atom 1) so that gradients reach every
    autograd leaf class of the path."""
    if name == "C1":
        base = Problem(torch.tensor([[0.0, 0.0], [8.0, 0.0]], dtype=torch.float64), C6_70,
                       [global_channel((RP.constant(1000, 5.0), RP.constant(1000, 0.0), 0.0))], rate=1.0,
                       evaluation_times="Minimal")
    elif name == "C2-small":
        k = torch.arange(1100, dtype=torch.float64) / 1100
        base = Problem(chain(4, 7.0), C6_60,
                       [Channel(6.0 * torch.sin(math.pi * k) ** 2, -8.0 + 16.0 * k, torch.zeros(1100, dtype=torch.float64))],
                       rate=0.05)
    else:
        base = kat_problem(name)
    ch = base.channels[0]
    T = ch.amp.numel()
    chirp = torch.arange(T, dtype=torch.float64) / T
    shift = torch.zeros_like(base.coords)
    shift[1, 0] = 1.0
    return Problem(base.coords + theta[3] * shift, base.c6,
                   [Channel(theta[0] * ch.amp, theta[1] * ch.det, ch.phase + theta[2] * chirp)],
                   rate=base.rate, noise=base.noise, evaluation_times=base.evaluation_times)


KAT_SOLVER = {"K-A": "dp5_se", "K-B": "krylov_se", "K-C": "krylov_se", "K-D": "krylov_se",
              "K-E": "krylov_se", "K-F": "dp5_me", "K-G": "dp5_se", "C1": "dp5_se", "C2-small": "dp5_se"}


def random_problem(n: int, seed: int = 0, T: int = 300, rate: float = 0.2, local: bool = False,
                   noise: Optional[dict] = None, grad: bool = True) -> Problem:
    g = torch.Generator().manual_seed(seed)
    coords = (torch.rand(n, 2, dtype=torch.float64, generator=g) * 3
              + chain(n, 7.0)).requires_grad_(grad)
    amp = (RP.blackman(T, 2.0) + 1.0 + 0.3 * torch.rand(T, dtype=torch.float64, generator=g)).detach().requires_grad_(grad)
    det = (RP.ramp(T, -3.0, 2.0)).detach().requires_grad_(grad)
    ph = torch.linspace(0, 1.0, T, dtype=torch.float64).requires_grad_(grad)
    chans = [Channel(amp, det, ph)]
    if local:
        chans.append(Channel(RP.constant(T, 1.5).requires_grad_(grad),
                             RP.constant(T, 0.7).requires_grad_(grad),
                             RP.constant(T, 0.3).requires_grad_(grad), "Local", min(1, n - 1)))
    return Problem(coords, C6_60, chans, rate=rate, noise=noise or {})
