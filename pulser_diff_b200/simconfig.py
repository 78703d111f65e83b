"""Noise configuration carrying the rates that reach the hot path.

Stand-in for reference ``pulser_diff/simconfig.py:15-132`` (a ``pulser_simulation.SimConfig``
subclass; Pulser is not installable here).  Only the Lindblad-type noises and the
deterministic laser-waist amplitude scaling are on the B200 path;
stochastic noises (doppler, SPAM, amplitude with amp_sigma > 0) are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence

LINDBLAD_NOISES = ("dephasing", "relaxation", "depolarizing", "eff_noise")
SUPPORTED_NOISES = {"ising": set(LINDBLAD_NOISES) | {"amplitude"}}


@dataclass(frozen=True)
class SimConfig:
    noise: Any = ()
    dephasing_rate: Any = 0.05
    relaxation_rate: Any = 0.01
    depolarizing_rate: Any = 0.05
    eff_noise_rates: Sequence = field(default_factory=list)
    eff_noise_opers: Sequence = field(default_factory=list)
    laser_waist: Any = None
    amp_sigma: float = 0.0
    runs: int = 15
    samples_per_run: int = 5

    def __post_init__(self) -> None:
        noise = (self.noise,) if isinstance(self.noise, str) else tuple(self.noise)
        object.__setattr__(self, "noise", noise)
        bad = set(noise) - SUPPORTED_NOISES["ising"]
        if bad:
            raise NotImplementedError(
                "Interaction mode 'ising' does not support simulation of noise types on the "
                f"B200 path: {', '.join(sorted(bad))}.")
        if len(self.eff_noise_rates) != len(self.eff_noise_opers):
            raise ValueError("eff_noise_rates and eff_noise_opers must have the same length")
        if "amplitude" in noise and self.amp_sigma != 0.0:
            raise NotImplementedError("stochastic amplitude noise (amp_sigma > 0) is out of scope")

    @property
    def noise_types(self) -> tuple:
        return self.noise

    @property
    def supported_noises(self) -> dict:
        return SUPPORTED_NOISES

    def to_noise_model(self) -> "SimConfig":
        return self

    @property
    def has_lindblad(self) -> bool:
        return any(n in self.noise for n in LINDBLAD_NOISES)
