"""Noise configuration carrying the rates that reach the hot path.

Stand-in for reference ``pulser_diff/simconfig.py:15-132`` (a ``pulser_simulation.SimConfig``
subclass; Pulser is not installable here).  Lindblad-type noises act inside one evolution; the
stochastic noises (doppler detuning offsets, shot-to-shot amplitude fluctuations with the Gaussian beam
profile, state-preparation and detection errors) resample the Hamiltonian ``runs`` times (reference
backend.py:568-611, hamiltonian.py:179-219, 270-286) -- on the B200 path those runs are ONE batch of
parameter sets (``ops.evolve_units``).  Field names and defaults follow Pulser's ``SimConfig``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence

LINDBLAD_NOISES = ("dephasing", "relaxation", "depolarizing", "eff_noise")
STOCHASTIC_NOISES = ("doppler", "amplitude", "SPAM")
SUPPORTED_NOISES = {"ising": set(LINDBLAD_NOISES) | set(STOCHASTIC_NOISES)}
KB, KEFF, MASS = 1.38e-23, 8.7, 1.45e-25     # Boltzmann constant, effective wave vector (1/um), Rb-87 mass


def doppler_sigma(temperature_k: float) -> float:
    """Standard deviation (rad/us) of the Doppler detuning at a temperature in kelvin."""
    return KEFF * (KB * temperature_k / MASS) ** 0.5


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
    temperature: float = 50.0          # uK
    eta: float = 0.005                 # state-preparation error probability per atom
    epsilon: float = 0.01              # detection false-positive probability
    epsilon_prime: float = 0.05        # detection false-negative probability

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
        if self.runs < 1 or self.samples_per_run < 1:
            raise ValueError("runs and samples_per_run must be positive")

    @property
    def noise_types(self) -> tuple:
        return self.noise

    @property
    def supported_noises(self) -> dict:
        return SUPPORTED_NOISES

    def to_noise_model(self) -> "SimConfig":
        return self

    @property
    def state_prep_error(self) -> float:
        return self.eta

    @property
    def needs_resampling(self) -> bool:
        """True when a run has to be repeated over random Hamiltonians (reference backend.py:532-566)."""
        n = set(self.noise)
        return ("doppler" in n or ("amplitude" in n and self.amp_sigma != 0.0)
                or ("SPAM" in n and self.eta > 0))

    @property
    def has_lindblad(self) -> bool:
        return any(n in self.noise for n in LINDBLAD_NOISES)
