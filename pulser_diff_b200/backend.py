"""``TorchEmulator`` for the B200 path: the reference's entry point, Pulser-free inputs.

Same public surface as reference ``pulser_diff/backend.py:35-611`` for everything that touches
the hot path -- ``run(time_grad, dist_grad, solver, **options)``, ``set_initial_state``,
``set_evaluation_times``, ``set_config``, ``evaluation_times``, ``sampling_times``,
``qq_distances``, ``get_hamiltonian``, ``build_operator`` -- with ``_run_solver`` re-routed to
:mod:`pulser_diff_b200.solvers`.  Inputs are :class:`~pulser_diff_b200.samples.SequenceSamples`
plus a ``{qubit id: coordinates}`` register and a device object exposing
``interaction_coeff`` (Pulser is not installable here; INTEGRATION.md shows the two-line change
that makes the reference's own emulator call this path).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Union

import torch
from torch import Tensor

from .hamiltonian import Hamiltonian
from .samples import SequenceSamples
from .simconfig import SimConfig
from .simresults import CoherentResults, NoisyResults
from .solvers import SolverType, mesolve, sesolve

C128 = torch.complex128


@dataclass
class DeviceSpec:
    """The one device property the hot path reads (reference hamiltonian.py:343)."""
    interaction_coeff: float
    name: str = "device"


MockDevice = DeviceSpec(5420158.53, "MockDevice")            # rydberg_level 70
Level60Device = DeviceSpec(865723.02, "VirtualDevice(rydberg_level=60)")


class TorchEmulator:
    def __init__(self, sampled_seq: SequenceSamples, register: Union[dict, Tensor], device: Any,
                 sampling_rate: float = 1.0, config: Optional[SimConfig] = None,
                 evaluation_times: Union[float, str, Any] = "Full",
                 torch_device: Union[str, torch.device] = "cuda") -> None:
        if not isinstance(sampled_seq, SequenceSamples):
            raise TypeError("The provided sequence has to be a valid SequenceSamples instance.")
        if sampled_seq.max_duration == 0:
            raise ValueError("SequenceSamples is empty.")
        if isinstance(register, Tensor):
            register = {f"q{i}": register[i] for i in range(register.shape[0])}
        self._register = dict(register)
        for ch in sampled_seq.channels:
            if ch.addressing == "Local" and not set(ch.targets or []) <= set(self._register):
                raise ValueError(
                    "The ids of qubits targeted in Local channels should be defined in register.")
        self._tot_duration = sampled_seq.max_duration
        self.samples_obj = sampled_seq.extend_duration(self._tot_duration + 1)
        if not (0 < sampling_rate <= 1.0):
            raise ValueError(
                f"The sampling rate (`sampling_rate` = {sampling_rate}) must be greater than 0 "
                "and less than or equal to 1.")
        if int(self._tot_duration * sampling_rate) < 4:
            raise ValueError("`sampling_rate` is too small, less than 4 data points.")
        self._hamiltonian = Hamiltonian(self.samples_obj, self._register, device, sampling_rate,
                                        config or SimConfig(), torch_device)
        self._eval_times_array: Tensor
        self.set_evaluation_times(evaluation_times)
        self._meas_basis = self._hamiltonian.basis_name
        self.set_initial_state("all-ground")
        self.dist_dict: dict[str, Tensor] = {}

    # ---- properties (reference backend.py:153-181, 248-251, 282-289) ------------------------
    @property
    def sampling_times(self) -> Tensor:
        return self._hamiltonian.sampling_times

    @property
    def _sampling_rate(self) -> float:
        return self._hamiltonian._sampling_rate

    @property
    def dim(self) -> int:
        return self._hamiltonian.dim

    @property
    def basis_name(self) -> str:
        return self._hamiltonian.basis_name

    @property
    def basis(self) -> dict:
        return self._hamiltonian.basis

    @property
    def config(self) -> SimConfig:
        return self._hamiltonian.config

    @property
    def initial_state(self) -> Tensor:
        return self._initial_state

    @property
    def evaluation_times(self) -> Tensor:
        return self._eval_times_array

    @property
    def qq_distances(self) -> dict[str, Tensor]:
        return self.dist_dict

    def set_config(self, cfg: SimConfig) -> None:
        if not isinstance(cfg, SimConfig):
            raise ValueError(f"Object {cfg} is not a valid `SimConfig`.")
        self._hamiltonian.set_config(cfg)

    def reset_config(self) -> None:
        self._hamiltonian.set_config(SimConfig())

    def set_initial_state(self, state: Union[str, Tensor]) -> None:
        """"all-ground" or a (2^N, B) tensor (reference backend.py:253-280)."""
        n = self._hamiltonian._size
        if isinstance(state, str) and state == "all-ground":
            psi = torch.zeros(2 ** n, 1, dtype=C128)
            psi[-1] = 1.0                       # kron(|g>, ..., |g>) with |g> = e1
            if self._hamiltonian.torch_device.type == "cuda" and torch.cuda.is_available():
                psi = psi.pin_memory()          # staged for the host->device copy of every run
            self._initial_state = psi
        else:
            legal = self._hamiltonian.dim ** n
            if state.shape[0] != legal:
                raise ValueError("Incompatible shape of initial state." +
                                 f"Expected {legal}, got {state.shape[0]}.")
            self._initial_state = state.to(C128)

    def set_evaluation_times(self, value: Union[str, Any, float]) -> None:
        """"Full" | "Minimal" | list of times (us) | float fraction (reference
        backend.py:312-375); 0 and the final time are always included, result sorted/unique."""
        st = self._hamiltonian.sampling_times
        if isinstance(value, str):
            if value == "Full":
                eval_times = torch.clone(st)
            elif value == "Minimal":
                eval_times = torch.tensor([], dtype=torch.float64)
            else:
                raise ValueError("Wrong evaluation time label. It should be `Full`, `Minimal`, "
                                 "an array of times or a float between 0 and 1.")
        elif isinstance(value, float):
            if value > 1 or value <= 0:
                raise ValueError("evaluation_times float must be between 0 and 1.")
            indices = torch.linspace(0, len(st) - 1, int(value * len(st)), dtype=torch.int)
            eval_times = st[indices]
        elif isinstance(value, (list, tuple, Tensor)):
            v = torch.as_tensor(value, dtype=torch.float64)
            if torch.max(v) > self._tot_duration / 1000:
                raise ValueError("Provided evaluation-time list extends further than sequence duration.")
            if torch.min(v) < 0:
                raise ValueError("Provided evaluation-time list contains negative values.")
            eval_times = v
        else:
            raise ValueError("Wrong evaluation time label. It should be `Full`, `Minimal`, "
                             "an array of times or a float between 0 and 1.")
        self._eval_times_array = torch.cat(
            [eval_times, torch.tensor([0.0, self._tot_duration / 1000], dtype=eval_times.dtype)]
        ).unique().requires_grad_(False)
        self._eval_times_instruction = value

    def build_operator(self, operations) -> Tensor:
        return self._hamiltonian.build_operator(operations)

    def get_hamiltonian(self, time: float) -> Tensor:
        """Sparse H at ``time`` (ns), small registers only (reference backend.py:401-427)."""
        if time > self._tot_duration:
            raise ValueError(f"Provided time (`time` = {time}) must be less than or equal to the "
                             f"sequence duration ({self._tot_duration}).")
        if time < 0:
            raise ValueError(f"Provided time (`time` = {time}) must be greater than or equal to 0.")
        return self._hamiltonian._hamiltonian(time / 1000)

    def run(self, time_grad: bool = False, dist_grad: bool = False,
            solver: SolverType = SolverType.DP5_SE, **options: Any) -> CoherentResults:
        """Evolve the register on the B200 and return the states at ``evaluation_times``."""
        if time_grad:
            self._eval_times_array.requires_grad_(True)
        if dist_grad:
            for k, v in self._hamiltonian._dist_dict.items():
                if v.is_leaf:
                    v.requires_grad_(True)
                else:
                    v.retain_grad()
                self.dist_dict[k] = v
            self._hamiltonian.refresh_couplings()
        if self.config.has_lindblad:
            solver = SolverType.DP5_ME
        if self.config.needs_resampling:
            return self._run_noisy(solver, options)
        ham = self._hamiltonian._hamiltonian
        if solver in (SolverType.DP5_SE, SolverType.KRYLOV_SE):
            result = sesolve(H=ham, psi0=self.initial_state, tsave=self._eval_times_array,
                             solver=solver, options=options)
        elif solver == SolverType.DP5_ME:
            psi = self.initial_state
            result = mesolve(H=ham, rho0=torch.matmul(psi, psi.mH).unsqueeze(-1),
                             L=self._hamiltonian._collapse_ops, tsave=self._eval_times_array,
                             solver=solver, options=options)
        else:
            raise ValueError(f"Solver {solver} not available.")
        self._last_result = result
        meas = ({"epsilon": self.config.epsilon, "epsilon_prime": self.config.epsilon_prime}
                if "SPAM" in self.config.noise else None)
        return CoherentResults(result.states, self._hamiltonian._size, self._hamiltonian.basis_name,
                               self._eval_times_array, self._meas_basis, meas)

    def _run_noisy(self, solver: SolverType, options: dict,
                   generator: Optional[torch.Generator] = None) -> NoisyResults:
        """Averages over the random Hamiltonians of the stochastic noises (reference backend.py:568-611).

        Where the reference loops ``runs`` times over ``_construct_hamiltonian`` + solver, the draws are
        made up front and all runs that share a set of badly prepared atoms go to the device as ONE
        batch of parameter sets (``ops.evolve_units``: a single launch for registers of <= 14 atoms);
        Lindblad noise on top runs them one after the other through ``mesolve``.  Bitstring counts are
        sampled on the device per run and evaluation time, ``samples_per_run`` shots each."""
        from collections import Counter
        from . import _cabi, ops
        from .hamiltonian import StructuredHamiltonian
        cfg, ham = self.config, self._hamiltonian
        n, dev = ham._size, ham.torch_device
        noise = set(cfg.noise)
        generator = generator if generator is not None else getattr(self, "noise_generator", None)
        resample = "doppler" in noise or ("amplitude" in noise and cfg.amp_sigma != 0.0)
        if resample:
            draws = [(ham.noise_realisation(generator), 1) for _ in range(cfg.runs)]
        else:       # SPAM only: distinct preparation patterns with their multiplicities, nothing else random
            pats = Counter(tuple((torch.rand(n, dtype=torch.float64, generator=generator) < cfg.eta).tolist())
                           for _ in range(cfg.runs))
            draws = [(ham.noise_realisation(generator, bad_atoms=torch.tensor(p), resample=False), reps)
                     for p, reps in pats.most_common()]
        self._last_noise_draws, self._last_noisy_states = draws, []
        meas = ({"epsilon": cfg.epsilon, "epsilon_prime": cfg.epsilon_prime} if "SPAM" in noise else None)
        times = self._eval_times_array
        total = [Counter() for _ in range(len(times))]
        groups: dict = {}
        for d, reps in draws:
            groups.setdefault(tuple(d["bad_atoms"].tolist()), []).append((d, reps))
        psi0 = self.initial_state.to(device=dev, dtype=_cabi.Options.from_dict(options).state_dtype).transpose(0, 1).contiguous()
        for members in groups.values():
            d0 = members[0][0]
            if solver == SolverType.DP5_SE:
                dv = torch.stack([d["det_values"] for d, _ in members])
                av = torch.stack([d["amp_values"] for d, _ in members])
                st = ops.evolve_units(psi0.repeat(len(members), 1, 1), times.detach(), dv, av, d0["pair_u"],
                                      n_qubits=n, dt=d0["dt"], det_masks=d0["det_masks"], amp_masks=d0["amp_masks"],
                                      options=_cabi.Options.from_dict(options))
                per_run = [st[u].permute(0, 2, 1) for u in range(len(members))]
            else:
                per_run = []
                for d, _ in members:
                    H = StructuredHamiltonian(n, d["pair_u"], d["dt"], d["n_samples"],
                                              list(zip(d["det_masks"], d["det_values"])),
                                              list(zip(d["amp_masks"], d["amp_values"])), dev)
                    if solver == SolverType.DP5_ME:
                        psi = self.initial_state
                        per_run.append(mesolve(H, torch.matmul(psi, psi.mH).unsqueeze(-1), ham._collapse_ops,
                                               times, solver, options).states)
                    else:
                        per_run.append(sesolve(H, self.initial_state, times, solver, options).states)
            self._last_noisy_states = getattr(self, "_last_noisy_states", [])
            self._last_noisy_states += [(d, st_) for st_, (d, _) in zip(per_run, members)]
            for states, (_, reps) in zip(per_run, members):
                res = CoherentResults(states, n, ham.basis_name, times, self._meas_basis, meas)
                for k, t in enumerate(times.tolist()):
                    total[k] += res.sample_state(t, n_samples=cfg.samples_per_run * reps)
        return NoisyResults(total, n, ham.basis_name, times, cfg.runs * cfg.samples_per_run)
