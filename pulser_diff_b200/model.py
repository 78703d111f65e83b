"""``QuantumModel``: the optimisation entry point on top of the B200 emulator.

Same role and method names as reference ``pulser_diff/model.py:30-431`` -- an ``nn.Module`` whose
parameters are pulse parameters (and optionally qubit coordinates), ``forward()`` ->
``(times, states)``, ``expectation(obs)`` -> ``(times, values)``, ``check_constraints()``,
``update_sequence()`` -- with one difference forced by the missing Pulser dependency: instead of
a parametrised ``pulser.Sequence`` the model takes ``sample_fn(params) -> SequenceSamples``, a
plain function from the current parameter tensors to the 1-sample-per-ns channel arrays (the
helpers in :mod:`pulser_diff_b200.samples` cover the waveforms the notebooks use).  As in the
reference, every forward builds a fresh emulator (model.py:405-414); rebuilding the sequence
is host-side work and out of scope to accelerate.
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Union

import torch
from torch import Tensor
from torch.nn import Module, Parameter, ParameterDict

from .backend import TorchEmulator
from .samples import SequenceSamples
from .simconfig import SimConfig
from .solvers import SolverType


class QuantumModel(Module):
    def __init__(self, register: dict, device: Any, sample_fn: Callable[[dict], SequenceSamples],
                 trainable_param_values: dict[str, Tensor], constraints: Optional[dict] = None,
                 sampling_rate: float = 1.0, solver: SolverType = SolverType.DP5_SE,
                 initial_state: Optional[Tensor] = None, noise_config: Optional[SimConfig] = None,
                 time_grad: bool = False, dist_grad: bool = False,
                 torch_device: Union[str, torch.device] = "cuda", **options: Any) -> None:
        super().__init__()
        self.device_spec, self.sample_fn = device, sample_fn
        self.constraints = constraints or {}
        self.sampling_rate, self.solver = sampling_rate, solver
        self.initial_state, self.noise_config = initial_state, noise_config
        self.time_grad, self.dist_grad, self.options = time_grad, dist_grad, options
        self.torch_device = torch.device(torch_device)
        reg_names = set(register)
        self.seq_param_values = ParameterDict({k: Parameter(v.detach().clone().to(torch.float64))
                                               for k, v in trainable_param_values.items()
                                               if k not in reg_names})
        self.reg_param_values = ParameterDict({str(k): Parameter(v.detach().clone().to(torch.float64))
                                               for k, v in trainable_param_values.items()
                                               if k in reg_names})
        self._fixed_register = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in register.items()
                                if str(k) not in self.reg_param_values}
        self._reg_order = list(register)
        self.update_sequence()

    # -- reference model.py:370-403 ------------------------------------------------------------
    def check_constraints(self) -> None:
        with torch.no_grad():
            for name, lim in self.constraints.items():
                for group in (self.seq_param_values, self.reg_param_values):
                    if name in group:
                        group[name].clamp_(min=lim.get("min"), max=lim.get("max"))

    def update_sequence(self) -> None:
        """Re-sample the pulses and rebuild the register from the current parameter values."""
        self.register = {k: (self.reg_param_values[str(k)] if str(k) in self.reg_param_values
                             else self._fixed_register[k]) for k in self._reg_order}
        self.built_samples = self.sample_fn(dict(self.seq_param_values.items()))
        if not isinstance(self.built_samples, SequenceSamples):
            self.built_samples = SequenceSamples(self.built_samples)

    # -- reference model.py:405-431 ------------------------------------------------------------
    def _run(self):
        sim = TorchEmulator(self.built_samples, self.register, self.device_spec, self.sampling_rate,
                            torch_device=self.torch_device)
        if self.initial_state is not None:
            sim.set_initial_state(self.initial_state)
        if self.noise_config is not None:
            sim.set_config(self.noise_config)
        results = sim.run(time_grad=self.time_grad, dist_grad=self.dist_grad, solver=self.solver,
                          **self.options)
        return sim.evaluation_times, results

    def forward(self) -> tuple[Tensor, Tensor]:
        times, results = self._run()
        return times, results.states

    def expectation(self, obs: Tensor) -> tuple[Tensor, Tensor]:
        times, results = self._run()
        return times, results.expect([obs])[0]
