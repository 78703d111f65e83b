"""Pulser-free restatement of the pulse sampling that feeds the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pulser itself (pinned at
fcf980463f47b92722901aba0e63bec9a28e01af, reference pyproject.toml:31-32) is
not installable here, so the few waveform rules the notebooks use are restated
from SURVEY.md Appendix B [UPSTREAM-RECALLED, KAT-CONFIRMED]:

* one sample per ns;
* ``ConstantWaveform(D, v)``  -> ``full(D, v)``
* ``RampWaveform(D, a, b)``   -> ``linspace(a, b, D)``
* ``BlackmanWaveform(D, A)``  -> ``clip(blackman(D), 0) * A / sum / 1e-3``
* ``KaiserWaveform(D, A, beta)`` -> same with ``kaiser(D, beta)``
* ``CustomWaveform(samples)`` -> as given
* ``extend_duration(T + 1)`` appends one sample with amp = det = 0
  (reference backend.py:114-115).
* duration mode (reference model.py:184-206, 324-368 and
  waveform_funcs.py:9-27): total samples ``sum(int(dur_us*1000)) + 5``, sample
  ``t`` carries the sum of tanh box envelopes.

Everything is torch float64 and differentiable w.r.t. tensor arguments so the
tape oracle can produce parameter gradients.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

F64 = torch.float64


def _t(x) -> Tensor:
    return x.to(F64) if isinstance(x, Tensor) else torch.tensor(x, dtype=F64)


def constant(duration: int, value) -> Tensor:
    return torch.ones(duration, dtype=F64) * _t(value).reshape(())


def ramp(duration: int, start, stop) -> Tensor:
    start, stop = _t(start).reshape(()), _t(stop).reshape(())
    if duration == 1:
        return start.reshape(1)
    k = torch.arange(duration, dtype=F64) / (duration - 1)
    return start + (stop - start) * k


def blackman(duration: int, area) -> Tensor:
    w = torch.tensor(np.clip(np.blackman(duration), 0, np.inf), dtype=F64)
    return w * (_t(area).reshape(()) / float(w.sum()) / 1e-3)


def kaiser(duration: int, area, beta: float = 14.0) -> Tensor:
    w = torch.tensor(np.clip(np.kaiser(duration, beta), 0, np.inf), dtype=F64)
    return w * (_t(area).reshape(()) / float(w.sum()) / 1e-3)


def custom(samples) -> Tensor:
    return _t(samples).reshape(-1)


def tanh_box_first(t_ns: Tensor, tf_us, value, steep: float = 1.0) -> Tensor:
    """reference waveform_funcs.py:16-17 (ti == 0 branch)."""
    return _t(value) * 0.5 * (1.0 + torch.tanh(steep * (-(t_ns - _t(tf_us) * 1000))))


def tanh_box(t_ns: Tensor, ti_us, tf_us, value, steep: float = 1.0) -> Tensor:
    """reference waveform_funcs.py:19-24 (ti != 0 branch)."""
    return _t(value) * (
        0.5 * (1.0 + torch.tanh(steep * (t_ns - _t(ti_us) * 1000)))
        + 0.5 * (1.0 + torch.tanh(steep * (-(t_ns - _t(tf_us) * 1000))))
        - 1.0
    )


def duration_mode_samples(durations_us, amps, dets, phases) -> dict[str, Tensor]:
    """1-ns constant pulses carrying summed tanh envelopes.

    Follows reference model.py:184-206 (one ConstantPulse(1, ...) per ns),
    :301-322 (total duration = sum(int(d*1000)) + 5) and :324-368 (envelopes,
    cumulative ti/tf in microseconds).
    """
    total = sum(int(float(d) * 1000) for d in durations_us) + 5
    t = torch.arange(total, dtype=F64)
    out = {"amp": torch.zeros(total, dtype=F64), "det": torch.zeros(total, dtype=F64),
           "phase": torch.zeros(total, dtype=F64)}
    ti = None
    for d, a, de, ph in zip(durations_us, amps, dets, phases):
        if ti is None:
            tf = _t(d).reshape(())
            for key, v in (("amp", a), ("det", de), ("phase", ph)):
                out[key] = out[key] + tanh_box_first(t, tf, v)
        else:
            tf = ti + _t(d).reshape(())
            for key, v in (("amp", a), ("det", de), ("phase", ph)):
                out[key] = out[key] + tanh_box(t, ti, tf, v)
        ti = tf
    return out


class GlobalChannelSamples:
    """Samples of one global Rydberg channel: concatenated pulses, 1 sample/ns.

    ``add(amp, det, phase)`` appends a pulse (amp/det arrays of equal length,
    scalar phase).  ``extended()`` returns the (T+1)-long arrays the reference
    emulator hands to ``Hamiltonian`` (backend.py:114-115).
    """

    def __init__(self) -> None:
        self.amp: list[Tensor] = []
        self.det: list[Tensor] = []
        self.phase: list[Tensor] = []

    def add(self, amp: Tensor, det: Tensor, phase=0.0) -> "GlobalChannelSamples":
        assert amp.numel() == det.numel()
        self.amp.append(amp)
        self.det.append(det)
        ph = _t(phase)
        self.phase.append(ph.reshape(-1) if ph.numel() == amp.numel()
                          else torch.ones(amp.numel(), dtype=F64) * ph.reshape(()))
        return self

    @property
    def duration(self) -> int:
        return int(sum(a.numel() for a in self.amp))

    def extended(self) -> dict[str, Tensor]:
        z = torch.zeros(1, dtype=F64)
        phase = torch.cat(self.phase)
        return {
            "amp": torch.cat(self.amp + [z]),
            "det": torch.cat(self.det + [z]),
            "phase": torch.cat([phase, phase[-1:].detach()]),
        }
