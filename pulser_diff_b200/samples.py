"""Pulser-free stand-ins for the sampled-sequence objects the emulator consumes.

The reference takes a ``pulser.sampler.samples.SequenceSamples`` (backend.py:61-115) produced
by Pulser @ fcf9804, which is not installable here.  The hot path only
needs, per channel, the 1-sample-per-ns ``amp`` / ``det`` / ``phase`` arrays, the channel's
addressing and targets -- this module carries exactly that, plus the waveform sample rules of
SURVEY.md Appendix B written with torch so pulse parameters stay differentiable.
A maintainer integrating with real Pulser passes ``SequenceSamples.from_pulser(...)``-style
data instead.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

F64 = torch.float64


def _t(x) -> Tensor:
    return x.to(F64) if isinstance(x, Tensor) else torch.tensor(x, dtype=F64)


# ---- waveform sample rules ---------------------------------------------
def constant_waveform(duration: int, value) -> Tensor:
    return torch.ones(int(duration), dtype=F64) * _t(value).reshape(())


def ramp_waveform(duration: int, start, stop) -> Tensor:
    start, stop = _t(start).reshape(()), _t(stop).reshape(())
    if duration == 1:
        return start.reshape(1)
    return start + (stop - start) * (torch.arange(int(duration), dtype=F64) / (duration - 1))


def blackman_waveform(duration: int, area) -> Tensor:
    w = torch.tensor(np.clip(np.blackman(int(duration)), 0, np.inf), dtype=F64)
    return w * (_t(area).reshape(()) / float(w.sum()) / 1e-3)


def kaiser_waveform(duration: int, area, beta: float = 14.0) -> Tensor:
    w = torch.tensor(np.clip(np.kaiser(int(duration), beta), 0, np.inf), dtype=F64)
    return w * (_t(area).reshape(()) / float(w.sum()) / 1e-3)


def custom_waveform(samples) -> Tensor:
    return _t(samples).reshape(-1)


def tanh_envelope(t_ns: Tensor, ti_us, tf_us, value, edge_steepness: float = 1.0) -> Tensor:
    """Smooth box between ti and tf (us) sampled at integer ns (reference
    waveform_funcs.py:9-27; the ``ti == 0`` branch drops the rising edge)."""
    tf = _t(tf_us) * 1000
    fall = 0.5 * (1.0 + torch.tanh(edge_steepness * (-(t_ns - tf))))
    if isinstance(ti_us, (int, float)) and ti_us == 0:
        return _t(value) * fall
    rise = 0.5 * (1.0 + torch.tanh(edge_steepness * (t_ns - _t(ti_us) * 1000)))
    return _t(value) * (rise + fall - 1.0)


def duration_mode_samples(durations_us: Sequence, amps: Sequence, dets: Sequence,
                          phases: Sequence) -> dict[str, Tensor]:
    """Samples of the 1-ns-pulse sequence ``QuantumModel`` builds when durations are trainable
    (reference model.py:184-206, 301-322, 324-368): ``sum(int(d*1000)) + 5`` samples, each the
    sum of the pulses' tanh envelopes."""
    total = sum(int(float(d) * 1000) for d in durations_us) + 5
    # built where the trainable parameters live: with pulse parameters on the GPU the envelopes (and their
    # autograd graph) never touch the host
    dev = next((x.device for x in (*durations_us, *amps, *dets, *phases) if isinstance(x, Tensor)), None)
    t = torch.arange(total, dtype=F64, device=dev)
    out = {k: torch.zeros(total, dtype=F64, device=dev) for k in ("amp", "det", "phase")}
    ti = 0
    for d, a, de, ph in zip(durations_us, amps, dets, phases):
        tf = ti + _t(d).reshape(()).to(t.device)
        for key, v in (("amp", a), ("det", de), ("phase", ph)):
            out[key] = out[key] + tanh_envelope(t, ti, tf, _t(v).to(t.device))
        ti = tf
    return out


# ---- sampled sequence -----------------------------------------------------------------------------
@dataclass
class Slot:
    ti: int
    tf: int
    targets: set


@dataclass
class ChannelSamples:
    """One channel: (T,) amp / det / phase, 1 sample per ns."""
    amp: Tensor
    det: Tensor
    phase: Tensor
    addressing: str = "Global"                 # "Global" | "Local"
    targets: Optional[Sequence] = None         # qubit ids for Local channels
    slots: list = field(default_factory=list)
    basis: str = "ground-rydberg"

    @property
    def duration(self) -> int:
        return int(self.amp.numel())


class PulseBuilder:
    """Concatenates pulses on one channel, like ``Sequence.add`` on a single channel."""

    def __init__(self, addressing: str = "Global", targets: Optional[Sequence] = None) -> None:
        self.addressing, self.targets = addressing, targets
        self._amp: list[Tensor] = []
        self._det: list[Tensor] = []
        self._phase: list[Tensor] = []
        self._slots: list[Slot] = []

    def add(self, amp: Tensor, det: Tensor, phase=0.0) -> "PulseBuilder":
        if amp.numel() != det.numel():
            raise ValueError("amplitude and detuning waveforms must have the same duration")
        ti = sum(int(a.numel()) for a in self._amp)
        ph = _t(phase)
        self._amp.append(amp.to(F64))
        self._det.append(det.to(F64))
        self._phase.append(ph.reshape(-1) if ph.numel() == amp.numel()
                           else torch.ones(amp.numel(), dtype=F64) * ph.reshape(()))
        self._slots.append(Slot(ti, ti + int(amp.numel()), set(self.targets or [])))
        return self

    def delay(self, duration: int) -> "PulseBuilder":
        z = torch.zeros(int(duration), dtype=F64)
        return self.add(z, z, 0.0)

    def build(self) -> ChannelSamples:
        return ChannelSamples(torch.cat(self._amp), torch.cat(self._det), torch.cat(self._phase),
                              self.addressing, self.targets, list(self._slots))


class SequenceSamples:
    """The slice of ``pulser.sampler.samples.SequenceSamples`` the emulator touches."""

    def __init__(self, channels: Iterable[ChannelSamples]) -> None:
        self.channels = list(channels)
        if not self.channels:
            raise ValueError("SequenceSamples is empty.")
        self.used_bases = {c.basis for c in self.channels}
        self._in_xy = False
        self._measurement = None

    @property
    def max_duration(self) -> int:
        return max(c.duration for c in self.channels)

    def extend_duration(self, new_duration: int) -> "SequenceSamples":
        """Zero-pad amp/det, edge-pad phase (reference backend.py:114-115 via Pulser)."""
        out = []
        for c in self.channels:
            pad = int(new_duration) - c.duration
            if pad < 0:
                raise ValueError("can not shorten samples")
            z = torch.zeros(pad, dtype=F64)
            ph_pad = c.phase[-1:].detach().expand(pad) if c.duration else z
            out.append(ChannelSamples(torch.cat([c.amp, z]), torch.cat([c.det, z]),
                                      torch.cat([c.phase, ph_pad]), c.addressing, c.targets,
                                      c.slots, c.basis))
        return SequenceSamples(out)

    def to_nested_dict(self, qubit_ids: Sequence, all_local: bool = False) -> dict:
        """{"Global": {basis: {amp,det,phase}}, "Local": {basis: {qid: {...}}}} as Pulser's
        ``to_nested_dict(all_local, samples_type="tensor")`` gives the reference
        (hamiltonian.py:177).  Global channels stay under "Global" (summed per basis) unless
        ``all_local``; Local channels are accumulated per target qubit inside each pulse slot
        ``[ti:tf]``.  The two groups are SEPARATE Hamiltonian terms (hamiltonian.py:478-481): a
        qubit driven by both sees ``0.5 a_g e^{-i ph_g} + 0.5 a_l e^{-i ph_l}``."""
        d = self.max_duration
        res: dict = {"Global": {}, "Local": {}}

        def zero():
            return {k: torch.zeros(d, dtype=F64) for k in ("amp", "det", "phase")}

        def padded(x: Tensor) -> Tensor:
            return x if x.numel() == d else torch.cat([x, torch.zeros(d - x.numel(), dtype=F64)])

        for c in self.channels:
            if c.addressing == "Global" and not all_local:
                dst = res["Global"].setdefault(c.basis, zero())
                for k in ("amp", "det", "phase"):
                    dst[k] = dst[k] + padded(getattr(c, k))
                continue
            everyone = list(qubit_ids) if c.addressing == "Global" else list(c.targets or [])
            per = res["Local"].setdefault(c.basis, {})
            for slot in (c.slots or [Slot(0, c.duration, set(everyone))]):
                window = torch.zeros(d, dtype=F64)
                window[slot.ti:slot.tf] = 1.0
                for q in (slot.targets or everyone):
                    dst = per.setdefault(q, zero())
                    for k in ("amp", "det", "phase"):
                        dst[k] = dst[k] + padded(getattr(c, k)) * window
        return res
