"""torch-facing operators of the B200 evolution engine.

* :func:`evolve` - differentiable evolution (``torch.autograd.Function``) whose leaves are
  exactly the tensors the reference's tape differentiates: the coefficient arrays captured by
  ``H`` (``det_values`` / ``amp_values``, reference hamiltonian.py:506-520), the pair couplings
  behind ``int_mat`` (hamiltonian.py:341-343), ``tsave`` (backend.py:453-455) and the initial
  state.  Backward = adjoint sweep on the device; ``retain_graph`` / repeated
  ``autograd.grad`` calls work because the step log is kept, not consumed
  (reference derivative.py:40,76 call it 40x in docs/basic_usage.ipynb cell 26).
* :func:`evolve_units` - the same for a batch of pulse-parameter sets of one register.
* ``torch.ops.pulser_diff_b200.{hpsi, rhs, evolve_states, expect_diag}`` - ``torch.library``
  custom ops over the same C ABI.  ``hpsi`` and ``rhs`` (one generator application) are
  differentiable w.r.t. the state, the coefficient arrays and the pair couplings (C ABI
  ``pd_rhs_vjp``), so a stepper written in torch on top of them back-propagates like the
  reference's ``H_t(t) @ psi`` (hamiltonian.py:526-546).

Internal state layout is batch-major ``(batch, dim)``; the reference layout ``(dim, batch)``
is converted in :mod:`pulser_diff_b200.solvers`.

Precision.  State tensors are complex128 (what the reference computes in, backend.py:271, 280) or -- north_star's
optional 1e-5 tier -- complex64: every operator here takes the precision from the dtype of the state it is given.
complex64 states of the bandwidth-bound shapes (kets of N >= 15, every density matrix) run in the complex64
build of the library (half the bytes per vector pass); kets of N <= 14 live in registers / L2 where nothing is
bandwidth-bound, so they are computed by the complex128 kernels and only cast at the boundary
(:func:`compute_dtype`).  Coefficients, times and every gradient w.r.t. them stay float64 / complex128.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _cabi
from ._cabi import PD_DENSITY, PD_KET, Options, Plan

_program_ids = itertools.count()
_PLAN_CACHE: dict[tuple, Plan] = {}
_PLAN_CACHE_MAX_PLANS = 16                 # small registers: cheap to keep
_PLAN_CACHE_MAX_AMPLITUDES = 1 << 27       # plans kept beside a new one hold at most this many amplitudes in total
_UNIT_PROGRAMS: dict[tuple, "Program"] = {}
_PROGRAMS: dict[tuple, "Program"] = {}     # content -> Program: unchanged coefficient tables keep their identity
_PROGRAMS_MAX = 32


@dataclass
class Program:
    """Everything that defines H(t) (and the dissipator) for one evolution, detached."""
    n_qubits: int
    kind: int
    dt: float
    det_masks: List[int]
    det_values: Tensor            # (n_det, n_samples) float64, cpu
    amp_masks: List[int]
    amp_values: Tensor            # (n_amp, n_samples) complex128, cpu
    pair_u: Tensor                # (N, N) float64, cpu
    collapse: Optional[Tensor] = None   # (n_ops, 2, 2) complex128, cpu
    id: int = field(default_factory=lambda: next(_program_ids))


def clear_plan_cache() -> None:
    """Drop cached plans (frees their device workspace)."""
    _PLAN_CACHE.clear()
    _UNIT_PROGRAMS.clear()
    _PROGRAMS.clear()


C64_MIN_KET_QUBITS = 15     # kets below this are served by the complex128 small-register kernels


def compute_dtype(state_dtype: torch.dtype, n_qubits: int, kind: int) -> torch.dtype:
    """Precision the library computes in for states of ``state_dtype`` (see the module docstring)."""
    if state_dtype == torch.complex64 and (kind == PD_DENSITY or n_qubits >= C64_MIN_KET_QUBITS):
        return torch.complex64
    if state_dtype in (torch.complex64, torch.complex128):
        return torch.complex128
    raise TypeError(f"state vectors must be complex128 or complex64, got {state_dtype}")


def get_plan(n_qubits: int, batch: int, kind: int, device: torch.device,
             dtype: torch.dtype = torch.complex128) -> Plan:
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    c64 = dtype == torch.complex64
    key = (n_qubits, batch, kind, str(device), c64,
           _cabi._lib_paths[c64] or (_cabi.DEFAULT_LIBRARY_C64 if c64 else _cabi.DEFAULT_LIBRARY))
    plan = _PLAN_CACHE.pop(key, None)
    if plan is None:
        # bounded, least recently used first out: a plan pins its whole workspace (tens of GiB at N >= 26),
        # so sweeping N or the batch size must not pile plans up until the device is full
        budget = _PLAN_CACHE_MAX_AMPLITUDES
        kept = 0
        for k in reversed(list(_PLAN_CACHE)):
            kept += _PLAN_CACHE[k].dim * _PLAN_CACHE[k].batch
            if kept > budget or len(_PLAN_CACHE) >= _PLAN_CACHE_MAX_PLANS:
                del _PLAN_CACHE[k]
        plan = Plan(n_qubits, batch, kind, device, dtype)
    _PLAN_CACHE[key] = plan          # most recently used last
    return plan


def configure(plan: Plan, prog: Program) -> None:
    if plan.program_id == prog.id:
        return
    plan.set_interaction(prog.pair_u)
    plan.set_terms(prog.dt, prog.det_masks, prog.det_values, prog.amp_masks, prog.amp_values)
    if plan.kind == PD_DENSITY:
        plan.set_collapse(prog.collapse)
    plan.program_id = prog.id


def make_program(n_qubits: int, kind: int, dt: float, det_masks: Sequence[int], det_values: Tensor,
                 amp_masks: Sequence[int], amp_values: Tensor, pair_u: Tensor,
                 collapse: Optional[Tensor]) -> Program:
    # a program may have no detuning (or no drive) term at all: keep the sample axis explicit
    n_s = int(det_values.shape[-1]) if len(det_masks) else (int(amp_values.shape[-1]) if len(amp_masks) else 0)
    dm, am = [int(m) for m in det_masks], [int(m) for m in amp_masks]
    dv = det_values.detach().to("cpu", torch.float64).reshape(len(dm), n_s).contiguous()
    av = amp_values.detach().to("cpu", torch.complex128).reshape(len(am), n_s).contiguous()
    pu = pair_u.detach().to("cpu", torch.float64).contiguous()
    co = None if collapse is None else collapse.detach().to("cpu", torch.complex128).contiguous()
    # Programs are identified by CONTENT (the tables are a few KiB): a stepper written on the hpsi / rhs ops calls
    # this once per generator application with the same tables, and a plan that sees the same Program id skips
    # its reconfiguration (diagonal rebuild + stream synchronisation + table copies, see configure)
    key = (n_qubits, kind, float(dt), tuple(dm), tuple(am), n_s, dv.numpy().tobytes(),
           torch.view_as_real(av).numpy().tobytes(), pu.numpy().tobytes(),
           None if co is None else torch.view_as_real(co).numpy().tobytes())
    prog = _PROGRAMS.pop(key, None)
    if prog is None:
        if len(_PROGRAMS) >= _PROGRAMS_MAX:
            del _PROGRAMS[next(iter(_PROGRAMS))]        # least recently used first
        prog = Program(n_qubits, kind, float(dt), dm, dv, am, av, pu, co)
    _PROGRAMS[key] = prog
    return prog


class _EvolveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state0: Tensor, tsave: Tensor, det_values: Tensor, amp_values: Tensor,
                pair_u: Tensor, n_qubits: int, kind: int, dt: float, det_masks, amp_masks,
                collapse, solver: int, opt: Options):
        batch = int(state0.shape[0])
        cd = compute_dtype(state0.dtype, n_qubits, kind)
        plan = get_plan(n_qubits, batch, kind, state0.device, cd)
        prog = make_program(n_qubits, kind, dt, det_masks, det_values, amp_masks, amp_values,
                            pair_u, collapse)
        configure(plan, prog)
        need = any(ctx.needs_input_grad[:5])
        states, tape = plan.evolve_forward(solver, opt, state0.detach().to(cd), tsave, want_tape=need)
        ctx.plan, ctx.prog, ctx.tape = plan, prog, tape
        ctx.state_dtype = state0.dtype
        ctx.meta = (tsave.device, tsave.dtype, det_values.device, det_values.dtype,
                    amp_values.device, amp_values.dtype, pair_u.device, pair_u.dtype,
                    tuple(det_values.shape), tuple(amp_values.shape))
        ctx.save_for_backward(states)
        ctx.set_materialize_grads(False)
        return states if cd == state0.dtype else states.to(state0.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_states: Optional[Tensor]):
        none = (None,) * 13
        if grad_states is None or ctx.tape is None:
            return none
        (states,) = ctx.saved_tensors
        plan, prog = ctx.plan, ctx.prog
        configure(plan, prog)
        w_s0, w_ts, w_det, w_amp, w_pair = ctx.needs_input_grad[:5]
        g_det, g_amp, g_pair, g_ts, g_s0 = plan.evolve_backward(
            ctx.tape, states, grad_states.to(plan.cdtype).contiguous(),
            want_det=w_det, want_amp=w_amp, want_pair=w_pair, want_tsave=w_ts, want_state0=w_s0)
        if g_s0 is not None:
            g_s0 = g_s0.to(ctx.state_dtype)
        (ts_dev, ts_dt, det_dev, det_dt, amp_dev, amp_dt, pu_dev, pu_dt, det_shape, amp_shape) = ctx.meta
        if g_ts is not None:
            g_ts = g_ts.to(device=ts_dev, dtype=ts_dt)
        if g_det is not None:
            g_det = g_det.reshape(det_shape).to(device=det_dev)
            g_det = g_det.to(det_dt) if not det_dt.is_complex else g_det.to(det_dt)
        elif w_det:
            g_det = torch.zeros(det_shape, dtype=det_dt, device=det_dev)
        if g_amp is not None:
            g_amp = g_amp.reshape(amp_shape).to(device=amp_dev)
            g_amp = g_amp.to(amp_dt) if amp_dt.is_complex else g_amp.real.to(amp_dt)
        elif w_amp:
            g_amp = torch.zeros(amp_shape, dtype=amp_dt, device=amp_dev)
        if g_pair is not None:
            g_pair = g_pair.to(device=pu_dev, dtype=pu_dt)
        return (g_s0, g_ts, g_det, g_amp, g_pair) + (None,) * 8


def evolve(state0: Tensor, tsave: Tensor, det_values: Tensor, amp_values: Tensor, pair_u: Tensor, *,
           n_qubits: int, kind: int, dt: float, det_masks: Sequence[int], amp_masks: Sequence[int],
           collapse: Optional[Tensor] = None, solver: int = _cabi.SOLVER_DP5_SE,
           options: Optional[Options] = None) -> Tensor:
    """states[k] = state at tsave[k]; shape (n_t, batch, dim), state0's dtype (complex128 / complex64) and device."""
    opt = options or Options()
    return _EvolveFn.apply(state0, tsave, det_values, amp_values, pair_u, int(n_qubits), int(kind),
                           float(dt), list(det_masks), list(amp_masks), collapse, int(solver), opt)


class _EvolveUnitsFn(torch.autograd.Function):
    """Batch of independent parameter sets of ONE register (BASELINE configs[2]): unit u evolves
    ``state0[u]`` under its own coefficient tables ``det_values[u]`` / ``amp_values[u]``.  Kets with
    ``2*batch*2^N <= 128`` run all units in a single kernel launch (csrc/small_ket.cuh)."""

    @staticmethod
    def forward(ctx, state0: Tensor, tsave: Tensor, det_values: Tensor, amp_values: Tensor,
                pair_u: Tensor, n_qubits: int, dt: float, det_masks, amp_masks, opt: Options):
        n_units, batch = int(state0.shape[0]), int(state0.shape[1])
        cd = compute_dtype(state0.dtype, n_qubits, PD_KET)
        plan = get_plan(n_qubits, batch, PD_KET, state0.device, cd)
        # the plan carries the register, masks, dt and sample count; every unit brings its own tables, so
        # the plan's own are placeholders and one Program serves every call with the same structure
        # (no per-call reconfiguration, no device->host copy of a table)
        pu = pair_u.detach().to("cpu", torch.float64).contiguous()
        key = (n_qubits, float(dt), tuple(int(m) for m in det_masks), tuple(int(m) for m in amp_masks),
               int(det_values.shape[-1]), int(amp_values.shape[-1]), pu.numpy().tobytes())
        prog = _UNIT_PROGRAMS.get(key)
        if prog is None:
            prog = make_program(n_qubits, PD_KET, dt, det_masks,
                                torch.zeros(det_values.shape[1:], dtype=torch.float64), amp_masks,
                                torch.zeros(amp_values.shape[1:], dtype=torch.complex128), pu, None)
            if len(_UNIT_PROGRAMS) > 64:
                _UNIT_PROGRAMS.clear()
            _UNIT_PROGRAMS[key] = prog
        configure(plan, prog)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        states, tape = plan.evolve_forward_units(opt, state0.detach().to(cd), tsave, det_values, amp_values,
                                                 want_tape=need)
        ctx.plan, ctx.prog, ctx.tape = plan, prog, tape
        ctx.state_dtype = state0.dtype
        ctx.opt, ctx.state0, ctx.tsave = opt, state0.detach().to(cd), tsave.detach()
        ctx.tables = (det_values.detach(), amp_values.detach())
        ctx.meta = (det_values.device, det_values.dtype, amp_values.device, amp_values.dtype)
        ctx.save_for_backward(states)
        ctx.set_materialize_grads(False)
        return states if cd == state0.dtype else states.to(state0.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_states: Optional[Tensor]):
        none = (None,) * 10
        if grad_states is None or ctx.tape is None:
            return none
        (states,) = ctx.saved_tensors
        configure(ctx.plan, ctx.prog)
        dv, av = ctx.tables
        gs = grad_states.to(ctx.plan.cdtype).contiguous()
        try:
            g_det, g_amp, g_s0 = ctx.plan.evolve_backward_units(ctx.tape, states, gs, dv, av,
                                                                want_state0=ctx.needs_input_grad[0])
        except (RuntimeError, ValueError) as exc:
            # The stage tape of a batch lives in plan-owned device memory; another evolution on the same plan
            # (a second live graph of the same shape) has overwritten it.  The forward sweep is deterministic:
            # run it again from the saved inputs and differentiate the fresh tape.
            if "overwritten" not in str(exc) and "still current" not in str(exc):
                raise
            _, ctx.tape = ctx.plan.evolve_forward_units(ctx.opt, ctx.state0, ctx.tsave, dv, av, want_tape=True)
            g_det, g_amp, g_s0 = ctx.plan.evolve_backward_units(ctx.tape, states, gs, dv, av,
                                                                want_state0=ctx.needs_input_grad[0])
        if g_s0 is not None:
            g_s0 = g_s0.to(ctx.state_dtype)
        det_dev, det_dt, amp_dev, amp_dt = ctx.meta
        if g_det is not None:
            g_det = g_det.to(device=det_dev, dtype=det_dt)
        if g_amp is not None:
            g_amp = g_amp.to(device=amp_dev)
            g_amp = g_amp.to(amp_dt) if amp_dt.is_complex else g_amp.real.to(amp_dt)
        return (g_s0, None, g_det, g_amp) + (None,) * 6


def evolve_units(state0: Tensor, tsave: Tensor, det_values: Tensor, amp_values: Tensor, pair_u: Tensor, *,
                 n_qubits: int, dt: float, det_masks: Sequence[int], amp_masks: Sequence[int],
                 options: Optional[Options] = None) -> Tensor:
    """DP5_SE evolution of ``U`` independent parameter sets of one register.

    ``state0`` (U, batch, 2^N) on the device; ``det_values`` (U, n_det, n_samples) float64 and
    ``amp_values`` (U, n_amp, n_samples) complex128 (the reference's coefficient arrays, one set per
    unit); returns states (U, n_t, batch, 2^N).  Differentiable w.r.t. ``state0``, ``det_values`` and
    ``amp_values`` (not ``tsave`` / ``pair_u``: use :func:`evolve` per unit for those)."""
    return _EvolveUnitsFn.apply(state0, tsave, det_values, amp_values, pair_u, int(n_qubits), float(dt),
                                list(det_masks), list(amp_masks), options or Options())


def last_step_log(states: Tensor) -> list[dict]:
    """Attempted-step records of the evolution that produced ``states`` (needs a grad graph)."""
    fn = states.grad_fn
    while fn is not None and not hasattr(fn, "tape"):
        nxt = [f for f, _ in fn.next_functions if f is not None]
        fn = nxt[0] if nxt else None
    if fn is None or fn.tape is None:
        raise RuntimeError("no step log: the evolution was run without gradient tracking")
    return fn.tape.records()


# ------------------------------------------------------------------------------------------
# torch.library custom ops (hpsi / rhs carry autograd formulas; evolve_states / expect_diag are
# the non-differentiable entry points -- differentiable evolution is ops.evolve)
# ------------------------------------------------------------------------------------------
def _prog_from_args(n_qubits, kind, dt, det_masks, det_values, amp_masks, amp_values, pair_u, collapse):
    coll = collapse if collapse is not None and collapse.numel() > 0 else None
    return make_program(n_qubits, kind, dt, det_masks, det_values, amp_masks, amp_values, pair_u, coll)


@torch.library.custom_op("pulser_diff_b200::hpsi", mutates_args=())
def hpsi(psi: Tensor, t: float, det_values: Tensor, amp_values: Tensor, pair_u: Tensor,
         det_masks: List[int], amp_masks: List[int], dt: float) -> Tensor:
    """H(t) @ psi for psi of shape (batch, 2**N)."""
    n = int(pair_u.shape[0])
    cd = compute_dtype(psi.dtype, n, PD_KET)
    plan = get_plan(n, int(psi.shape[0]), PD_KET, psi.device, cd)
    configure(plan, _prog_from_args(n, PD_KET, dt, det_masks, det_values, amp_masks, amp_values,
                                    pair_u, None))
    return plan.hpsi(t, psi.to(cd)).to(psi.dtype)


@hpsi.register_fake
def _(psi, t, det_values, amp_values, pair_u, det_masks, amp_masks, dt):
    return torch.empty_like(psi)


def _vjp_common(ctx, cot: Tensor, kind: int, collapse):
    """Shared reverse mode of one generator application (C ABI pd_rhs_vjp)."""
    state, det_values, amp_values, pair_u = ctx.saved_tensors
    n = int(pair_u.shape[0])
    cd = compute_dtype(state.dtype, n, kind)
    plan = get_plan(n, int(state.shape[0]), kind, state.device, cd)
    configure(plan, _prog_from_args(n, kind, ctx.dt, ctx.det_masks, det_values, ctx.amp_masks,
                                    amp_values, pair_u, collapse))
    w_state, _, w_det, w_amp, w_pair = ctx.needs_input_grad[:5]
    g_state, g_det, g_amp, g_pair, _ = plan.rhs_vjp(ctx.t, state.to(cd), cot.to(cd).contiguous(),
                                                    want_state=w_state, want_det=w_det, want_amp=w_amp,
                                                    want_pair=w_pair)
    if g_state is not None:
        g_state = g_state.to(state.dtype)
    if w_det:
        g_det = (g_det if g_det is not None else torch.zeros(det_values.shape, dtype=torch.float64))
        g_det = g_det.to(device=det_values.device, dtype=det_values.dtype)
    if w_amp:
        g_amp = (g_amp if g_amp is not None else torch.zeros(amp_values.shape, dtype=torch.complex128))
        g_amp = g_amp.to(amp_values.device)
        g_amp = g_amp.to(amp_values.dtype) if amp_values.dtype.is_complex else g_amp.real.to(amp_values.dtype)
    if w_pair:
        g_pair = g_pair.to(device=pair_u.device, dtype=pair_u.dtype)
    return g_state, g_det, g_amp, g_pair


def _hpsi_setup(ctx, inputs, output):
    psi, t, det_values, amp_values, pair_u, det_masks, amp_masks, dt = inputs
    ctx.save_for_backward(psi, det_values, amp_values, pair_u)
    ctx.t, ctx.det_masks, ctx.amp_masks, ctx.dt = float(t), list(det_masks), list(amp_masks), float(dt)


def _hpsi_backward(ctx, cot):
    # H psi = i * (-i H psi): the cotangent on the right-hand side is -i * cot
    g_state, g_det, g_amp, g_pair = _vjp_common(ctx, -1j * cot, PD_KET, None)
    return g_state, None, g_det, g_amp, g_pair, None, None, None


hpsi.register_autograd(_hpsi_backward, setup_context=_hpsi_setup)


@torch.library.custom_op("pulser_diff_b200::rhs", mutates_args=())
def rhs(state: Tensor, t: float, det_values: Tensor, amp_values: Tensor, pair_u: Tensor,
        collapse: Tensor, det_masks: List[int], amp_masks: List[int], dt: float, kind: int) -> Tensor:
    """-i H(t) psi (kind 0) or the Lindblad right-hand side on vec(rho) (kind 1)."""
    n = int(pair_u.shape[0])
    cd = compute_dtype(state.dtype, n, kind)
    plan = get_plan(n, int(state.shape[0]), kind, state.device, cd)
    configure(plan, _prog_from_args(n, kind, dt, det_masks, det_values, amp_masks, amp_values,
                                    pair_u, collapse))
    return plan.hpsi(t, state.to(cd), rhs=True).to(state.dtype)


@rhs.register_fake
def _(state, t, det_values, amp_values, pair_u, collapse, det_masks, amp_masks, dt, kind):
    return torch.empty_like(state)


def _rhs_setup(ctx, inputs, output):
    state, t, det_values, amp_values, pair_u, collapse, det_masks, amp_masks, dt, kind = inputs
    ctx.save_for_backward(state, det_values, amp_values, pair_u)
    ctx.collapse, ctx.kind = collapse, int(kind)
    ctx.t, ctx.det_masks, ctx.amp_masks, ctx.dt = float(t), list(det_masks), list(amp_masks), float(dt)


def _rhs_backward(ctx, cot):
    g_state, g_det, g_amp, g_pair = _vjp_common(ctx, cot, ctx.kind, ctx.collapse)
    return g_state, None, g_det, g_amp, g_pair, None, None, None, None, None


rhs.register_autograd(_rhs_backward, setup_context=_rhs_setup)


@torch.library.custom_op("pulser_diff_b200::evolve_states", mutates_args=())
def evolve_states(state0: Tensor, tsave: Tensor, det_values: Tensor, amp_values: Tensor,
                  pair_u: Tensor, collapse: Tensor, det_masks: List[int], amp_masks: List[int],
                  dt: float, kind: int, solver: int, atol: float, rtol: float) -> Tensor:
    """Forward-only evolution; returns (n_t, batch, dim)."""
    n = int(pair_u.shape[0])
    cd = compute_dtype(state0.dtype, n, kind)
    plan = get_plan(n, int(state0.shape[0]), kind, state0.device, cd)
    configure(plan, _prog_from_args(n, kind, dt, det_masks, det_values, amp_masks, amp_values,
                                    pair_u, collapse))
    states, _ = plan.evolve_forward(solver, Options(atol=atol, rtol=rtol), state0.to(cd), tsave, False)
    return states.to(state0.dtype)


@evolve_states.register_fake
def _(state0, tsave, det_values, amp_values, pair_u, collapse, det_masks, amp_masks, dt, kind,
      solver, atol, rtol):
    return state0.new_empty((tsave.numel(),) + tuple(state0.shape))


@torch.library.custom_op("pulser_diff_b200::expect_diag", mutates_args=())
def expect_diag(states: Tensor, obs_diag: Tensor, kind: int) -> Tensor:
    """sum over batch of <psi|diag(obs)|psi> (kind 0) or tr(diag(obs) rho) (kind 1), per time."""
    n = int(obs_diag.numel()).bit_length() - 1
    cd = compute_dtype(states.dtype, n, kind)
    plan = get_plan(n, int(states.shape[1]), kind, states.device, cd)
    return plan.expect_diag(states.to(cd), obs_diag)


@expect_diag.register_fake
def _(states, obs_diag, kind):
    return torch.empty(states.shape[0], dtype=torch.complex128)
