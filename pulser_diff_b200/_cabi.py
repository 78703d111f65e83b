"""ctypes binding of ``include/pulser_diff_b200.h``.

The product loads ``pulser_diff_b200/libpulser_diff_b200.so`` (built for sm_100a by
``__graft_entry__.build()``) and nothing else: there is NO CPU fallback.  If the CUDA
library is missing, or a tensor is not on a CUDA device, the call raises.

Tests of the host-side logic may inject the host stand-in of the same ABI
(``tests/emu/libpd_emu.so``) with :func:`use_library`; that is the only way a non-CUDA
library is ever bound, and the package never does it by itself.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIBRARY = os.path.join(_HERE, "libpulser_diff_b200.so")
# complex64 build of the same sources and symbols (north_star's optional 1e-5 tier): state vectors are
# complex64 in device memory, the bandwidth-bound kernel families only (include/pulser_diff_b200.h)
DEFAULT_LIBRARY_C64 = os.path.join(_HERE, "libpulser_diff_b200_c64.so")

PD_KET, PD_DENSITY = 0, 1
SOLVER_DP5_SE, SOLVER_KRYLOV_SE, SOLVER_DP5_ME = 0, 1, 2
_ERR_INVALID = 1


class pd_options(C.Structure):
    _fields_ = [
        ("atol", C.c_double), ("rtol", C.c_double), ("max_steps", C.c_int64),
        ("safety_factor", C.c_double), ("min_factor", C.c_double), ("max_factor", C.c_double),
        ("max_krylov", C.c_int32), ("exp_tolerance", C.c_double), ("norm_tolerance", C.c_double),
        ("n_replay", C.c_int32), ("replay_dt", C.POINTER(C.c_double)),
        ("replay_clipped", C.POINTER(C.c_uint8)), ("path", C.c_int32),
    ]


class pd_step_record(C.Structure):
    _fields_ = [("t", C.c_double), ("dt", C.c_double), ("error", C.c_double),
                ("accepted", C.c_int32), ("clipped", C.c_int32), ("interval", C.c_int32),
                ("_pad", C.c_int32)]


# every symbol include/pulser_diff_b200.h declares (checked by tests/test_cabi_symbols.py)
EXPORTS = (
    "pd_abi_version", "pd_last_error", "pd_options_default", "pd_plan_create", "pd_plan_destroy",
    "pd_plan_set_interaction", "pd_plan_set_terms", "pd_plan_set_collapse", "pd_plan_set_path",
    "pd_hpsi", "pd_rhs", "pd_rhs_vjp", "pd_pair_gradient_flush",
    "pd_evolve_forward", "pd_evolve_backward", "pd_tape_n_records", "pd_tape_records",
    "pd_evolve_forward_units", "pd_evolve_backward_units", "pd_tape_unit_steps",
    "pd_tape_destroy", "pd_expect_diag", "pd_sharded_accumulate", "pd_sharded_accumulate_range", "pd_lincomb", "pd_dp5_error_sumsq", "pd_bench_hpsi", "pd_bench_dp5_steps",
    "pd_plan_launch_count", "pd_transfer_counters", "pd_is_cuda", "pd_amplitude_bytes",
)

_libs: dict = {}                       # complex64? -> loaded library
_lib_paths: dict = {False: None, True: None}


def _declare(lib: C.CDLL) -> None:
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pdbl, pu64 = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    lib.pd_abi_version.restype = C.c_int
    lib.pd_last_error.restype = C.c_char_p
    lib.pd_is_cuda.restype = C.c_int
    lib.pd_amplitude_bytes.restype = C.c_int
    lib.pd_options_default.argtypes = [C.POINTER(pd_options)]
    lib.pd_options_default.restype = None
    lib.pd_plan_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    lib.pd_plan_destroy.argtypes = [vp]
    lib.pd_plan_set_interaction.argtypes = [vp, pdbl, vp]
    lib.pd_plan_set_terms.argtypes = [vp, i32, dbl, i32, pu64, pdbl, i32, pu64, pdbl]
    lib.pd_plan_set_collapse.argtypes = [vp, i32, pdbl]
    lib.pd_plan_set_path.argtypes = [vp, i32]
    lib.pd_hpsi.argtypes = [vp, vp, dbl, vp, vp]
    lib.pd_rhs.argtypes = [vp, vp, dbl, vp, vp]
    lib.pd_evolve_forward.argtypes = [vp, vp, i32, C.POINTER(pd_options), vp, pdbl, i32, vp,
                                      C.POINTER(vp)]
    lib.pd_evolve_backward.argtypes = [vp, vp, vp, vp, vp, pdbl, pdbl, pdbl, pdbl, vp]
    lib.pd_evolve_forward_units.argtypes = [vp, vp, C.POINTER(pd_options), i32, vp, pdbl, i32, pdbl, pdbl, vp,
                                            C.POINTER(vp)]
    lib.pd_evolve_backward_units.argtypes = [vp, vp, vp, vp, vp, pdbl, pdbl, pdbl, pdbl, vp]
    lib.pd_tape_unit_steps.argtypes = [vp, i32, C.POINTER(i32)]
    lib.pd_tape_unit_steps.restype = i64
    lib.pd_tape_n_records.argtypes = [vp]
    lib.pd_tape_n_records.restype = i64
    lib.pd_tape_records.argtypes = [vp, C.POINTER(pd_step_record), i64]
    lib.pd_tape_destroy.argtypes = [vp]
    lib.pd_expect_diag.argtypes = [vp, vp, vp, i32, vp, pdbl]
    lib.pd_rhs_vjp.argtypes = [vp, vp, dbl, vp, vp, vp, pdbl, pdbl, pdbl, pdbl, i32]
    lib.pd_pair_gradient_flush.argtypes = [vp, vp, pdbl]
    lib.pd_lincomb.argtypes = [vp, vp, vp, i32, C.POINTER(vp), pdbl]
    lib.pd_dp5_error_sumsq.argtypes = [vp, vp, C.POINTER(vp), pdbl, vp, vp, dbl, dbl, pdbl]
    lib.pd_sharded_accumulate.argtypes = [vp, vp, vp, vp, dbl, i32, C.POINTER(vp), pdbl]
    lib.pd_sharded_accumulate_range.argtypes = [vp, vp, vp, vp, dbl, i32, C.POINTER(vp), pdbl, C.c_uint64]
    lib.pd_bench_hpsi.argtypes = [vp, vp, dbl, i32, vp, vp, pdbl]
    lib.pd_bench_dp5_steps.argtypes = [vp, vp, dbl, dbl, i32, vp, pdbl]
    lib.pd_plan_launch_count.argtypes = [vp]
    lib.pd_plan_launch_count.restype = i64
    lib.pd_transfer_counters.argtypes = [C.POINTER(i64), C.POINTER(i64), i32]


def transfer_counters(reset: bool = False) -> tuple[int, int]:
    """(host->device, device->host) bytes copied by the library so far (see pd_transfer_counters)."""
    a, b = C.c_int64(0), C.c_int64(0)
    _check(lib().pd_transfer_counters(C.byref(a), C.byref(b), 1 if reset else 0))
    return int(a.value), int(b.value)


def use_library(path: Optional[str], path_c64: Optional[str] = None) -> None:
    """Bind explicit libraries (tests only) or reset to the defaults with ``None``."""
    _libs.clear()
    _lib_paths[False], _lib_paths[True] = path, path_c64


def lib(torch_complex64: bool = False) -> C.CDLL:
    """The complex128 library (default) or the complex64 build of the same ABI."""
    key = bool(torch_complex64)
    if key not in _libs:
        path = _lib_paths[key] or (DEFAULT_LIBRARY_C64 if key else DEFAULT_LIBRARY)
        if key and _lib_paths[False] and not _lib_paths[True]:
            raise RuntimeError("an explicit complex128 library is bound without its complex64 counterpart")
        if not os.path.exists(path):
            raise RuntimeError(
                f"pulser_diff_b200: CUDA library not found at {path}. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                "There is no CPU fallback.")
        handle = C.CDLL(path)
        _declare(handle)
        if handle.pd_abi_version() != 1:
            raise RuntimeError("pulser_diff_b200: ABI version mismatch")
        if handle.pd_amplitude_bytes() != (8 if key else 16):
            raise RuntimeError(f"pulser_diff_b200: {path} is not the "
                               f"{'complex64' if key else 'complex128'} build")
        _libs[key] = handle
    return _libs[key]


def is_cuda_library() -> bool:
    return bool(lib().pd_is_cuda())


def _check(status: int, handle: Optional[C.CDLL] = None) -> None:
    if status == 0:
        return
    msg = (handle or lib()).pd_last_error().decode()
    if status == _ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def _require_device(t: torch.Tensor, what: str) -> None:
    if is_cuda_library():
        if not t.is_cuda:
            raise RuntimeError(
                f"pulser_diff_b200: {what} must live on a CUDA device (got {t.device}); "
                "this package has no CPU path.")
    elif t.is_cuda:
        raise RuntimeError("host stand-in library bound but tensor is on CUDA")


def _stream(device: torch.device) -> C.c_void_p:
    if device.type == "cuda":
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    return C.c_void_p(0)


def _dptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _hdbl(t: Optional[torch.Tensor]):
    if t is None:
        return C.POINTER(C.c_double)()
    return C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_double))


@dataclass
class Options:
    """Solver options forwarded by ``TorchEmulator.run(**options)`` (reference backend.py:435)."""
    atol: float = 1e-8
    rtol: float = 1e-6
    max_steps: int = 100_000
    safety_factor: float = 0.9
    min_factor: float = 0.2
    max_factor: float = 5.0
    max_krylov: int = 80
    exp_tolerance: float = 1e-10
    norm_tolerance: float = 1e-10
    use_sparse: bool = False          # accepted for API parity; the path is matrix free
    replay: Optional[Sequence] = None  # accepted steps [(dt, clipped), ...]: shared-step protocol
    path: int = 0                     # 0 auto, 1 gather, 2 tiled, 3 small-register, 4 stream kernels
    dtype: str = "complex128"         # "complex64": north_star's optional 1e-5 tier (state vectors in complex64;
                                      # the reference itself is complex128 only, backend.py:271, 280)

    @classmethod
    def from_dict(cls, d: Optional[dict]) -> "Options":
        d = dict(d or {})
        unknown = set(d) - set(cls.__dataclass_fields__)
        if unknown:
            raise TypeError(f"unknown solver option(s): {sorted(unknown)}")
        if "dtype" in d:
            d["dtype"] = str(d["dtype"]).replace("torch.", "")
            if d["dtype"] not in ("complex128", "complex64"):
                raise ValueError("dtype must be complex128 or complex64")
        return cls(**d)

    @property
    def state_dtype(self) -> torch.dtype:
        return torch.complex64 if self.dtype == "complex64" else torch.complex128


class Tape:
    """Owner of a ``pd_tape*`` (step log of one forward evolution)."""

    def __init__(self, ptr: int, handle: Optional[C.CDLL] = None) -> None:
        self._ptr = C.c_void_p(ptr)
        self._lib = handle or lib()

    @property
    def ptr(self) -> C.c_void_p:
        return self._ptr

    def records(self) -> list[dict]:
        n = self._lib.pd_tape_n_records(self._ptr)
        buf = (pd_step_record * max(n, 1))()
        _check(self._lib.pd_tape_records(self._ptr, buf, n), self._lib)
        return [dict(t=r.t, dt=r.dt, error=r.error, accepted=bool(r.accepted),
                     clipped=bool(r.clipped), interval=r.interval) for r in buf[:n]]

    def __del__(self) -> None:
        try:
            if self._ptr:
                self._lib.pd_tape_destroy(self._ptr)
                self._ptr = C.c_void_p(0)
        except Exception:
            pass


class Plan:
    """Owner of a ``pd_plan*``: one register geometry + workspace on one device."""

    def __init__(self, n_qubits: int, batch: int, kind: int, device: torch.device,
                 dtype: torch.dtype = torch.complex128) -> None:
        self.n_qubits, self.batch, self.kind = int(n_qubits), int(batch), int(kind)
        self.device = torch.device(device)
        if dtype not in (torch.complex128, torch.complex64):
            raise TypeError("state vectors are complex128 or complex64")
        self.cdtype = dtype                       # dtype of every amplitude tensor of this plan
        self._lib = lib(dtype == torch.complex64)
        if is_cuda_library() and self.device.type != "cuda":
            raise RuntimeError(
                f"pulser_diff_b200 runs on CUDA devices only (asked for {self.device}); "
                "there is no CPU fallback.")
        ordinal = self.device.index if self.device.index is not None else (
            torch.cuda.current_device() if self.device.type == "cuda" else 0)
        p = C.c_void_p()
        self._ck(self._lib.pd_plan_create(C.byref(p), self.n_qubits, self.batch, self.kind, ordinal))
        self._ptr = p
        self.dim = 2 ** (self.n_qubits if kind == PD_KET else 2 * self.n_qubits)
        self.program_id = -1
        self.n_det = self.n_amp = self.n_samples = 0

    def _ck(self, status: int) -> None:
        _check(status, self._lib)

    def __del__(self) -> None:
        try:
            if self._ptr:
                self._lib.pd_plan_destroy(self._ptr)
                self._ptr = C.c_void_p(0)
        except Exception:
            pass

    # ---- setup -----------------------------------------------------------------------------
    def set_interaction(self, pair_u: torch.Tensor) -> None:
        u = pair_u.detach().to("cpu", torch.float64).contiguous()
        if tuple(u.shape) != (self.n_qubits, self.n_qubits):
            raise ValueError("pair_u must be (N, N)")
        self._ck(self._lib.pd_plan_set_interaction(self._ptr, _hdbl(u), _stream(self.device)))

    def set_terms(self, dt: float, det_masks: Sequence[int], det_values: torch.Tensor,
                  amp_masks: Sequence[int], amp_values: torch.Tensor) -> None:
        dv = det_values.detach().to("cpu", torch.float64).contiguous()
        av = torch.view_as_real(amp_values.detach().to("cpu", torch.complex128).contiguous()).contiguous()
        n_det, n_amp = len(det_masks), len(amp_masks)
        n_samples = int(dv.shape[1]) if n_det else (int(av.shape[1]) if n_amp else 2)
        if n_det and tuple(dv.shape) != (n_det, n_samples):
            raise ValueError("det_values must be (n_det, n_samples)")
        if n_amp and tuple(av.shape) != (n_amp, n_samples, 2):
            raise ValueError("amp_values must be (n_amp, n_samples)")
        dm = (C.c_uint64 * max(n_det, 1))(*det_masks)
        am = (C.c_uint64 * max(n_amp, 1))(*amp_masks)
        self._ck(self._lib.pd_plan_set_terms(self._ptr, n_samples, float(dt), n_det, dm, _hdbl(dv),
                                       n_amp, am, _hdbl(av)))
        self.n_det, self.n_amp, self.n_samples = n_det, n_amp, n_samples

    def set_collapse(self, ops: Optional[torch.Tensor]) -> None:
        if ops is None or ops.numel() == 0:
            self._ck(self._lib.pd_plan_set_collapse(self._ptr, 0, C.POINTER(C.c_double)()))
            return
        o = torch.view_as_real(ops.detach().to("cpu", torch.complex128).contiguous()).contiguous()
        if tuple(o.shape[1:]) != (2, 2, 2):
            raise ValueError("collapse operators must be (n_ops, 2, 2) complex")
        self._ck(self._lib.pd_plan_set_collapse(self._ptr, int(o.shape[0]), _hdbl(o)))

    def set_path(self, path: int) -> None:
        self._ck(self._lib.pd_plan_set_path(self._ptr, int(path)))

    # ---- applications ------------------------------------------------------------------------
    def _vec(self, t: torch.Tensor, what: str, lead: tuple = ()) -> torch.Tensor:
        _require_device(t, what)
        if t.dtype != self.cdtype:
            raise TypeError(f"{what} must be {str(self.cdtype).replace('torch.', '')}")
        want = lead + (self.batch, self.dim)
        if tuple(t.shape) != want:
            raise ValueError(f"{what} must have shape {want}, got {tuple(t.shape)}")
        return t.contiguous()

    def hpsi(self, t: float, psi: torch.Tensor, rhs: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        psi = self._vec(psi, "psi")
        if out is None:
            out = torch.empty_like(psi)
        elif not out.is_contiguous() or out.data_ptr() == psi.data_ptr():
            raise ValueError("out must be a contiguous buffer distinct from psi")
        else:
            self._vec(out, "out")
        fn = self._lib.pd_rhs if rhs else self._lib.pd_hpsi
        self._ck(fn(self._ptr, _stream(self.device), float(t), _dptr(psi), _dptr(out)))
        return out

    def rhs_vjp(self, t: float, state: torch.Tensor, cot: torch.Tensor, want_state: bool = True,
                want_det: bool = True, want_amp: bool = True, want_pair: bool = False,
                defer_pair: bool = False):
        """Reverse mode of ``k = rhs(t, state)``: returns ``(grad_state, g_det, g_amp, g_pair,
        g_t)`` for the cotangent ``cot`` on ``k`` (entries not asked for are None).
        ``defer_pair``: accumulate the interaction weights inside the plan; collect dL/dU_ij
        once with :meth:`pair_gradient_flush`."""
        state = self._vec(state, "state")
        cot = self._vec(cot, "cot")
        g_state = torch.empty_like(state) if want_state else None
        g_det = torch.zeros((self.n_det, self.n_samples), dtype=torch.float64) if want_det and self.n_det else None
        g_amp = torch.zeros((self.n_amp, self.n_samples, 2), dtype=torch.float64) if want_amp and self.n_amp else None
        g_pair = torch.zeros((self.n_qubits, self.n_qubits), dtype=torch.float64) if want_pair and not defer_pair else None
        g_t = C.c_double(0.0)
        self._ck(self._lib.pd_rhs_vjp(self._ptr, _stream(self.device), float(t), _dptr(state), _dptr(cot),
                                _dptr(g_state), _hdbl(g_det), _hdbl(g_amp), _hdbl(g_pair), C.byref(g_t),
                                int(bool(defer_pair))))
        if g_amp is not None:
            g_amp = torch.view_as_complex(g_amp)
        return g_state, g_det, g_amp, g_pair, float(g_t.value)

    def pair_gradient_flush(self) -> torch.Tensor:
        """dL/dU_ij of every ``rhs_vjp(..., defer_pair=True)`` since the last flush."""
        g_pair = torch.zeros((self.n_qubits, self.n_qubits), dtype=torch.float64)
        self._ck(self._lib.pd_pair_gradient_flush(self._ptr, _stream(self.device), _hdbl(g_pair)))
        return g_pair

    def evolve_forward(self, solver: int, opt: Options, state0: torch.Tensor, tsave: torch.Tensor,
                       want_tape: bool) -> tuple[torch.Tensor, Optional[Tape]]:
        state0 = self._vec(state0, "state0")
        ts = tsave.detach().to("cpu", torch.float64).contiguous()
        n_t = int(ts.numel())
        states = torch.empty((n_t, self.batch, self.dim), dtype=self.cdtype, device=state0.device)
        o = pd_options()
        self._lib.pd_options_default(C.byref(o))
        for k in ("atol", "rtol", "max_steps", "safety_factor", "min_factor", "max_factor",
                  "max_krylov", "exp_tolerance", "norm_tolerance", "path"):
            setattr(o, k, getattr(opt, k))
        keep = None
        if opt.replay:
            n = len(opt.replay)
            dts = (C.c_double * n)(*[float(r[0]) for r in opt.replay])
            cl = (C.c_uint8 * n)(*[1 if r[1] else 0 for r in opt.replay])
            o.n_replay, o.replay_dt, o.replay_clipped = n, dts, cl
            keep = (dts, cl)
        tape_ptr = C.c_void_p()
        self._ck(self._lib.pd_evolve_forward(self._ptr, _stream(self.device), int(solver), C.byref(o),
                                       _dptr(state0), _hdbl(ts), n_t, _dptr(states),
                                       C.byref(tape_ptr) if want_tape else None))
        del keep
        return states, (Tape(tape_ptr.value, self._lib) if want_tape else None)

    def _options_struct(self, opt: Options):
        o = pd_options()
        self._lib.pd_options_default(C.byref(o))
        for k in ("atol", "rtol", "max_steps", "safety_factor", "min_factor", "max_factor",
                  "max_krylov", "exp_tolerance", "norm_tolerance", "path"):
            setattr(o, k, getattr(opt, k))
        return o

    def evolve_forward_units(self, opt: Options, state0: torch.Tensor, tsave: torch.Tensor,
                             det_values: torch.Tensor, amp_values: torch.Tensor,
                             want_tape: bool) -> tuple[torch.Tensor, Optional[Tape]]:
        """Batch of parameter sets: ``state0`` (U, batch, dim) on the device, ``det_values``
        (U, n_det, n_samples) float64 / ``amp_values`` (U, n_amp, n_samples) complex128 on the
        host.  Returns states (U, n_t, batch, dim)."""
        n_units = int(state0.shape[0])
        state0 = self._vec(state0, "state0", (n_units,))
        ts = tsave.detach().to("cpu", torch.float64).contiguous()
        n_t = int(ts.numel())
        dv, av, dvp, avp = self._unit_tables(det_values, amp_values, n_units)
        states = torch.empty((n_units, n_t, self.batch, self.dim), dtype=self.cdtype, device=state0.device)
        o = self._options_struct(opt)
        tape_ptr = C.c_void_p()
        self._ck(self._lib.pd_evolve_forward_units(self._ptr, _stream(self.device), C.byref(o), n_units,
                                             _dptr(state0), _hdbl(ts), n_t, dvp, avp,
                                             _dptr(states), C.byref(tape_ptr) if want_tape else None))
        return states, (Tape(tape_ptr.value, self._lib) if want_tape else None)

    def _unit_tables(self, det_values: torch.Tensor, amp_values: torch.Tensor, n_units: int):
        """Per-unit coefficient tables as (tensors kept alive, double* det, double* amp).  Tables that
        already live on the plan's device cross the ABI as device pointers (no host round trip); the
        library tells the two apart (include/pulser_diff_b200.h, pd_evolve_forward_units)."""
        on_dev = det_values.is_cuda and amp_values.is_cuda and det_values.device == amp_values.device
        where = det_values.device if on_dev else "cpu"
        dv = det_values.detach().to(where, torch.float64).contiguous()
        av = torch.view_as_real(amp_values.detach().to(where, torch.complex128).contiguous()).contiguous()
        if tuple(dv.shape) != (n_units, self.n_det, self.n_samples) or \
                tuple(av.shape) != (n_units, self.n_amp, self.n_samples, 2):
            raise ValueError("per-unit coefficient tables must be (U, n_terms, n_samples) matching the plan")
        pd_ = C.POINTER(C.c_double)
        if on_dev:
            return dv, av, C.cast(C.c_void_p(dv.data_ptr()), pd_), C.cast(C.c_void_p(av.data_ptr()), pd_)
        return dv, av, _hdbl(dv), _hdbl(av)

    def evolve_backward_units(self, tape: Tape, states: torch.Tensor, grad_states: torch.Tensor,
                              det_values: torch.Tensor, amp_values: torch.Tensor, want_state0: bool):
        n_units, n_t = int(states.shape[0]), int(states.shape[1])
        states = self._vec(states, "states", (n_units, n_t))
        grad_states = self._vec(grad_states, "grad_states", (n_units, n_t))
        dv, av, dvp, avp = self._unit_tables(det_values, amp_values, n_units)
        where = dv.device
        g_det = torch.zeros((n_units, self.n_det, self.n_samples), dtype=torch.float64, device=where) if self.n_det else None
        g_amp = torch.zeros((n_units, self.n_amp, self.n_samples, 2), dtype=torch.float64, device=where) if self.n_amp else None
        g_s0 = torch.empty((n_units, self.batch, self.dim), dtype=self.cdtype,
                           device=states.device) if want_state0 else None
        pd_ = C.POINTER(C.c_double)

        def ptr(x):
            if x is None:
                return None
            return C.cast(C.c_void_p(x.data_ptr()), pd_) if x.is_cuda else _hdbl(x)

        self._ck(self._lib.pd_evolve_backward_units(self._ptr, _stream(self.device), tape.ptr, _dptr(states),
                                              _dptr(grad_states), dvp, avp, ptr(g_det), ptr(g_amp), _dptr(g_s0)))
        if g_amp is not None:
            g_amp = torch.view_as_complex(g_amp)
        return g_det, g_amp, g_s0

    def unit_steps(self, tape: Tape, unit: int) -> tuple[int, int]:
        att = C.c_int32(0)
        acc = self._lib.pd_tape_unit_steps(tape.ptr, int(unit), C.byref(att))
        return int(acc), int(att.value)

    def evolve_backward(self, tape: Tape, states: torch.Tensor, grad_states: torch.Tensor,
                        want_det: bool, want_amp: bool, want_pair: bool, want_tsave: bool,
                        want_state0: bool):
        n_t = int(states.shape[0])
        states = self._vec(states, "states", (n_t,))
        grad_states = self._vec(grad_states, "grad_states", (n_t,))
        g_det = torch.zeros((self.n_det, self.n_samples), dtype=torch.float64) if want_det and self.n_det else None
        g_amp = torch.zeros((self.n_amp, self.n_samples, 2), dtype=torch.float64) if want_amp and self.n_amp else None
        g_pair = torch.zeros((self.n_qubits, self.n_qubits), dtype=torch.float64) if want_pair else None
        g_ts = torch.zeros(n_t, dtype=torch.float64) if want_tsave else None
        g_s0 = torch.empty((self.batch, self.dim), dtype=self.cdtype, device=states.device) if want_state0 else None
        self._ck(self._lib.pd_evolve_backward(self._ptr, _stream(self.device), tape.ptr, _dptr(states),
                                        _dptr(grad_states), _hdbl(g_det), _hdbl(g_amp), _hdbl(g_pair),
                                        _hdbl(g_ts), _dptr(g_s0)))
        if g_amp is not None:
            g_amp = torch.view_as_complex(g_amp)
        return g_det, g_amp, g_pair, g_ts, g_s0

    def expect_diag(self, states: torch.Tensor, obs_diag: torch.Tensor) -> torch.Tensor:
        n_t = int(states.shape[0])
        states = self._vec(states, "states", (n_t,))
        _require_device(obs_diag, "obs_diag")
        obs = obs_diag.to(torch.float64).contiguous()
        if obs.numel() != 2 ** self.n_qubits:
            raise ValueError("diagonal observable must have 2**N entries")
        out = torch.zeros(n_t, 2, dtype=torch.float64)
        self._ck(self._lib.pd_expect_diag(self._ptr, _stream(self.device), _dptr(states), n_t, _dptr(obs),
                                    _hdbl(out)))
        return torch.view_as_complex(out)

    def lincomb(self, out: torch.Tensor, ins: Sequence[torch.Tensor], w: Sequence[float]) -> torch.Tensor:
        """``out = sum_j w[j] * ins[j]`` in one pass (at most 8 inputs)."""
        if not out.is_contiguous():
            raise ValueError("out must be contiguous (it is written in place)")
        self._vec(out, "out")
        ins = [self._vec(x, "input") for x in ins]
        if len(ins) != len(w) or not 1 <= len(ins) <= 8:
            raise ValueError("1..8 inputs, one weight each")
        ptrs = (C.c_void_p * len(ins))(*[C.c_void_p(x.data_ptr()) for x in ins])
        ws = (C.c_double * len(ins))(*[float(x) for x in w])
        self._ck(self._lib.pd_lincomb(self._ptr, _stream(self.device), _dptr(out), len(ins), ptrs, ws))
        return out

    def dp5_error_sumsq(self, k: Sequence[Optional[torch.Tensor]], ew: Sequence[float], y0: torch.Tensor,
                        y1: torch.Tensor, atol: float, rtol: float) -> torch.Tensor:
        """Per-column local share of the DP5 error norm (see include/pulser_diff_b200.h)."""
        if len(k) != 7 or len(ew) != 7:
            raise ValueError("seven slopes and seven weights")
        ks = [None if x is None else self._vec(x, "k") for x in k]
        ptrs = (C.c_void_p * 7)(*[C.c_void_p(0 if x is None else x.data_ptr()) for x in ks])
        ws = (C.c_double * 7)(*[float(x) for x in ew])
        out = torch.zeros(self.batch, dtype=torch.float64)
        self._ck(self._lib.pd_dp5_error_sumsq(self._ptr, _stream(self.device), ptrs, ws,
                                        _dptr(self._vec(y0, "y0")), _dptr(self._vec(y1, "y1")),
                                        float(atol), float(rtol), _hdbl(out)))
        return out

    def sharded_accumulate(self, out: torch.Tensor, psi: torch.Tensor, shift: float,
                           peer_ptrs: Sequence[int], coefs: Sequence[complex]) -> None:
        """``out += shift*psi + sum_k coefs[k] * slice_at(peer_ptrs[k])`` (in place).

        ``peer_ptrs`` are raw device addresses of the partner ranks' slices (peer-mapped memory
        from ``torch.distributed._symmetric_memory``, or local tensors' ``data_ptr()``)."""
        if not out.is_contiguous():
            raise ValueError("out must be contiguous (it is updated in place)")
        out = self._vec(out, "out")
        psi = self._vec(psi, "psi")
        k = len(peer_ptrs)
        if k != len(coefs):
            raise ValueError("one coefficient per peer slice")
        ptrs = (C.c_void_p * max(k, 1))(*[C.c_void_p(int(a)) for a in peer_ptrs])
        cf = (C.c_double * max(2 * k, 1))()
        for i, c in enumerate(coefs):
            cf[2 * i], cf[2 * i + 1] = complex(c).real, complex(c).imag
        self._ck(self._lib.pd_sharded_accumulate(self._ptr, _stream(self.device), _dptr(out), _dptr(psi),
                                           float(shift), k, ptrs, cf))

    def sharded_accumulate_range(self, out_ptr: int, psi_ptr: int, shift: float, peer_ptrs: Sequence[int],
                                 coefs: Sequence[complex], n_amp: int) -> None:
        """:meth:`sharded_accumulate` on ``n_amp`` amplitudes starting at the given raw addresses (the
        caller offsets every pointer to the start of the range)."""
        k = len(peer_ptrs)
        if k != len(coefs):
            raise ValueError("one coefficient per peer slice")
        ptrs = (C.c_void_p * max(k, 1))(*[C.c_void_p(int(a)) for a in peer_ptrs])
        cf = (C.c_double * max(2 * k, 1))()
        for i, c in enumerate(coefs):
            cf[2 * i], cf[2 * i + 1] = complex(c).real, complex(c).imag
        self._ck(self._lib.pd_sharded_accumulate_range(self._ptr, _stream(self.device), C.c_void_p(int(out_ptr)),
                                                 C.c_void_p(int(psi_ptr)), float(shift), k, ptrs, cf,
                                                 C.c_uint64(int(n_amp))))

    def bench_hpsi(self, t: float, psi: torch.Tensor, reps: int) -> float:
        """Average device ms of one H(t)·psi (CUDA events on the current stream)."""
        psi = self._vec(psi, "psi")
        out = torch.empty_like(psi)
        ms = C.c_double()
        self._ck(self._lib.pd_bench_hpsi(self._ptr, _stream(self.device), float(t), int(reps), _dptr(psi),
                                   _dptr(out), C.byref(ms)))
        return ms.value

    def bench_dp5_steps(self, t0: float, dt: float, steps: int, y: torch.Tensor) -> float:
        """Average device ms of one fixed-size DP5 step (updates ``y`` in place)."""
        y = self._vec(y, "y")
        ms = C.c_double()
        self._ck(self._lib.pd_bench_dp5_steps(self._ptr, _stream(self.device), float(t0), float(dt),
                                        int(steps), _dptr(y), C.byref(ms)))
        return ms.value

    @property
    def launch_count(self) -> int:
        return int(self._lib.pd_plan_launch_count(self._ptr))
