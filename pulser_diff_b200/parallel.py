"""Multi-GPU partitioning of the evolution path.  One process per GPU.

Two axes, exactly the two ``north_star`` names:

1. **Independent parameter sets** (configs[2]: batches of pulse-parameter sets).  Units are
   dealt round-robin to ranks; every rank evolves its own units with the single-GPU engine; there
   is NO data-path collective.  Only the small results (losses / gradients) are gathered at the
   end (:func:`gather_results`).
2. **One large register sharded by its highest-order qubits** (configs[4]).  State index MSBs =
   register qubits 0..g-1 = rank bits (qubit 0 is the most significant bit), so
   each rank owns 2^(N-g) contiguous amplitudes and its own slice of the interaction diagonal.
   :class:`ShardedKet` applies ``H(t)`` as

       local part   : the single-GPU kernels on the N-g local qubits, with the global qubits'
                      interaction folded into per-qubit detunings + one energy shift, and
       global flips : the slice of rank ^ (1 << k) per global qubit.  On GPUs
                      (``peer_memory=True``) every rank keeps its slice in a symmetric buffer the
                      peers have mapped over NVLink.  The copy engines pull the partner slices
                      on a second stream while the local kernels (HBM-bound, all SMs) run, then
                      ONE kernel (``pd_sharded_accumulate``) folds every flip into the result.
                      ``peer_memory="read"`` skips the local copies: the same kernel reads the
                      partner slices in place over NVLink (no receive buffers; measured slower,
                      profiles/r01_multi_gpu.md).  Otherwise (gloo in the CPU tests) one pairwise
                      send/recv per global qubit, posted before the local kernels.

   The reference has nothing to mirror here (single process, no collectives).
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from . import _cabi, ops

C128 = torch.complex128


# ---------------------------------------------------------------------------------------------
# axis 1: independent units
# ---------------------------------------------------------------------------------------------
def shard_units(n_units: int, rank: Optional[int] = None, world: Optional[int] = None) -> list[int]:
    """Indices of the units this rank owns (round-robin; balanced to within one unit)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return list(range(rank, n_units, world))


def run_units(n_units: int, fn: Callable[[int], Tensor]) -> dict[int, Tensor]:
    """Evaluate ``fn(unit)`` for this rank's units.  No communication."""
    return {u: fn(u) for u in shard_units(n_units)}


def gather_results(local: dict[int, Tensor], n_units: int, device=None) -> Tensor:
    """Stack per-unit result tensors (same shape everywhere) from all ranks, in unit order.

    The only collective of axis 1; it moves results, not states.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    sample = next(iter(local.values())) if local else None
    shape = [list(sample.shape), str(sample.dtype)] if sample is not None else None
    shapes = [None] * world
    dist.all_gather_object(shapes, shape)
    shape = next(s for s in shapes if s is not None)
    dtype = getattr(torch, shape[1].split(".")[-1])
    device = device or (sample.device if sample is not None else "cpu")
    per_rank = (n_units + world - 1) // world
    buf = torch.zeros([per_rank] + shape[0], dtype=dtype, device=device)
    for slot, u in enumerate(shard_units(n_units, rank, world)):
        buf[slot] = local[u].detach().to(device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    out = torch.zeros([n_units] + shape[0], dtype=dtype, device=device)
    for r in range(world):
        for slot, u in enumerate(shard_units(n_units, r, world)):
            out[u] = bufs[r][slot]
    return out


# ---------------------------------------------------------------------------------------------
# axis 2: one register sharded by its top qubits
# ---------------------------------------------------------------------------------------------
class ShardedKet:
    """H(t) for a register whose state vector is sharded over ``world = 2^g`` ranks.

    Args:
        n_qubits: total register size N.
        pair_u:   (N, N) upper-triangular couplings C6 / r_ij^6.
        dt, det_masks, det_values, amp_masks, amp_values: the term structure that crosses the
                  C ABI (same meaning as in :func:`pulser_diff_b200.ops.evolve`).
        device:   this rank's device.
        peer_memory: ``True``/``"copy"``: partner slices come through NVLink peer memory (CUDA
                  ranks of one node), pulled by the copy engines beside the local kernels;
                  ``"read"``: read in place by the accumulation kernel; ``False``: send/recv.
        dtype:    ``torch.complex128`` (default) or ``torch.complex64`` -- the optional 1e-5 tier: slices,
                  exchange buffers and every stage vector in complex64 (the complex64 build of the library),
                  which halves the HBM and the NVLink bytes and fits one more qubit per GPU (2^30 amplitudes
                  = N = 33 on eight 180 GB GPUs).
    """

    def __init__(self, n_qubits: int, pair_u: Tensor, dt: float, det_masks: Sequence[int],
                 det_values: Tensor, amp_masks: Sequence[int], amp_values: Tensor,
                 device: torch.device, group=None, peer_memory: bool | str = False,
                 dtype: torch.dtype = torch.complex128) -> None:
        self.group = group
        if dtype not in (torch.complex128, torch.complex64):
            raise TypeError("sharded registers are complex128 or complex64")
        self.cdtype = dtype
        self._amp_bytes = 16 if dtype == torch.complex128 else 8
        # without a process group: one rank owning the whole register (no global qubits)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        g = self.world.bit_length() - 1
        if 1 << g != self.world:
            raise ValueError("the number of ranks must be a power of two")
        if g >= n_qubits:
            raise ValueError("more shard bits than qubits")
        self.n, self.g, self.nl = n_qubits, g, n_qubits - g
        self.device = torch.device(device)
        self.dt = float(dt)
        self.n_samples = int(det_values.shape[1]) if len(det_masks) else int(amp_values.shape[1])
        self.det_masks, self.amp_masks = list(det_masks), list(amp_masks)
        self.det_values = det_values.detach().to("cpu", torch.float64)
        self.amp_values = amp_values.detach().to("cpu", C128)
        u = pair_u.detach().to("cpu", torch.float64)
        # Rydberg occupation of the global qubits on this rank: qubit q <-> rank bit (g-1-q),
        # bit value 0 = Rydberg.
        self.r_glob = [1 - ((self.rank >> (g - 1 - q)) & 1) for q in range(g)]
        # local plan on the N-g low qubits (register qubits g..N-1 keep their order)
        loc = slice(g, n_qubits)
        self.plan = ops.get_plan(self.nl, 1, _cabi.PD_KET, self.device, dtype)
        low = (1 << self.nl) - 1
        dm = [(m >> g) & low for m in self.det_masks]          # mask bit q -> local bit q-g
        am = [(m >> g) & low for m in self.amp_masks]
        dv, av = self.det_values, self.amp_values
        # interaction of local qubit j with the occupied global qubits = static detuning on j;
        # det coefficient c enters H as 2*c*n_j (reference hamiltonian.py:537-540)
        extra_m, extra_v = [], []
        for j in range(self.nl):
            shift = sum(float(u[q, g + j]) * self.r_glob[q] for q in range(g))
            if shift != 0.0:
                extra_m.append(1 << j)
                extra_v.append(torch.full((self.n_samples,), 0.5 * shift, dtype=torch.float64))
        keep = [k for k, m in enumerate(dm) if m]
        self._keep_det, self._extra_qubits = keep, [m.bit_length() - 1 for m in extra_m]
        dm_all = [dm[k] for k in keep] + extra_m
        dv_all = torch.stack([dv[k] for k in keep] + extra_v) if dm_all else torch.zeros(0, self.n_samples)
        keep_a = [k for k, m in enumerate(am) if m]
        self._keep_amp = keep_a
        am_all = [am[k] for k in keep_a]
        av_all = torch.stack([av[k] for k in keep_a]) if am_all else torch.zeros(0, self.n_samples, dtype=C128)
        self._prog = ops.make_program(self.nl, _cabi.PD_KET, self.dt, dm_all, dv_all, am_all, av_all,
                                      u[loc, loc].contiguous(), None)
        # energy shift from global-global interaction (static) -- detuning part is time dependent
        self.e_static = sum(float(u[p, q]) * self.r_glob[p] * self.r_glob[q]
                            for p in range(g) for q in range(p + 1, g))
        self._sym = self._hdl = self._side = self._events = None
        self._pull_chunks = 8
        self._mode = "read" if peer_memory == "read" else "copy"
        if peer_memory:
            if self.device.type != "cuda":
                raise ValueError("peer_memory needs CUDA ranks")
            import torch.distributed._symmetric_memory as symm_mem
            # real view of the complex slice: (re, im) pairs, the layout the kernels read
            self._sym = symm_mem.empty((1, 2 << self.nl), device=self.device,
                                       dtype=torch.float64 if dtype == torch.complex128 else torch.float32)
            self._hdl = symm_mem.rendezvous(self._sym, group if group is not None else dist.group.WORLD)
            self._peer_ptrs = [int(p) for p in self._hdl.buffer_ptrs]

    def state_buffer(self) -> Tensor:
        """The peer-visible (1, 2^(N-g)) slice buffer.  A state kept here is read by the partner
        ranks without the staging copy :meth:`hpsi` otherwise makes."""
        if self._sym is None:
            raise RuntimeError("state_buffer() needs peer_memory=True")
        return torch.view_as_complex(self._sym.view(1, 1 << self.nl, 2))

    # -- scalar coefficients of the global qubits at time t (reference interpolation rule) -----
    def _interp(self, values: Tensor, t: float):
        n = self.n_samples
        i1 = max(int(min(math.floor(t / self.dt), n - 2)), 0)
        i2 = min(i1 + 1, n - 2)
        return values[i1] + (values[i2] - values[i1]) * (t - i1 * self.dt) / self.dt

    def _global_coefficients(self, t: float):
        d = [0.0] * self.g
        c = [0j] * self.g
        for m, v in zip(self.det_masks, self.det_values):
            val = 2.0 * float(self._interp(v, t))
            for q in range(self.g):
                if m >> q & 1:
                    d[q] += val
        for m, v in zip(self.amp_masks, self.amp_values):
            val = complex(self._interp(v, t))
            for q in range(self.g):
                if m >> q & 1:
                    c[q] += val
        return d, c

    def hpsi(self, t: float, psi_local: Tensor, keep_partners: bool = False, rhs: bool = False,
             out: Optional[Tensor] = None):
        """``(H(t) psi)`` restricted to this rank's slice (``rhs``: ``-i H(t) psi``, the factor
        folded into the kernels' coefficients).  ``psi_local``: (1, 2^(N-g)).

        ``keep_partners``: also return the partner slices (one per global qubit, valid until the
        next call) -- the adjoint sweep takes its drive correlations from them."""
        d, c = self._global_coefficients(t)
        if self._hdl is not None:
            return self._hpsi_peer(t, psi_local, d, c, keep_partners, rhs, out)
        # 1. post the pairwise exchanges (one per global qubit) before any local work
        recv = [torch.empty_like(psi_local) for _ in range(self.g)]
        reqs = []
        for q in range(self.g):
            peer = self.rank ^ (1 << (self.g - 1 - q))
            ops_ = [dist.P2POp(dist.isend, psi_local, peer, self.group),
                    dist.P2POp(dist.irecv, recv[q], peer, self.group)]
            reqs += dist.batch_isend_irecv(ops_)
        # 2. local qubits: the single-GPU kernels (overlaps the transfers)
        ops.configure(self.plan, self._prog)
        out = self.plan.hpsi(t, psi_local, rhs=rhs, out=out)
        shift = self.e_static + sum(d[q] * self.r_glob[q] for q in range(self.g))
        # 3. global flips: this rank's bit for qubit q is 1 (ground) -> coefficient c, else conj(c)
        for r in reqs:
            r.wait()
        f = -1j if rhs else 1.0
        coefs = [f * shift] + [f * (c[q] if self.r_glob[q] == 0 else c[q].conjugate()) for q in range(self.g)]
        self.plan.sharded_accumulate(out, psi_local, 0.0, [psi_local.data_ptr()] + [r.data_ptr() for r in recv],
                                     coefs)
        return (out, recv) if keep_partners else out

    def _hpsi_peer(self, t: float, psi_local: Tensor, d, c, keep_partners: bool = False,
                   rhs: bool = False, out: Optional[Tensor] = None):
        """Peer-memory variants.  "read": one kernel accumulates the partner slices in place over
        NVLink.  "copy": the copy engines pull the partner slices into local buffers on a second
        stream while the local kernels run; the same kernel then accumulates them from HBM."""
        buf = self.state_buffer()
        if psi_local.data_ptr() != buf.data_ptr():
            buf.copy_(psi_local)
        self._hdl.barrier(channel=0)                 # every slice published
        shift = self.e_static + sum(d[q] * self.r_glob[q] for q in range(self.g))
        qs = [q for q in range(self.g) if c[q] != 0 or keep_partners]
        f = -1j if rhs else 1.0
        coefs = [f * shift] + [f * (c[q] if self.r_glob[q] == 0 else c[q].conjugate()) for q in qs]
        peers = [self.rank ^ (1 << (self.g - 1 - q)) for q in qs]
        copy = self._mode == "copy" or keep_partners
        if copy:
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(self.device)
                self._recv = [torch.empty_like(self._sym) for _ in range(self.g)]
                self._peer_bufs = [self._hdl.get_buffer(r, self._sym.shape, self._sym.dtype)
                                   for r in range(self.world)]
            # the pulls go chunk by chunk (all partners' chunk 0 first) so that the accumulate of a chunk can
            # run while the copy engines are still busy with the next ones
            n_loc = 1 << self.nl
            n_ch = max(1, min(self._pull_chunks, n_loc >> 20))
            ch = n_loc // n_ch
            if self._events is None or len(self._events) != n_ch:
                self._events = [torch.cuda.Event() for _ in range(n_ch)]
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                for c_ in range(n_ch):
                    a_, b_ = 2 * c_ * ch, 2 * (c_ + 1) * ch          # real view: (re, im) pairs
                    for k, r in enumerate(peers):
                        self._recv[k][:, a_:b_].copy_(self._peer_bufs[r][:, a_:b_])
                    self._events[c_].record(self._side)
            ptrs = [self._recv[k].data_ptr() for k in range(len(peers))]
        else:
            ptrs = [self._peer_ptrs[r] for r in peers]
        ops.configure(self.plan, self._prog)
        out = self.plan.hpsi(t, buf, rhs=rhs, out=out)
        if copy and n_ch > 1:
            for c_ in range(n_ch):
                main.wait_event(self._events[c_])
                off = self._amp_bytes * c_ * ch
                self.plan.sharded_accumulate_range(out.data_ptr() + off, buf.data_ptr() + off, 0.0,
                                                   [buf.data_ptr() + off] + [p_ + off for p_ in ptrs], coefs, ch)
        else:
            if copy:
                main.wait_stream(self._side)
            self.plan.sharded_accumulate(out, buf, 0.0, [buf.data_ptr()] + ptrs, coefs)
        self._hdl.barrier(channel=1)                 # partners are done reading this slice
        if keep_partners:
            n_loc = 1 << self.nl
            return out, [torch.view_as_complex(self._recv[k].view(1, n_loc, 2)) for k in range(self.g)]
        return out

    def local_slice(self, full: Tensor) -> Tensor:
        """This rank's (1, 2^(N-g)) slice of a full (1, 2^N) vector (tests)."""
        n_loc = 1 << self.nl
        return full[:, self.rank * n_loc:(self.rank + 1) * n_loc].contiguous()

    # -- DP5 evolution of the sharded register (configs[4]: full pulse sequence + gradient) ------
    # Same integrator, controller and discrete adjoint as the single-GPU engine
    # (csrc/engine.hpp forward_dp5 / adjoint_step; reference call backend.py:488-494), driven
    # from the host because every H.psi is one exchange step; the error norm is the only
    # reduction on the forward path (one scalar all-reduce per attempted step).
    def rhs(self, t: float, psi_local: Tensor, out: Optional[Tensor] = None) -> Tensor:
        return self.hpsi(t, psi_local, rhs=True, out=out)

    def _sum(self, x: float) -> float:
        if self.world == 1:
            return x
        v = torch.tensor([x], dtype=torch.float64, device=self.device)
        dist.all_reduce(v, group=self.group)
        return float(v.item())

    def _scaled_norm(self, x: Tensor, ref_abs: Tensor, atol: float, rtol: float) -> float:
        loc = ((x.abs() / (atol + rtol * ref_abs)) ** 2).sum().item()
        return math.sqrt(self._sum(loc) / float(1 << self.n))

    def _stage_input(self, y: Tensor, k: list, i: int, h: float, publish: bool = False,
                     out: Optional[Tensor] = None) -> Tensor:
        """Y_i = y + h sum_j beta_ij k_j in one pass; ``publish``: written straight into the
        peer-visible buffer (valid until the next generator application)."""
        ins, w = [y], [1.0]
        for j in range(i):
            if _BETA[i - 1][j] != 0.0:
                ins.append(k[j])
                w.append(h * _BETA[i - 1][j])
        if publish and self._hdl is not None:
            out = self.state_buffer()
        elif out is None:
            out = torch.empty_like(y)
        return self.plan.lincomb(out, ins, w)

    def _dp5_step(self, t: float, h: float, y: Tensor, k0: Tensor, upto: int = 6, ring=None):
        """k[0..upto], y_new of one step (stage 7's input is y_new: FSAL; ``upto=5`` stops before it).

        ``ring``: ``(slopes, y_next)`` = seven preallocated slope vectors (``slopes[0] is k0``) and the
        vector that receives y_new -- no allocation inside the step (the forward evolution; at 2^29
        amplitudes per GPU a fresh 8 GiB tensor per stage costs more than the stage itself)."""
        k = [k0]
        y_new = None
        for i in range(1, upto + 1):
            Y = self._stage_input(y, k, i, h, publish=i < 6, out=None if ring is None or i < 6 else ring[1])
            k.append(self.rhs(t + h * _ALPHA[i - 1], Y, out=None if ring is None else ring[0][i]))
            if i == 6:
                y_new = Y
        return k, y_new

    def evolve(self, psi0_local: Tensor, tsave: Sequence[float], atol: float = 1e-8, rtol: float = 1e-6,
               safety_factor: float = 0.9, min_factor: float = 0.2, max_factor: float = 5.0,
               max_steps: int = 100000, replay: Optional[Sequence[tuple]] = None):
        """Adaptive DP5 from ``psi0_local`` over ``tsave``.  Returns ``(states, steps)``:
        ``states`` (n_t, 1, 2^(N-g)) = this rank's slices at the save times, ``steps`` = the
        accepted steps ``(t, dt, interval, clipped)`` (identical on every rank; the tape of
        :meth:`evolve_backward`).  ``replay``: run exactly these steps instead of controlling."""
        ts = [float(x) for x in tsave]
        y = psi0_local.detach().clone()
        t = ts[0]
        # the integrator's whole working set, allocated once: y, y_next and seven slopes (+ the
        # peer-visible stage buffer and the g receive buffers of the exchange)
        slopes = [torch.empty_like(y) for _ in range(7)]
        y_next = torch.empty_like(y)
        k0 = self.rhs(t, y, out=slopes[0])
        steps = []
        # the saved states go straight into the result (no list + stack: two more vectors at the peak)
        states = torch.empty((len(ts),) + tuple(y.shape), dtype=y.dtype, device=y.device)
        if replay is None:
            d0 = self._scaled_norm(y, y.abs(), atol, rtol)
            d1 = self._scaled_norm(k0, y.abs(), atol, rtol)
            h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
            f1 = self.rhs(t + h0, y + h0 * k0)
            d2 = self._scaled_norm(f1 - k0, y.abs(), atol, rtol) / h0
            h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 6.0)
            dt = min(100 * h0, h1)
        else:
            dt, replay = 0.0, list(replay)
        error, pos = 1.0, 0
        ew = [_B5[j] - _B4[j] for j in range(7)]
        for kk, t_next in enumerate(ts):
            cache_dt, cache_err, n_att = dt, error, 0
            while t < t_next:
                if replay is None:
                    if error == 0.0:
                        dt = dt * max_factor
                    else:
                        fac = safety_factor * error ** (-0.2)
                        dt = dt * (max(1.0, min(max_factor, fac)) if error <= 1.0
                                   else min(0.9, max(min_factor, fac)))
                    clipped = t + dt >= t_next
                else:
                    _, dt, _, clipped = replay[pos]
                    pos += 1
                if clipped:
                    cache_dt, cache_err, dt = dt, error, t_next - t
                k, y_new = self._dp5_step(t, dt, y, k0, ring=(slopes, y_next))
                loc = self.plan.dp5_error_sumsq(k, [dt * e for e in ew], y, y_new, atol, rtol)
                error = math.sqrt(self._sum(float(loc[0])) / float(1 << self.n))
                if error != error:
                    raise RuntimeError("non-finite error norm in DP5 step")
                if replay is not None or error <= 1.0:
                    steps.append((t, dt, kk, bool(clipped)))
                    t = t_next if clipped else t + dt
                    y, y_next = y_new, y                 # FSAL: the last slope opens the next step
                    slopes[0], slopes[6] = slopes[6], slopes[0]
                    k0 = slopes[0]
                n_att += 1
                if n_att >= max_steps:
                    raise RuntimeError("max_steps reached")
            dt, error = cache_dt, cache_err
            states[kk].copy_(y)
        return states, steps

    def _vjp(self, t: float, Y: Tensor, kbar: Tensor, acc: dict) -> Tensor:
        """Reverse mode of k = -i H(t) Y on the sharded register: returns this rank's slice of
        ``(-i H)^dagger kbar`` and adds this rank's share of the parameter gradients to ``acc``
        (summed over ranks by the caller)."""
        g, n = self.g, self.n_samples
        # (-i H)^dagger kbar = i H kbar = -(-i H kbar)
        hk, partners = self.hpsi(t, kbar, keep_partners=True, rhs=True)
        hk = hk.neg_()
        i1 = max(int(min(math.floor(t / self.dt), n - 2)), 0)
        i2 = min(i1 + 1, n - 2)
        x = (t - i1 * self.dt) / self.dt
        # local qubits: the engine's own correlation kernels (C ABI pd_rhs_vjp)
        ops.configure(self.plan, self._prog)
        _, gl_det, gl_amp, _, _ = self.plan.rhs_vjp(t, Y, kbar, want_state=False, defer_pair=True)
        nk = len(self._keep_det)
        if gl_det is None:
            gl_det = torch.zeros(nk + len(self._extra_qubits), n, dtype=torch.float64)
        if gl_amp is None:
            gl_amp = torch.zeros(len(self._keep_amp), n, dtype=C128)
        for j, kidx in enumerate(self._keep_det):
            acc["det"][kidx] += gl_det[j]
        for j, (q_loc) in enumerate(self._extra_qubits):      # static detuning = interaction with
            tot = float(gl_det[nk + j].sum())                  # the occupied global qubits
            for q in range(g):
                acc["pair"][q, g + q_loc] += 0.5 * self.r_glob[q] * tot
        for j, kidx in enumerate(self._keep_amp):
            acc["amp"][kidx] += gl_amp[j]
        # energy shift of the global qubits: k has -i * shift * Y
        g_shift = torch.vdot(kbar.reshape(-1), Y.reshape(-1)).imag.item()
        for p_ in range(g):
            for q in range(p_ + 1, g):
                acc["pair"][p_, q] += self.r_glob[p_] * self.r_glob[q] * g_shift
        for kidx, m in enumerate(self.det_masks):
            w = 2.0 * g_shift * sum(self.r_glob[q] for q in range(g) if m >> q & 1)
            if w != 0.0:
                acc["det"][kidx, i1] += w * (1.0 - x)
                acc["det"][kidx, i2] += w * x
        # flips of the global qubits: with the partner's kbar slice in hand this rank evaluates
        # the PARTNER's drive correlation (its k has -i * coef * Y_mine); the sum over ranks is
        # what the caller reduces, so every contribution is counted once.
        for q in range(g):
            zb = -1j * torch.vdot(partners[q].reshape(-1), Y.reshape(-1)).item()
            gc = zb.conjugate() if (1 - self.r_glob[q]) == 0 else zb
            for kidx, m in enumerate(self.amp_masks):
                if m >> q & 1:
                    acc["amp"][kidx, i1] += gc * (1.0 - x)
                    acc["amp"][kidx, i2] += gc * x
        return hk

    def _slope_cache_steps(self, like: Tensor) -> int:
        """How many steps' slopes (6 slices each) the adjoint sweep may keep."""
        per_step = 6 * like.numel() * like.element_size()
        if like.device.type != "cuda":
            return (1 << 28) // per_step
        free, _ = torch.cuda.mem_get_info(like.device)
        return int(free // 3 // per_step)

    def evolve_backward(self, states: Tensor, grad_states: Tensor, steps: Sequence[tuple]) -> dict:
        """Discrete adjoint of :meth:`evolve` (the recorded step sequence, recomputed stage by
        stage as in csrc/engine.hpp adjoint_step).  ``grad_states``: cotangents on this rank's
        ``states``.  Returns ``{"det", "amp", "pair", "state0"}``: gradients w.r.t. the
        coefficient samples and ``pair_u`` (complete, identical on every rank) and this rank's
        slice of the gradient w.r.t. the initial state."""
        acc = {"det": torch.zeros_like(self.det_values), "amp": torch.zeros_like(self.amp_values),
               "pair": torch.zeros(self.n, self.n, dtype=torch.float64)}
        n_t = int(states.shape[0])
        ops.configure(self.plan, self._prog)
        self.plan.pair_gradient_flush()              # start from an empty accumulator
        lam = grad_states[n_t - 1].detach().clone()
        hi = len(steps)
        for kk in range(n_t - 1, 0, -1):
            lo = hi
            while lo > 0 and steps[lo - 1][2] == kk:
                lo -= 1
            # step-start states of the interval, recomputed from the saved state; the slopes of
            # that pass are kept while they fit a third of the free memory (engine.hpp does the same)
            ys, ks = [states[kk - 1]], []
            room = self._slope_cache_steps(states[kk - 1])
            k0 = None
            for (t, h, _, _) in steps[lo:hi - 1]:
                k, y1 = self._dp5_step(t, h, ys[-1], k0 if k0 is not None else self.rhs(t, ys[-1]))
                ys.append(y1)
                ks.append(k[:6] if len(ks) < room else None)
                k0 = k[6]                                   # FSAL: first slope of the next step
            for s_idx in range(hi - 1, lo - 1, -1):
                t, h, _, _ = steps[s_idx]
                y_n = ys[s_idx - lo]
                k = ks.pop() if len(ks) > s_idx - lo else None
                if k is None:
                    k, _ = self._dp5_step(t, h, y_n, self.rhs(t, y_n), upto=5)
                yb = [None] * 6
                for i in range(5, -1, -1):
                    ins, w = ([lam], [h * _B5[i]]) if _B5[i] != 0.0 else ([], [])
                    for j in range(i + 1, 6):
                        if _BETA[j - 1][i] != 0.0:
                            ins.append(yb[j])
                            w.append(h * _BETA[j - 1][i])
                    kbar = self.plan.lincomb(self.state_buffer() if self._hdl is not None
                                             else torch.empty_like(lam), ins, w)
                    ts_i = t + h * (0.0 if i == 0 else _ALPHA[i - 1])
                    yb[i] = self._vjp(ts_i, self._stage_input(y_n, k, i, h), kbar, acc)
                self.plan.lincomb(lam, [lam] + yb, [1.0] * 7)
            hi = lo
            lam.add_(grad_states[kk - 1])
        ops.configure(self.plan, self._prog)
        acc["pair"][self.g:, self.g:] += self.plan.pair_gradient_flush()
        for key in ("det", "amp", "pair") if self.world > 1 else ():
            v = acc[key].to(self.device)
            if v.is_complex():
                v = torch.view_as_real(v).contiguous()
                dist.all_reduce(v, group=self.group)
                acc[key] = torch.view_as_complex(v).cpu()
            else:
                dist.all_reduce(v, group=self.group)
                acc[key] = v.cpu()
        acc["state0"] = lam
        return acc


# Dormand-Prince 5(4) tableau (the one csrc/pd_common.hpp holds for the device paths)
_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_B5 = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
_B4 = [5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40]
