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
    """

    def __init__(self, n_qubits: int, pair_u: Tensor, dt: float, det_masks: Sequence[int],
                 det_values: Tensor, amp_masks: Sequence[int], amp_values: Tensor,
                 device: torch.device, group=None, peer_memory: bool | str = False) -> None:
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
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
        self.plan = ops.get_plan(self.nl, 1, _cabi.PD_KET, self.device)
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
        dm_all = [dm[k] for k in keep] + extra_m
        dv_all = torch.stack([dv[k] for k in keep] + extra_v) if dm_all else torch.zeros(0, self.n_samples)
        keep_a = [k for k, m in enumerate(am) if m]
        am_all = [am[k] for k in keep_a]
        av_all = torch.stack([av[k] for k in keep_a]) if am_all else torch.zeros(0, self.n_samples, dtype=C128)
        self._prog = ops.make_program(self.nl, _cabi.PD_KET, self.dt, dm_all, dv_all, am_all, av_all,
                                      u[loc, loc].contiguous(), None)
        # energy shift from global-global interaction (static) -- detuning part is time dependent
        self.e_static = sum(float(u[p, q]) * self.r_glob[p] * self.r_glob[q]
                            for p in range(g) for q in range(p + 1, g))
        self._sym = self._hdl = self._side = None
        self._mode = "read" if peer_memory == "read" else "copy"
        if peer_memory:
            if self.device.type != "cuda":
                raise ValueError("peer_memory needs CUDA ranks")
            import torch.distributed._symmetric_memory as symm_mem
            # float64 view of the complex slice: (re, im) pairs, the layout the kernels read
            self._sym = symm_mem.empty((1, 2 << self.nl), dtype=torch.float64, device=self.device)
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

    def hpsi(self, t: float, psi_local: Tensor) -> Tensor:
        """``(H(t) psi)`` restricted to this rank's slice.  ``psi_local``: (1, 2^(N-g))."""
        d, c = self._global_coefficients(t)
        if self._hdl is not None:
            return self._hpsi_peer(t, psi_local, d, c)
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
        out = self.plan.hpsi(t, psi_local)
        shift = self.e_static + sum(d[q] * self.r_glob[q] for q in range(self.g))
        if shift != 0.0:
            out.add_(psi_local, alpha=shift)
        # 3. global flips: this rank's bit for qubit q is 1 (ground) -> coefficient c, else conj(c)
        for r in reqs:
            r.wait()
        for q in range(self.g):
            if c[q] == 0:
                continue
            coef = c[q] if self.r_glob[q] == 0 else c[q].conjugate()
            out.add_(recv[q], alpha=coef)
        return out

    def _hpsi_peer(self, t: float, psi_local: Tensor, d, c) -> Tensor:
        """Peer-memory variants.  "read": one kernel accumulates the partner slices in place over
        NVLink.  "copy": the copy engines pull the partner slices into local buffers on a second
        stream while the local kernels run; the same kernel then accumulates them from HBM."""
        buf = self.state_buffer()
        if psi_local.data_ptr() != buf.data_ptr():
            buf.copy_(psi_local)
        self._hdl.barrier(channel=0)                 # every slice published
        shift = self.e_static + sum(d[q] * self.r_glob[q] for q in range(self.g))
        qs = [q for q in range(self.g) if c[q] != 0]
        coefs = [c[q] if self.r_glob[q] == 0 else c[q].conjugate() for q in qs]
        peers = [self.rank ^ (1 << (self.g - 1 - q)) for q in qs]
        if self._mode == "copy":
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(self.device)
                self._recv = [torch.empty_like(self._sym) for _ in range(self.g)]
                self._peer_bufs = [self._hdl.get_buffer(r, self._sym.shape, self._sym.dtype)
                                   for r in range(self.world)]
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                for k, r in enumerate(peers):
                    self._recv[k].copy_(self._peer_bufs[r])
            ptrs = [self._recv[k].data_ptr() for k in range(len(peers))]
        else:
            ptrs = [self._peer_ptrs[r] for r in peers]
        ops.configure(self.plan, self._prog)
        out = self.plan.hpsi(t, buf)
        if self._mode == "copy":
            main.wait_stream(self._side)
        self.plan.sharded_accumulate(out, buf, shift, ptrs, coefs)
        self._hdl.barrier(channel=1)                 # partners are done reading this slice
        return out

    def local_slice(self, full: Tensor) -> Tensor:
        """This rank's (1, 2^(N-g)) slice of a full (1, 2^N) vector (tests)."""
        n_loc = 1 << self.nl
        return full[:, self.rank * n_loc:(self.rank + 1) * n_loc].contiguous()
