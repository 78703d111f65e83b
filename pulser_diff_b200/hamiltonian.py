"""Hamiltonian assembly for the B200 path: structure instead of sparse matrices.

Same constructor, attributes and conventions as reference ``pulser_diff/hamiltonian.py`` for
the ising / ground-rydberg branch (the only one in ``north_star``'s H):

* sampling times / sub-sampling                  reference hamiltonian.py:67-91
* ``basis`` / ``op_matrix`` / ``build_operator``  reference hamiltonian.py:221-268, 288-318
* van-der-Waals couplings + ``_dist_dict``        reference hamiltonian.py:333-344, 385-404
* coefficient arrays, zero terms dropped          reference hamiltonian.py:406-454
* collapse operators                              reference hamiltonian.py:98-143
* the ``H_t`` closure                             reference hamiltonian.py:499-548

What differs is WHAT is produced: no 2^N x 2^N operator is ever built.  ``_hamiltonian`` is a
:class:`StructuredHamiltonian` that (a) still behaves as ``H(t) -> sparse COO`` for small
registers, so code written against the reference keeps working, and (b) exposes the term
structure (qubit masks + coefficient tensors + pair couplings) that crosses the C ABI.
XY mode, SLM masks, the digital / all bases and stochastic noise resampling are out of scope
(SURVEY.md 2, row 2).
"""
from __future__ import annotations

import itertools
from math import floor
from typing import Any, Optional, Sequence, Union

import torch
from torch import Tensor

from .samples import SequenceSamples
from .simconfig import SimConfig
from .utils import XMAT, YMAT, ZMAT, basis_state, kron

C128 = torch.complex128


class CollapseOperators(list):
    """The list ``Hamiltonian._collapse_ops``: one entry per (local operator, qubit), as in the
    reference (hamiltonian.py:139-143), but entries are ``(op2x2, qubit)`` descriptors; the
    dense 2^N x 2^N matrix is only materialised on request (``dense(i)``)."""

    def __init__(self, local_ops: Sequence[Tensor], n_qubits: int) -> None:
        super().__init__((op, q) for op in local_ops for q in range(n_qubits))
        self.local_ops = [op.to(C128) for op in local_ops]
        self.n_qubits = n_qubits

    def stacked(self) -> Optional[Tensor]:
        return torch.stack(self.local_ops) if self.local_ops else None

    def dense(self, i: int) -> Tensor:
        op, q = self[i]
        facs = [torch.eye(2, dtype=C128) for _ in range(self.n_qubits)]
        facs[q] = op.to(C128)
        return kron(*facs)


class StructuredHamiltonian:
    """H(t) of a Rydberg register as structure.

    Attributes (all torch, autograd-connected to pulse parameters / coordinates):
        pair_u       (N, N) upper-triangular C6 / r_ij^6  (net coefficient of n_i n_j)
        det_terms    list of (qubit mask, coefficient tensor (n_samples,))   [-0.5 * det]
        amp_terms    list of (qubit mask, coefficient tensor (n_samples,))   [0.5 * amp * e^{-i phase}]
        dt, n_samples  interpolation grid of the closure (hamiltonian.py:523-524)
    """

    def __init__(self, n_qubits: int, pair_u: Tensor, dt: float, n_samples: int,
                 det_terms: list, amp_terms: list, device: torch.device) -> None:
        self.n_qubits, self.pair_u, self.dt, self.n_samples = n_qubits, pair_u, dt, n_samples
        self.det_terms, self.amp_terms = det_terms, amp_terms
        self.device = torch.device(device)

    # -- what crosses the C ABI --------------------------------------------------------------
    def masks_and_values(self):
        dm = [m for m, _ in self.det_terms]
        am = [m for m, _ in self.amp_terms]
        dv = (torch.stack([v.real if v.is_complex() else v for _, v in self.det_terms])
              if dm else torch.zeros(0, self.n_samples, dtype=torch.float64))
        av = (torch.stack([v.to(C128) for _, v in self.amp_terms])
              if am else torch.zeros(0, self.n_samples, dtype=C128))
        return dm, dv, am, av

    # -- compatibility: H(t) as a sparse COO matrix (small registers only) --------------------
    def coefficient(self, values: Tensor, t) -> Tensor:
        """Linear interpolation with the reference's index rule (hamiltonian.py:532-542)."""
        if not isinstance(t, Tensor):
            t = torch.tensor(t, dtype=torch.float64)
        i1 = max(int(min(floor(float(t) / self.dt), self.n_samples - 2)), 0)
        i2 = min(i1 + 1, self.n_samples - 2)
        return values[i1] + (values[i2] - values[i1]) * (t - i1 * self.dt) / self.dt

    def __call__(self, t) -> Tensor:
        n = self.n_qubits
        if n > 14:
            raise ValueError("materialising H(t) is limited to 14 qubits; use the structured path")
        s = torch.arange(2 ** n)
        r = torch.stack([1 - ((s >> (n - 1 - q)) & 1) for q in range(n)]).to(torch.float64)
        diag = torch.zeros(2 ** n, dtype=C128)
        for i, j in itertools.combinations(range(n), 2):
            diag = diag + self.pair_u[i, j].to("cpu") * r[i] * r[j]
        for mask, v in self.det_terms:
            c = self.coefficient(v.to("cpu"), t)
            for q in range(n):
                if mask >> q & 1:
                    diag = diag + 2 * c * r[q]
        idx, val = [torch.stack([s, s])], [diag]
        for mask, v in self.amp_terms:
            c = self.coefficient(v.to("cpu"), t).to(C128)
            for q in range(n):
                if mask >> q & 1:
                    m = 1 << (n - 1 - q)
                    g_rows = s[(s & m) != 0]          # bit 1 = |g>:  <g| H |r> = c
                    idx += [torch.stack([g_rows, g_rows ^ m]), torch.stack([g_rows ^ m, g_rows])]
                    val += [c.expand(g_rows.numel()), c.conj().expand(g_rows.numel())]
        return torch.sparse_coo_tensor(torch.cat(idx, 1), torch.cat(val), (2 ** n, 2 ** n)).coalesce()


class Hamiltonian:
    r"""Generates the structured Hamiltonian from a sampled sequence and noise.

    Args:
        samples_obj: sampled sequence whose channels have the same duration.
        qdict: qubit id -> coordinates (tensor, may require grad).
        device: object with ``interaction_coeff`` (C6, rad/us * um^6).
        sampling_rate: fraction of samples kept (0 < rate <= 1).
        config: noise configuration.
        torch_device: CUDA device the evolution runs on.
    """

    def __init__(self, samples_obj: SequenceSamples, qdict: dict, device: Any, sampling_rate: float,
                 config: SimConfig, torch_device: Union[str, torch.device] = "cuda") -> None:
        self.samples_obj = samples_obj
        self._qdict = {k: torch.as_tensor(v, dtype=torch.float64) if not isinstance(v, Tensor)
                       else v.to(torch.float64) for k, v in qdict.items()}
        self._device = device
        self._sampling_rate = sampling_rate
        self.torch_device = torch.device(torch_device)
        self._dist_dict: dict[str, Tensor] = {}
        self._interaction = "ising"
        self._size = len(self._qdict)
        self._qid_index = {qid: i for i, qid in enumerate(self._qdict)}
        self._duration = self.samples_obj.max_duration
        self.sampling_times = self._adapt_to_sampling_rate(
            torch.arange(self._duration, dtype=torch.double) / 1000)
        self._collapse_ops: CollapseOperators = CollapseOperators([], self._size)
        self.set_config(config)

    def _adapt_to_sampling_rate(self, full_array: Tensor) -> Tensor:
        """Keep ``int(rate * duration)`` samples at integer positions spread evenly over the array,
        end points included (sub-sampling rule of reference hamiltonian.py:83-91)."""
        keep = int(self._sampling_rate * self._duration)
        return full_array[torch.linspace(0, len(full_array) - 1, keep, dtype=torch.int)]

    @property
    def config(self) -> SimConfig:
        return self._config

    def set_config(self, cfg: SimConfig) -> None:
        if not isinstance(cfg, SimConfig):
            raise ValueError(f"Object {cfg} is not a valid `NoiseModel`.")
        if not hasattr(self, "basis_name"):
            self._build_basis_and_op_matrices()
        self._build_collapse_operators(cfg)
        self._config = cfg
        self._construct_hamiltonian()

    def _build_basis_and_op_matrices(self) -> None:
        self.basis_name = "ground-rydberg"
        self.dim = 2
        self.basis = {b: basis_state(2, i) for i, b in enumerate(["r", "g"])}
        self.op_matrix = {"I": torch.eye(2).to_sparse()}
        for proj in ["gr", "rr", "gg"]:
            self.op_matrix["sigma_" + proj] = (self.basis[proj[0]] * self.basis[proj[1]].mH).to_sparse()

    def _build_collapse_operators(self, config: SimConfig) -> None:
        local = []
        if "dephasing" in config.noise_types:
            local.append(torch.sqrt(torch.as_tensor(config.dephasing_rate) / 2) * ZMAT)
        if "relaxation" in config.noise_types:
            local.append(torch.sqrt(torch.as_tensor(config.relaxation_rate))
                         * self.op_matrix["sigma_gr"].to_dense().to(C128))
        if "depolarizing" in config.noise_types:
            coeff = torch.sqrt(torch.as_tensor(config.depolarizing_rate) / 4)
            local += [coeff * XMAT, coeff * YMAT, coeff * ZMAT]
        if "eff_noise" in config.noise_types:
            for rate, op in zip(config.eff_noise_rates, config.eff_noise_opers):
                local.append(torch.sqrt(torch.as_tensor(rate)) * torch.as_tensor(op).to(C128))
        self._collapse_ops = CollapseOperators(local, self._size)

    def build_operator(self, operations: Union[list, tuple]) -> Tensor:
        """Tensor-product operator from ``[(op, qubit ids), ...]`` (identity elsewhere); an entry
        ``(op, "global")`` yields the sum of ``op`` over every qubit instead.  ``op`` is a 2x2
        tensor or a key of ``op_matrix``.  Same contract and exceptions as reference
        hamiltonian.py:221-268."""
        entries = operations if isinstance(operations, list) else [operations]
        placed: dict[int, Tensor] = {}
        for op, where in entries:
            if where == "global":
                terms = [self.build_operator([(op, [qid])]) for qid in self._qdict]
                return sum(terms[1:], terms[0])
            ids = list(where)
            if len(set(ids)) != len(ids):
                raise ValueError("Duplicate atom ids in argument list.")
            unknown = set(ids) - self._qdict.keys()
            if unknown:
                raise ValueError(f"Invalid qubit names: {unknown}")
            if isinstance(op, str):
                if op not in self.op_matrix:
                    raise ValueError(f"{op} is not a valid operator")
                op = self.op_matrix[op]
            placed.update({self._qid_index[q]: op for q in ids})
        return kron(*[placed.get(k, self.op_matrix["I"]) for k in range(self._size)])

    def _extract_samples(self) -> None:
        samples = self.samples_obj.to_nested_dict(list(self._qdict), all_local=False)
        cfg = self._config
        if "amplitude" in cfg.noise_types and cfg.laser_waist is not None:
            # deterministic Gaussian-beam amplitude loss on global pulses
            # (reference hamiltonian.py:196-204 with noise_amp_base = 1 since amp_sigma = 0)
            samples = self.samples_obj.to_nested_dict(list(self._qdict), all_local=True)
            w0 = torch.as_tensor(cfg.laser_waist)
            glob = [c for c in self.samples_obj.channels if c.addressing == "Global"]
            for basis, per in samples["Local"].items():
                for qid in per:
                    frac = torch.exp(-((torch.linalg.norm(self._qdict[qid]) / w0) ** 2))
                    g_amp = sum(c.amp for c in glob if c.basis == basis)
                    per[qid]["amp"] = per[qid]["amp"] - g_amp + g_amp * frac
        self.samples = samples

    def _construct_hamiltonian(self, update: bool = True) -> None:
        self._extract_samples()
        n = self._size
        qids = list(self._qdict)
        pair_u = torch.zeros(n, n, dtype=torch.float64)
        for q1, q2 in itertools.combinations(qids, 2):
            dist = torch.linalg.norm(self._qdict[q1] - self._qdict[q2])
            self._dist_dict[f"{q1}-{q2}"] = dist
        self._pair_index = {f"{q1}-{q2}": (self._qid_index[q1], self._qid_index[q2])
                            for q1, q2 in itertools.combinations(qids, 2)}
        det_terms, amp_terms = [], []

        def add_terms(s: dict, mask: int) -> None:
            coeffs = [0.5 * s["amp"] * torch.exp(-1j * s["phase"]), -0.5 * s["det"]]
            for coeff, dst in zip(coeffs, (amp_terms, det_terms)):
                if torch.any(coeff != 0):
                    dst.append((mask, self._adapt_to_sampling_rate(coeff)))

        for basis, s in self.samples["Global"].items():
            add_terms(s, (1 << n) - 1)
        for basis, per in self.samples["Local"].items():
            for qid, s in per.items():
                add_terms(s, 1 << self._qid_index[qid])
        self._det_terms, self._amp_terms = det_terms, amp_terms
        self._dt = 0.001 / self._sampling_rate
        self._n_samples = int(self._sampling_rate * self._duration)
        self._hamiltonian = StructuredHamiltonian(n, self._pair_couplings(), self._dt,
                                                  self._n_samples, det_terms, amp_terms,
                                                  self.torch_device)

    def _pair_couplings(self) -> Tensor:
        """(N, N) upper triangle of C6 / r_ij^6 from the CURRENT ``_dist_dict`` tensors, so that
        ``dist_grad`` (reference backend.py:456-460) differentiates the same handles."""
        n = self._size
        out = torch.zeros(n, n, dtype=torch.float64)
        if not self._pair_index:
            return out
        keys = list(self._pair_index)
        dists = torch.stack([self._dist_dict[k] for k in keys])
        # 2 * (0.5 * C6 / r^6): hamiltonian.py:343 and the factor 2 of :536
        vals = 2 * (0.5 * self._device.interaction_coeff / dists ** 6)
        ii = torch.tensor([self._pair_index[k][0] for k in keys])
        jj = torch.tensor([self._pair_index[k][1] for k in keys])
        return out.index_put((ii, jj), vals.to(torch.float64))

    # ---- stochastic noise (reference hamiltonian.py:179-219, 270-286; backend.py:568-611) ---------
    def noise_realisation(self, generator: Optional[torch.Generator] = None,
                          bad_atoms: Optional[Tensor] = None, resample: bool = True) -> dict:
        """One random Hamiltonian of the noisy sequence, as structure: per-qubit coefficient tables
        (``det_values`` (N, n_samples) float64, ``amp_values`` (N, n_samples) complex128 with masks
        ``1 << q``), the pair couplings with badly prepared atoms removed, and the draws themselves.

        Follows the reference's order of operations: every channel is expanded to per-qubit samples
        (``to_nested_dict(all_local=True)``), then, channel by channel and pulse slot by pulse slot, the
        Doppler offset of each targeted atom is added to its detuning inside the slot and -- for Global
        channels -- the amplitude inside the slot is scaled by one Gaussian draw per slot times the beam
        profile ``exp(-(r / w0)^2)``; samples of badly prepared atoms are zeroed and they drop out of
        the interaction.  ``resample=False`` keeps the deterministic parts only (the reference's
        ``update=False`` runs over SPAM configurations)."""
        from .simconfig import doppler_sigma
        cfg, n, D = self._config, self._size, self._duration
        qids = list(self._qdict)
        noise = set(cfg.noise_types)

        def normal(mean: float, std: float, size: int) -> Tensor:
            if std == 0.0:
                return torch.full((size,), float(mean), dtype=torch.float64)
            return mean + std * torch.randn(size, dtype=torch.float64, generator=generator)

        if bad_atoms is None:
            bad_atoms = (torch.rand(n, dtype=torch.float64, generator=generator) < cfg.eta
                         if ("SPAM" in noise and cfg.eta > 0 and resample) else torch.zeros(n, dtype=torch.bool))
        bad_atoms = torch.as_tensor(bad_atoms, dtype=torch.bool)
        doppler = (normal(0.0, doppler_sigma(cfg.temperature * 1e-6), n) if ("doppler" in noise and resample)
                   else torch.zeros(n, dtype=torch.float64))
        arr = {k: torch.zeros(n, D, dtype=torch.float64) for k in ("amp", "det", "phase")}
        plan = []
        for c in self.samples_obj.channels:
            everyone = qids if c.addressing == "Global" else list(c.targets or [])
            from .samples import Slot
            for slot in (c.slots or [Slot(0, c.duration, set(everyone))]):
                rows = [self._qid_index[q] for q in (slot.targets or everyone)]
                for k in ("amp", "det", "phase"):
                    arr[k][rows, slot.ti:slot.tf] += getattr(c, k).detach()[slot.ti:slot.tf]
                plan.append((c.addressing == "Global", slot.ti, slot.tf, rows))
        amp_draws = []
        for is_global, ti, tf, rows in plan:
            base = max(0.0, float(normal(1.0, cfg.amp_sigma if resample else 0.0, 1)))
            amp_draws.append(base)
            if "doppler" in noise:
                arr["det"][rows, ti:tf] += doppler[rows, None]
            if "amplitude" in noise and is_global:
                frac = torch.ones(len(rows), dtype=torch.float64)
                if cfg.laser_waist is not None:
                    r = torch.stack([torch.linalg.norm(self._qdict[qids[i]].detach()) for i in rows])
                    frac = torch.exp(-((r / float(cfg.laser_waist)) ** 2))
                arr["amp"][rows, ti:tf] *= (base * frac)[:, None]
        for k in arr:
            arr[k][bad_atoms] = 0.0
        keep = torch.linspace(0, D - 1, int(self._sampling_rate * D), dtype=torch.int).long()
        pair_u = self._pair_couplings().detach().clone()
        pair_u[bad_atoms, :] = 0.0
        pair_u[:, bad_atoms] = 0.0
        if n - int(bad_atoms.sum()) <= 1:
            pair_u.zero_()
        return {"det_masks": [1 << q for q in range(n)], "amp_masks": [1 << q for q in range(n)],
                "det_values": (-0.5 * arr["det"])[:, keep],
                "amp_values": (0.5 * arr["amp"] * torch.exp(-1j * arr["phase"]))[:, keep],
                "pair_u": pair_u, "bad_atoms": bad_atoms, "doppler": doppler, "amp_draws": amp_draws,
                "samples": arr,                      # the (N, T+1) per-qubit arrays the tables were cut from
                "dt": self._dt, "n_samples": self._n_samples}

    def refresh_couplings(self) -> None:
        self._hamiltonian.pair_u = self._pair_couplings()
