"""Stochastic-noise runs (reference backend.py:568-611, hamiltonian.py:179-219, 270-286):
``runs`` random Hamiltonians (Doppler offsets, per-pulse amplitude fluctuation x beam profile, badly prepared
atoms) evolved as ONE batch of parameter sets, bitstring counts sampled on the device."""
import pytest
import torch

from helpers import C6_60
from oracle.ref_emulator import RefEmulator
import pulser_diff_b200 as pdb
from pulser_diff_b200 import _cabi, ops
from pulser_diff_b200.samples import PulseBuilder, SequenceSamples
from pulser_diff_b200.utils import total_magnetization_diag

T = 200


def _emulator(device, cfg, n=3, local=False):
    one = torch.ones(T, dtype=torch.float64)
    b = PulseBuilder().add(4.0 * one, 1.0 * one, 0.0).add(2.0 * one, -1.0 * one, 0.3)
    chans = [b.build()]
    if local:
        chans.append(PulseBuilder("Local", ["q1"]).delay(100).add(1.5 * one[:150], 0.5 * one[:150], 0.2).delay(150).build())
    reg = {f"q{i}": torch.tensor([7.0 * i, 0.5 * i], dtype=torch.float64) for i in range(n)}
    em = pdb.TorchEmulator(SequenceSamples(chans), reg, pdb.DeviceSpec(C6_60), 0.2, cfg, "Minimal",
                           torch_device=device)
    em.noise_generator = torch.Generator().manual_seed(11)
    return em, reg


def test_noise_realisations_against_oracle_and_sequential_runs(engine_device):
    """Every run of the batch equals (a) the same realisation evolved on its own through ``ops.evolve`` and
    (b) the ORACLE fed with the realisation's per-qubit sample arrays (Doppler + amplitude draws)."""
    cfg = pdb.SimConfig(noise=("doppler", "amplitude"), amp_sigma=0.07, laser_waist=40.0, temperature=300.0,
                        runs=5, samples_per_run=10)
    em, reg = _emulator(engine_device, cfg, local=True)
    tight = dict(atol=1e-13, rtol=1e-12)       # the pulse edges make two default controllers part ways (~rtol)
    res = em.run(**tight)
    assert isinstance(res, pdb.simresults.NoisyResults) and res.n_measures == 50
    assert all(sum(c.values()) == 50 for c in res.results)
    coords = torch.stack(list(reg.values()))
    psi0 = em.initial_state.to(engine_device).transpose(0, 1).contiguous()
    assert len(em._last_noisy_states) == 5
    seen = set()
    for d, st in em._last_noisy_states:
        seen.add(round(float(d["doppler"][0]), 12))
        single = ops.evolve(psi0, em.evaluation_times, d["det_values"], d["amp_values"], d["pair_u"], n_qubits=3,
                            kind=_cabi.PD_KET, dt=d["dt"], det_masks=d["det_masks"], amp_masks=d["amp_masks"],
                            options=_cabi.Options(**tight))
        assert (single.permute(0, 2, 1) - st).abs().max() < 1e-12
        smp = {"Local": {q: {k: d["samples"][k][q] for k in ("amp", "det", "phase")} for q in range(3)}}
        ref = RefEmulator(coords, C6_60, smp, rate=0.2, evaluation_times="Minimal")
        want = ref.run(**tight).states
        assert (st.cpu() - want).abs().max() < 1e-9
    assert len(seen) == 5                      # five different Doppler draws


def test_spam_only_runs_group_preparation_patterns(engine_device):
    """SPAM without resampled noise: distinct preparation patterns are run once each with their multiplicity
    (reference backend.py:546-566); a badly prepared atom stays in |g> and out of the interaction."""
    cfg = pdb.SimConfig(noise=("SPAM",), eta=0.4, epsilon=0.0, epsilon_prime=0.0, runs=30, samples_per_run=20)
    em, _ = _emulator(engine_device, cfg)
    res = em.run()
    reps = [r for _, r in em._last_noise_draws]
    assert sum(reps) == 30 and len(reps) <= 8
    assert sum(res.results[-1].values()) == 600
    for d, st in em._last_noisy_states:
        bad = d["bad_atoms"]
        probs = (st[-1].abs() ** 2).sum(dim=1).cpu()          # index bit 0 = |r>, qubit 0 = most significant
        s = torch.arange(8)
        for q in range(3):
            if bad[q]:
                assert probs[((s >> (2 - q)) & 1) == 0].sum() < 1e-14      # never excited
                assert not d["pair_u"][q].any() and not d["pair_u"][:, q].any()


def test_noiseless_limit_reproduces_coherent_statistics(engine_device):
    """With every noise source switched to zero strength the averaged counts follow the coherent run's
    distribution (5-sigma binomial bound), and detection errors move <sum Z> of the initial state by the
    expected 2 N epsilon."""
    quiet = pdb.SimConfig(noise=("doppler", "SPAM"), temperature=0.0, eta=0.0, epsilon=0.0, epsilon_prime=0.0,
                          runs=1, samples_per_run=1)
    assert not quiet.needs_resampling or quiet.temperature == 0.0
    cfg = pdb.SimConfig(noise=("doppler", "amplitude"), temperature=0.0, amp_sigma=1e-300, runs=4, samples_per_run=2500)
    em, _ = _emulator(engine_device, cfg)
    noisy = em.run()
    em2, _ = _emulator(engine_device, pdb.SimConfig())
    clean = em2.run()
    p = clean._weights(len(clean) - 1).cpu()
    f = noisy.probabilities(float(em.evaluation_times[-1]))
    sigma = torch.sqrt(p * (1 - p) / noisy.n_measures)
    assert ((f - p).abs() <= 5 * sigma + 1e-4).all()
    cfg3 = pdb.SimConfig(noise=("SPAM", "doppler"), temperature=0.0, eta=0.0, epsilon=0.05, epsilon_prime=0.0,
                         runs=2, samples_per_run=20000)
    em3, _ = _emulator(engine_device, cfg3)
    z0 = em3.run().expect([total_magnetization_diag(3)])[0][0]
    assert abs(z0.item() - (-3 + 2 * 3 * 0.05)) < 0.02


def test_noisy_run_is_reproducible_with_a_seeded_generator(engine_device):
    cfg = pdb.SimConfig(noise=("doppler", "amplitude", "SPAM"), amp_sigma=0.05, eta=0.1, runs=6, samples_per_run=5)
    outs = []
    for _ in range(2):
        em, _ = _emulator(engine_device, cfg)
        em.run()
        outs.append(torch.stack([d["doppler"] for d, _ in em._last_noise_draws]))
    assert torch.equal(outs[0], outs[1])


def test_noise_runs_in_complex64_mode(engine_device):
    """The batch of noise realisations honours ``dtype="complex64"``: complex64 states out, the same realisations
    (same generator seed) within 1e-6 of the complex128 run -- N = 3 is served by the complex128 kernels behind a
    cast (ops.compute_dtype), so only the storage precision of the returned states differs."""
    cfg = pdb.SimConfig(noise=("doppler",), temperature=300.0, runs=4, samples_per_run=5)
    em, _ = _emulator(engine_device, cfg)
    em.run()
    ref = [st for _, st in em._last_noisy_states]
    em64, _ = _emulator(engine_device, cfg)
    res = em64.run(dtype="complex64")
    assert isinstance(res, pdb.simresults.NoisyResults)
    got = [st for _, st in em64._last_noisy_states]
    assert len(got) == len(ref) == 4
    for a, b in zip(got, ref):
        assert a.dtype == torch.complex64 and b.dtype == torch.complex128
        assert (a.to(torch.complex128) - b).abs().max() < 1e-6
