"""Duration-mode envelopes built where the parameters live (reference
waveform_funcs.py:9-27, model.py:324-368)."""
import pytest
import torch

from pulser_diff_b200.samples import duration_mode_samples


def _params(dev):
    f = lambda v: torch.tensor(v, dtype=torch.float64, device=dev, requires_grad=True)
    return [f(0.4), f(0.35), f(0.25)], [f(2.0), f(5.0), f(3.0)], [f(0.5), f(0.0), f(1.0)], [f(0.0), f(0.3), f(0.0)]


def test_duration_mode_samples_cpu_gradients_flow():
    d, a, de, ph = _params("cpu")
    s = duration_mode_samples(d, a, de, ph)
    assert s["amp"].shape == (1005,)
    g = torch.autograd.grad(s["amp"].sum() + s["det"].sum(), d + a + de)
    assert all(torch.isfinite(x) for x in g)


@pytest.mark.gpu
def test_duration_mode_samples_on_device(cuda_device):
    """Same samples and gradients on the GPU as on the CPU; nothing lands on the host."""
    dc, ac, dec, phc = _params("cpu")
    dg, ag, deg, phg = _params(cuda_device)
    sc = duration_mode_samples(dc, ac, dec, phc)
    sg = duration_mode_samples(dg, ag, deg, phg)
    for k in ("amp", "det", "phase"):
        assert sg[k].device.type == "cuda"
        assert (sg[k].cpu() - sc[k]).abs().max() < 1e-13
    gc = torch.autograd.grad(sc["amp"].sum() + (sc["det"] ** 2).sum(), dc + ac + dec)
    gg = torch.autograd.grad(sg["amp"].sum() + (sg["det"] ** 2).sum(), dg + ag + deg)
    for x, y in zip(gc, gg):
        assert abs(x.item() - y.item()) < 1e-10 * max(1.0, abs(x.item()))
