"""GPU-only checks at sizes the oracle cannot reach: size-independent properties.

* unitarity (norm conservation) and linearity of the evolution,
* H Hermitian: <a|H b> == conj(<b|H a>),
* gather kernels vs tiled kernels produce the same vectors,
* forward + adjoint gradient against a central finite difference of the same engine.
"""
import pytest
import torch

import pulser_diff_b200 as pdb
from pulser_diff_b200 import _cabi, ops

pytestmark = pytest.mark.gpu


def _program(n, T=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    dv = (torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5) * 4
    av = torch.complex(torch.rand(1, T, dtype=torch.float64, generator=g) * 3,
                       torch.rand(1, T, dtype=torch.float64, generator=g) - 0.5)
    x = torch.arange(n, dtype=torch.float64) * 7.0
    u = torch.zeros(n, n, dtype=torch.float64)
    for i in range(n):
        for j in range(i + 1, n):
            u[i, j] = 865723.02 / float(abs(x[i] - x[j])) ** 6
    full = (1 << n) - 1
    return dict(dt=0.002, det_masks=[full], det_values=dv, amp_masks=[full], amp_values=av, pair_u=u)


@pytest.mark.parametrize("n", [10, 16, 20])
def test_hermitian_and_linear(cuda_device, n):
    pr = _program(n)
    dev = cuda_device
    g = torch.Generator().manual_seed(n)
    a = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=g).to(dev)
    b = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=g).to(dev)
    args = (0.0371, pr["det_values"], pr["amp_values"], pr["pair_u"], pr["det_masks"], pr["amp_masks"], pr["dt"])
    Ha = torch.ops.pulser_diff_b200.hpsi(a, *args)
    Hb = torch.ops.pulser_diff_b200.hpsi(b, *args)
    lhs = torch.vdot(a.flatten(), Hb.flatten())
    rhs = torch.vdot(b.flatten(), Ha.flatten()).conj()
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)
    Hab = torch.ops.pulser_diff_b200.hpsi(a + 2j * b, *args)
    assert (Hab - (Ha + 2j * Hb)).abs().max() < 1e-10 * Hab.abs().max()


@pytest.mark.parametrize("n", [14, 20])
def test_norm_conservation_and_fd_gradient(cuda_device, n):
    pr = _program(n, T=16)
    dev = cuda_device
    psi0 = torch.zeros(1, 2 ** n, dtype=torch.complex128, device=dev)
    psi0[0, -1] = 1.0
    tsave = torch.tensor([0.0, 0.012, 0.024], dtype=torch.float64)
    av = pr["amp_values"].clone().requires_grad_(True)
    obs = torch.arange(2 ** n, device=dev).remainder(7).to(torch.float64)

    def f(av_):
        st = ops.evolve(psi0, tsave, pr["det_values"], av_, pr["pair_u"], n_qubits=n,
                        kind=_cabi.PD_KET, dt=pr["dt"], det_masks=pr["det_masks"],
                        amp_masks=pr["amp_masks"], options=_cabi.Options(atol=1e-12, rtol=1e-10))
        return st, (obs * st[-1, 0].abs() ** 2).sum()

    st, val = f(av)
    assert (st.detach().norm(dim=-1) - 1).abs().max() < 1e-8
    (g,) = torch.autograd.grad(val, [av])
    eps = 1e-5
    for idx in (3, 9):
        d = torch.zeros_like(av.detach())
        d[0, idx] = eps
        fd = (f(av.detach() + d)[1] - f(av.detach() - d)[1]) / (2 * eps)
        assert abs(fd.item() - g[0, idx].real.item()) < 1e-6 * max(1.0, abs(fd.item()))


@pytest.mark.parametrize("n", [14, 18, 22])
def test_tiled_equals_gather(cuda_device, n):
    pr = _program(n, T=16)
    dev = cuda_device
    psi0 = torch.randn(1, 2 ** n, dtype=torch.complex128, generator=torch.Generator().manual_seed(5)).to(dev)
    psi0 /= psi0.norm()
    tsave = torch.tensor([0.0, 0.004], dtype=torch.float64)
    outs = []
    for path in (1, 0):
        outs.append(ops.evolve(psi0, tsave, pr["det_values"], pr["amp_values"], pr["pair_u"],
                               n_qubits=n, kind=_cabi.PD_KET, dt=pr["dt"], det_masks=pr["det_masks"],
                               amp_masks=pr["amp_masks"], options=_cabi.Options(path=path)))
    assert (outs[0] - outs[1]).abs().max() < 1e-12
