"""Independent second opinion for the oracle: NumPy / SciPy only, dense matrices, no code shared with
``oracle/`` or ``pulser_diff_b200/``.

The reference pins no state to better than 4-6 printed digits and no gradient at all, and
its solver lives in pyqtorch, which is not installable here.  This script restates the PHYSICS of the path

    H(t) = sum_{i<j} C6/r_ij^6 n_i n_j + sum_q [ Omega_q(t)/2 (e^{-i phi_q(t)} |g><r|_q + h.c.) - delta_q(t) n_q ]

with the reference's sampling rules (sub-sampling hamiltonian.py:83-91, the interpolation rule with its
index clamp hamiltonian.py:532-542, evaluation times backend.py:329-373) and solves it with methods
unrelated to the oracle's Dormand-Prince / Lanczos code:

* DP5_SE / DP5_ME cases: ``scipy.integrate.solve_ivp(method="DOP853", rtol=1e-13, atol=1e-15)``, restarted
  at every kink of the piecewise-linear coefficients -- the exact solution of the ODE to ~1e-12;
* KRYLOV_SE cases: ``scipy.linalg.expm(-i H(t_next) dt)`` per interval (H frozen at the interval end,
  SURVEY.md Appendix A.5);
* gradients: fourth-order central differences (Richardson) of the same solutions w.r.t. four scalar
  parameters that reach every autograd leaf class of the path: an amplitude scale, a detuning scale, a
  phase chirp and the x coordinate of atom 1 (i.e. the pair couplings).

Output: ``tests/golden/independent_kats.json``.  ``tests/test_independent_kats.py`` checks the oracle (and,
on the GPU tier, the CUDA path) against it with both at tight solver tolerances.

The one call into torch is ``torch.linspace(..., dtype=torch.int)``: the sub-sampling INDEX rule is
"whatever torch.linspace rounds to" in the reference, so the same library function defines it here.

Run:  python tests/golden/make_independent_kats.py      (about a minute)
"""
import json
import math
import os

import numpy as np
import torch
from scipy.integrate import solve_ivp
from scipy.linalg import expm

C6_70, C6_60 = 5420158.53, 865723.02

N_OP = np.array([[1, 0], [0, 0]], dtype=complex)        # |r><r|, basis order (r, g)
SIG_GR = np.array([[0, 0], [1, 0]], dtype=complex)      # |g><r|
Z = np.array([[1, 0], [0, -1]], dtype=complex)


def embed(n, q, op):
    out = np.array([[1.0 + 0j]])
    for k in range(n):
        out = np.kron(out, op if k == q else np.eye(2))
    return out


# ---- pulse samples ---------------------------------------------------------
def constant(d, v):
    return np.full(d, float(v))


def ramp(d, a, b):
    return np.linspace(a, b, d)


def blackman(d, area):
    w = np.clip(np.blackman(d), 0, np.inf)
    return w * area / w.sum() / 1e-3


def tanh_sequence(durs_us, amps, dets, phases):
    """duration mode: 1-ns pulses carrying summed tanh boxes (model.py:301-368, waveform_funcs.py:9-27)."""
    total = sum(int(d * 1000) for d in durs_us) + 5
    t = np.arange(total, dtype=float)
    out = {k: np.zeros(total) for k in ("amp", "det", "phase")}
    ti = None
    for d, a, de, ph in zip(durs_us, amps, dets, phases):
        tf = d if ti is None else ti + d
        fall = 0.5 * (1 + np.tanh(-(t - tf * 1000)))
        env = fall if ti is None else 0.5 * (1 + np.tanh(t - ti * 1000)) + fall - 1
        for key, v in (("amp", a), ("det", de), ("phase", ph)):
            out[key] += v * env
        ti = tf
    return out["amp"], out["det"], out["phase"]


def cases():
    two = [[-4.0, 0.0], [4.0, 0.0]]
    kb_amp = np.concatenate([constant(1000, 5.0), blackman(800, math.pi)])
    kb_det = np.concatenate([constant(1000, 0.0), ramp(800, 5.0, 0.0)])
    z18 = np.zeros(1800)
    out = {}
    out["K-A"] = dict(coords=[[0, 0], [0, 8], [8, 0], [8, 8]], c6=C6_70, rate=0.1, solver="dp5_se",
                      amp=np.concatenate([blackman(800, math.pi), constant(800, 5.0)]),
                      det=np.concatenate([ramp(800, -5.0, 0.0), constant(800, 0.0)]), phase=np.zeros(1600))
    out["K-B"] = dict(coords=two, c6=C6_70, rate=0.5, solver="krylov_se", amp=kb_amp, det=kb_det, phase=z18)
    out["K-C"] = dict(coords=[[0.5, 0.4], [8.3, 0.1]], c6=C6_70, rate=0.5, solver="krylov_se",
                      amp=np.concatenate([constant(1000, 5.0), blackman(800, 3.14)]), det=kb_det, phase=z18)
    a, d, p = tanh_sequence([0.4, 0.4, 0.2], [2.0, 5.0, 3.0], [0.5, 0.0, 1.0], [0.0, 0.0, 0.0])
    out["K-D"] = dict(coords=two, c6=C6_70, rate=0.5, solver="krylov_se", amp=a, det=d, phase=p)
    x = np.arange(300) / 300
    out["K-E"] = dict(coords=two, c6=C6_70, rate=0.5, solver="krylov_se",
                      amp=np.concatenate([kb_amp, 6.0 * np.sin(math.pi * x) * np.exp(-2.0 * x)]),
                      det=np.concatenate([kb_det, constant(300, 1.5)]), phase=np.zeros(2100))
    out["K-F"] = dict(coords=two, c6=C6_70, rate=0.5, solver="dp5_me", amp=kb_amp, det=kb_det, phase=z18,
                      dephasing_rate=2.0)
    out["K-G"] = dict(coords=[[-3.25, 0.0], [3.25, 0.0]], c6=C6_60, rate=0.05, solver="dp5_se",
                      amp=constant(8 * 131, 5.0), det=constant(8 * 131, 5.0), phase=constant(8 * 131, 5.0),
                      psi0="eye")
    # C1: two atoms, constant global pulse, every sample kept
    out["C1"] = dict(coords=[[0.0, 0.0], [8.0, 0.0]], c6=C6_70, rate=1.0, solver="dp5_se",
                     amp=constant(1000, 5.0), det=constant(1000, 0.0), phase=np.zeros(1000), tsave="Minimal")
    # C2-small: 4-atom chain, smooth sweep, the loss of the state-preparation workload
    k = np.arange(1100) / 1100
    out["C2-small"] = dict(coords=[[7.0 * i, 0.0] for i in range(4)], c6=C6_60, rate=0.05, solver="dp5_se",
                           amp=6.0 * np.sin(math.pi * k) ** 2, det=-8.0 + 16.0 * k, phase=np.zeros(1100),
                           loss="nn")
    return out


# ---- the path, restated ------------------------------------------------------------------------------
class System:
    def __init__(self, c, theta):
        s_amp, s_det, chirp, dx = theta
        self.c = c
        coords = np.array(c["coords"], dtype=float)
        coords[1, 0] += dx
        self.n = n = len(coords)
        T = len(c["amp"])
        z = np.zeros(1)
        amp = np.concatenate([s_amp * c["amp"], z])                    # extend_duration(T + 1)
        det = np.concatenate([s_det * c["det"], z])
        ph = c["phase"] + chirp * np.arange(T) / T
        ph = np.concatenate([ph, ph[-1:]])
        L = T + 1
        ns = int(c["rate"] * L)
        idx = torch.linspace(0, L - 1, ns, dtype=torch.int).numpy()
        self.drive = (0.5 * amp * np.exp(-1j * ph))[idx]
        self.det = (-0.5 * det)[idx]
        self.dt = 0.001 / c["rate"]
        self.ns = ns
        self.T_us = T / 1000
        times = (np.arange(L) / 1000)[idx]
        ev = np.array([]) if c.get("tsave") == "Minimal" else times
        self.tsave = np.unique(np.concatenate([ev, [0.0, self.T_us]]))
        self.H0 = np.zeros((2 ** n, 2 ** n), dtype=complex)
        for i in range(n):
            for j in range(i + 1, n):
                r = np.linalg.norm(coords[i] - coords[j])
                self.H0 += c["c6"] / r ** 6 * embed(n, i, N_OP) @ embed(n, j, N_OP)
        self.Sgr = sum(embed(n, q, SIG_GR) for q in range(n))
        self.Nsum = sum(embed(n, q, N_OP) for q in range(n))

    def coef(self, v, t):
        i1 = max(int(min(math.floor(t / self.dt), self.ns - 2)), 0)
        i2 = min(i1 + 1, self.ns - 2)
        return v[i1] + (v[i2] - v[i1]) * (t - i1 * self.dt) / self.dt

    def H(self, t):
        g = self.coef(self.drive, t)
        d = self.coef(self.det, t)
        return self.H0 + g * self.Sgr + np.conj(g) * self.Sgr.conj().T + 2 * d * self.Nsum

    def generator(self):
        if self.c["solver"] != "dp5_me":
            return lambda t: -1j * self.H(t)
        S = 2 ** self.n
        I = np.eye(S)
        Ls = [math.sqrt(self.c["dephasing_rate"] / 2) * embed(self.n, q, Z) for q in range(self.n)]
        D = sum(np.kron(L, L.conj()) - 0.5 * np.kron(L.conj().T @ L, I) - 0.5 * np.kron(I, (L.conj().T @ L).T)
                for L in Ls)                                   # row-major vec(rho)
        return lambda t: -1j * (np.kron(self.H(t), I) - np.kron(I, self.H(t).T)) + D

    def solve(self):
        S = 2 ** self.n
        if self.c.get("psi0") == "eye":
            y = np.eye(S, dtype=complex)
        else:
            y = np.zeros((S, 1), dtype=complex)
            y[-1] = 1.0
        if self.c["solver"] == "dp5_me":
            y = (y @ y.conj().T).reshape(-1, 1)
        states = [y]
        if self.c["solver"] == "krylov_se":
            for a, b in zip(self.tsave[:-1], self.tsave[1:]):
                y = expm(-1j * (b - a) * self.H(b)) @ y
                states.append(y)
            return np.array(states)
        G = self.generator()
        shape = y.shape
        kinks = np.arange(1, self.ns) * self.dt
        for a, b in zip(self.tsave[:-1], self.tsave[1:]):
            pts = [a] + [k for k in kinks if a + 1e-12 < k < b - 1e-12] + [b]
            for p, q in zip(pts[:-1], pts[1:]):
                sol = solve_ivp(lambda t, v: (G(t) @ v.reshape(shape)).reshape(-1), (p, q), y.reshape(-1),
                                method="DOP853", rtol=1e-13, atol=1e-15)
                y = sol.y[:, -1].reshape(shape)
            states.append(y)
        return np.array(states)

    def observables(self, states):
        n, S = self.n, 2 ** self.n
        zdiag = np.real(np.diag(sum(embed(n, q, Z) for q in range(n))))
        if self.c["solver"] == "dp5_me":
            rho = states.reshape(len(states), S, S)
            sumz = np.real(np.einsum("tii,i->t", rho, zdiag))
        else:
            sumz = np.real(np.einsum("tsb,s,tsb->t", states.conj(), zdiag, states))
        if self.c.get("psi0") == "eye":
            h = np.array([[1, 1], [1, -1]]) / math.sqrt(2)
            loss = 1 - abs(np.trace(np.kron(h, h).conj().T @ states[-1])) / S
        elif self.c.get("loss") == "nn":
            nn = np.real(np.diag(sum(embed(n, q, N_OP) @ embed(n, q + 1, N_OP) for q in range(n - 1))))
            loss = float(np.real(np.einsum("sb,s,sb->", states[-1].conj(), nn, states[-1])))
        else:
            loss = float(sumz[-1])
        return sumz, float(loss)


THETA0 = (1.0, 1.0, 0.3, 0.0)
STEPS = (5e-4, 5e-4, 2e-3, 2e-4)      # small against the phase each parameter winds up (<= ~50 rad per unit)


def main():
    gold = {"_doc": "see tests/golden/make_independent_kats.py", "theta0": list(THETA0),
            "theta_names": ["amp scale", "det scale", "phase chirp (rad over the sequence)", "x of atom 1 (um)"]}
    for name, c in cases().items():
        sysm = System(c, THETA0)
        st = sysm.solve()
        sumz, loss = sysm.observables(st)
        grad = []
        for k, h in enumerate(STEPS):
            f = {}
            for m in (-2, -1, 1, 2):
                th = list(THETA0)
                th[k] += m * h
                s2 = System(c, th)
                f[m] = s2.observables(s2.solve())[1]
            grad.append((8 * (f[1] - f[-1]) - (f[2] - f[-2])) / (12 * h))
        last = st[-1]
        gold[name] = {"solver": c["solver"], "tsave": sysm.tsave.tolist(), "sum_z": sumz.tolist(), "loss": loss,
                      "grad": grad, "final_re": last.real.reshape(-1).tolist(),
                      "final_im": last.imag.reshape(-1).tolist(), "final_shape": list(last.shape)}
        print(name, "loss", loss, "grad", grad, flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "independent_kats.json")
    json.dump(gold, open(path, "w"))
    print("wrote", path)


if __name__ == "__main__":
    main()
