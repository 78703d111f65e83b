// Shared host-side types of the pulser_diff_b200 engine.
//
// Everything here is plain C++ (no CUDA) so the orchestration in engine.hpp can be
// compiled against the CUDA backend (product, libpulser_diff_b200.so) and against the
// host stand-in under tests/emu (test infrastructure only).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pulser_diff_b200.h"

#if defined(__CUDACC__)
#define PD_HD __host__ __device__
#else
#define PD_HD
#endif

namespace pd {

constexpr int kMaxQubits = 40;   // ket: up to 2^40 amplitudes addressable; density: N <= 16
constexpr int kMaxSitesDensity = 16;

struct alignas(16) cplx {
  double re, im;
};
PD_HD inline cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
PD_HD inline cplx operator-(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
PD_HD inline cplx operator*(cplx a, cplx b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
PD_HD inline cplx operator*(double a, cplx b) { return {a * b.re, a * b.im}; }
PD_HD inline cplx conj(cplx a) { return {a.re, -a.im}; }
PD_HD inline void fma_acc(cplx& acc, cplx a, cplx b) {
  acc.re = fma(a.re, b.re, fma(-a.im, b.im, acc.re));
  acc.im = fma(a.re, b.im, fma(a.im, b.re, acc.im));
}

// Storage type of state-vector amplitudes in device memory.  The default build keeps complex128 (amp_t is
// cplx itself); the complex64 build (-DPD_C64 -> libpulser_diff_b200_c64.so, north_star's optional 1e-5 tier)
// stores {float re, im}: half the bytes per vector pass.  Coefficients, reductions, gradients, times and the
// step controller stay double in both builds.  The gather kernels convert on load/store and compute in
// double; the stream kernels compute in the storage precision (areal).
#if defined(PD_C64)
using areal = float;
struct alignas(8) amp_t {
  float re, im;
  amp_t() = default;
  PD_HD amp_t(float r, float i) : re(r), im(i) {}
  PD_HD amp_t(cplx c) : re((float)c.re), im((float)c.im) {}
  PD_HD operator cplx() const { return cplx{(double)re, (double)im}; }
};
constexpr bool kC64 = true;
#else
using areal = double;
using amp_t = cplx;
constexpr bool kC64 = false;
#endif

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// Dormand-Prince 5(4) tableau (the published standard tableau).
struct Tableau {
  double alpha[6] = {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
  double beta[6][6] = {
      {1.0 / 5, 0, 0, 0, 0, 0},
      {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
      {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
      {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
      {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
      {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
  double b5[7] = {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0.0};
  double b4[7] = {5179.0 / 57600, 0.0,          7571.0 / 16695, 393.0 / 640,
                  -92097.0 / 339200, 187.0 / 2100, 1.0 / 40};
};

// What one application of the generator needs: a complex scale for the static diagonal
// and one small matrix per site (qubit).  width 1: ket, T is 2x2 over the qubit's bit.
// width 2: density, T is 4x4 over (row bit, col bit), pair index p = 2*rowbit + colbit.
//   out[idx] = kappa * Dstat(idx) * v[idx] + sum_q sum_p' T_q[p(idx)][p'] * v[idx with site q := p']
struct SiteOps {
  int nsites;
  int width;
  cplx kappa;
  cplx T[kMaxQubits * 4];  // ket layout [q][a][a']; density uses SiteOpsDensity below
};
// The same super-operator in its generator-specific form (gather_kernels.cu k_apply_lindblad): instead of a
// general 4x4 matrix per site, the pieces it is made of --
//   out[e] = kappa * [ (Dint[r] - Dint[c] + sum_b d_b ((a_b == 0) - (c_b == 0))) rho[e]
//                      + sum_b ( h_b(a_b) rho[e ^ row bit b] - conj(h_b(c_b)) rho[e ^ column bit b] ) ]
//            + sum_b dd[p_b] rho[e] + sum over the dissipator's off-diagonal entries (p <- p') and the sites b with
//              p_b = p of off_c * rho[e with site b set to p'],
// h_b(x) = x ? g_b : conj(g_b), p_b = 2 a_b + c_b; arrays indexed by BIT position b (site q = nq - 1 - b).
struct LindbladForm {
  int ok = 0, nq = 0, real_drive = 0, n_off = 0;
  cplx kappa{0, 0};
  double d[kMaxSitesDensity], gre[kMaxSitesDensity], gim[kMaxSitesDensity];
  cplx dd[4];
  int off_p[12], off_pp[12];
  cplx off_c[12];
};
struct SiteOpsDensity {
  int nsites;
  cplx kappa;
  unsigned nzmask[kMaxSitesDensity];  // bit (p*4+p') set <=> T[q][p][p'] != 0 for some use
  cplx T[kMaxSitesDensity * 16];      // [q][p][p']
  LindbladForm form;
};

struct Geometry {
  int kind;      // PD_KET / PD_DENSITY
  int nq;        // qubits
  int nbits;     // nq or 2*nq
  size_t dim;    // 2^nbits
  int batch;
  const double* diag;  // device Dint[2^nq]
  // kets of nq >= 16: Dint split for 4096-amplitude tiles (stream family), tile = index >> 12:
  //   [0, 4096)            Dlow[l]   pairs inside the low 12 bits
  //   [4096, 4096 + T)     Dhh[tile] pairs inside the bits above            (T = 2^(nq-12))
  //   [4096 + T, ... 13T)  ch[tile][p] = sum over occupied high bits of U(p, .), p < 12
  // so that Dint[s] = Dhh + Dlow + sum over occupied low bits of ch (nullptr: not built)
  const double* diag_parts = nullptr;
};

// one accepted step as handed to the small-register adjoint sweep
struct SkStepHost {
  double t, dt;
  int interval, clipped;
};

// Interpolation rule of the reference closure H_t, quirk included
// (reference pulser_diff/hamiltonian.py:532-542).
struct Interp {
  int i1, i2;
  double x;  // (t - i1*dt) / dt is applied as  v1 + (v2 - v1) * (t - i1*dt) / dt
};
inline Interp interp_index(double t, double dt, int n_samples) {
  double fl = std::floor(t / dt);
  long i1 = (long)std::min(fl, (double)(n_samples - 2));
  if (i1 < 0) i1 = 0;
  long i2 = std::min<long>(i1 + 1, n_samples - 2);
  if (i2 < 0) i2 = 0;
  return {(int)i1, (int)i2, 0.0};
}
inline double interp_real(const double* v, double t, double dt, const Interp& ix) {
  return v[ix.i1] + (v[ix.i2] - v[ix.i1]) * (t - ix.i1 * dt) / dt;
}

// Host description of the pulse program + register (what crosses the boundary).
struct Program {
  int nq = 0;
  int kind = PD_KET;
  int n_samples = 0;
  double dt = 0.0;
  std::vector<uint64_t> det_masks, amp_masks;
  std::vector<double> det_values;  // [n_det][n_samples]
  std::vector<double> amp_values;  // [n_amp][n_samples][2]
  std::vector<double> pair_u;      // [nq*nq]
  uint64_t version = 0;            // bumped whenever the coefficient tables change (device-side caches)
  int n_collapse = 0;
  std::vector<cplx> collapse;      // [n_ops][2][2]
  cplx dsup[16];                   // static dissipator on the (row,col) pair space, [p][p']
  bool has_dsup = false;
  int n_det() const { return (int)det_masks.size(); }
  int n_amp() const { return (int)amp_masks.size(); }

  // per-qubit coefficients at time t: d_q (of r_q) and g_q (of |g><r|_q)
  void coefficients(double t, double* d, cplx* g) const {
    for (int q = 0; q < nq; ++q) { d[q] = 0.0; g[q] = {0.0, 0.0}; }
    if (n_samples < 2) return;
    Interp ix = interp_index(t, dt, n_samples);
    for (int k = 0; k < n_det(); ++k) {
      double c = interp_real(&det_values[(size_t)k * n_samples], t, dt, ix);
      for (int q = 0; q < nq; ++q)
        if (det_masks[k] >> q & 1) d[q] += c + c;  // ham_mat + ham_mat.adjoint()
    }
    for (int k = 0; k < n_amp(); ++k) {
      const double* v = &amp_values[(size_t)k * n_samples * 2];
      double re = v[2 * ix.i1] + (v[2 * ix.i2] - v[2 * ix.i1]) * (t - ix.i1 * dt) / dt;
      double im = v[2 * ix.i1 + 1] + (v[2 * ix.i2 + 1] - v[2 * ix.i1 + 1]) * (t - ix.i1 * dt) / dt;
      for (int q = 0; q < nq; ++q)
        if (amp_masks[k] >> q & 1) { g[q].re += re; g[q].im += im; }
    }
  }

  // Static dissipator super-operator on one qubit's (row bit, col bit) space:
  //   D[(a,b)][(a',b')] = sum_k L[a][a'] conj(L[b][b']) - 1/2 M[a][a'] d_bb' - 1/2 d_aa' M[b'][b]
  // with M = sum_k L^dagger L   (hamiltonian.py:98-143 ops).
  void build_dsup() {
    for (auto& z : dsup) z = {0, 0};
    has_dsup = n_collapse > 0;
    cplx M[2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
    for (int k = 0; k < n_collapse; ++k) {
      const cplx* L = &collapse[(size_t)k * 4];
      for (int a = 0; a < 2; ++a)
        for (int ap = 0; ap < 2; ++ap)
          for (int c = 0; c < 2; ++c) fma_acc(M[a][ap], conj(L[c * 2 + a]), L[c * 2 + ap]);
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
          for (int ap = 0; ap < 2; ++ap)
            for (int bp = 0; bp < 2; ++bp)
              fma_acc(dsup[(a * 2 + b) * 4 + (ap * 2 + bp)], L[a * 2 + ap], conj(L[b * 2 + bp]));
    }
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int ap = 0; ap < 2; ++ap)
          for (int bp = 0; bp < 2; ++bp) {
            cplx& z = dsup[(a * 2 + b) * 4 + (ap * 2 + bp)];
            if (b == bp) z = z - 0.5 * M[a][ap];
            if (a == ap) z = z - 0.5 * M[bp][b];
          }
  }

  // mode 0: forward generator (-iH, or the Lindbladian); 1: its adjoint; 2: plain H (ket).
  void site_ops_ket(double t, int mode, SiteOps& so) const {
    double d[kMaxQubits];
    cplx g[kMaxQubits];
    coefficients(t, d, g);
    cplx phase = mode == 2 ? cplx{1, 0} : (mode == 0 ? cplx{0, -1} : cplx{0, 1});
    so.nsites = nq;
    so.width = 1;
    so.kappa = phase;
    for (int q = 0; q < nq; ++q) {
      cplx* T = &so.T[q * 4];
      T[0] = phase * cplx{d[q], 0};  // a = 0 (Rydberg): r_q = 1
      T[3] = {0, 0};
      T[2] = phase * g[q];           // row a=1 (g) <- a'=0 (r): |g><r|
      T[1] = phase * conj(g[q]);     // row a=0 (r) <- a'=1 (g)
    }
  }
  void site_ops_density(double t, int mode, SiteOpsDensity& so) const {
    double d[kMaxQubits];
    cplx g[kMaxQubits];
    coefficients(t, d, g);
    so.nsites = nq;
    so.kappa = mode == 0 ? cplx{0, -1} : cplx{0, 1};
    {
      LindbladForm& lf = so.form;
      lf = LindbladForm{};
      lf.ok = 1; lf.nq = nq; lf.kappa = so.kappa; lf.real_drive = 1;
      for (int q = 0; q < nq; ++q) {
        const int b = nq - 1 - q;
        lf.d[b] = d[q]; lf.gre[b] = g[q].re; lf.gim[b] = g[q].im;
        if (g[q].im != 0.0) lf.real_drive = 0;
      }
      for (int p = 0; p < 4; ++p)
        for (int pp = 0; pp < 4; ++pp) {
          cplx z = !has_dsup ? cplx{0, 0} : (mode == 0 ? dsup[p * 4 + pp] : conj(dsup[pp * 4 + p]));
          if (p == pp) lf.dd[p] = z;
          else if (z.re != 0.0 || z.im != 0.0) {
            lf.off_p[lf.n_off] = p; lf.off_pp[lf.n_off] = pp; lf.off_c[lf.n_off] = z;
            ++lf.n_off;
          }
        }
    }
    const cplx mi{0, -1}, pi{0, 1};
    for (int q = 0; q < nq; ++q) {
      cplx F[16];
      for (int i = 0; i < 16; ++i) F[i] = has_dsup ? dsup[i] : cplx{0, 0};
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          int p = a * 2 + b;
          double dd = d[q] * ((a == 0 ? 1.0 : 0.0) - (b == 0 ? 1.0 : 0.0));
          F[p * 4 + p] = F[p * 4 + p] + mi * cplx{dd, 0};
          cplx ha = a == 1 ? g[q] : conj(g[q]);   // H[r, r^m]
          cplx hb = b == 1 ? g[q] : conj(g[q]);
          int prow = (1 - a) * 2 + b, pcol = a * 2 + (1 - b);
          F[p * 4 + prow] = F[p * 4 + prow] + mi * ha;        // -i (H rho)
          F[p * 4 + pcol] = F[p * 4 + pcol] + pi * conj(hb);  // +i (rho H)
        }
      cplx* T = &so.T[q * 16];
      unsigned nz = 0;
      for (int p = 0; p < 4; ++p)
        for (int pp = 0; pp < 4; ++pp) {
          cplx z = mode == 0 ? F[p * 4 + pp] : conj(F[pp * 4 + p]);
          T[p * 4 + pp] = z;
          if (z.re != 0.0 || z.im != 0.0) nz |= 1u << (p * 4 + pp);
        }
      so.nzmask[q] = nz;
    }
  }
};

// Symmetric tridiagonal eigen-decomposition (implicit QL, as in EISPACK tql2).
// d[n] diagonal -> eigenvalues; e[n] sub-diagonal in e[1..n-1]; z (n x n, row-major) -> vectors.
inline void tridiag_eig(int n, std::vector<double>& d, std::vector<double>& e,
                        std::vector<double>& z) {
  z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) z[(size_t)i * n + i] = 1.0;
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) throw Error(PD_ERR_STATE, "tridiag_eig: no convergence");
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i], b = c * e[i];
          r = std::hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r; c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (int k = 0; k < n; ++k) {
            double f2 = z[(size_t)k * n + i + 1];
            z[(size_t)k * n + i + 1] = s * z[(size_t)k * n + i] + c * f2;
            z[(size_t)k * n + i] = c * z[(size_t)k * n + i] - s * f2;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p; e[l] = g; e[m] = 0.0;
      }
    } while (m != l);
  }
}

// w = exp(-i * delta * T) e_1 for the real symmetric tridiagonal T (alpha, beta).
inline std::vector<cplx> tridiag_expm_e1(const std::vector<double>& alpha,
                                         const std::vector<double>& beta, double delta) {
  int n = (int)alpha.size();
  std::vector<double> d(alpha), e(n, 0.0), z;
  for (int i = 1; i < n; ++i) e[i] = beta[i - 1];
  tridiag_eig(n, d, e, z);
  std::vector<cplx> w(n, cplx{0, 0});
  for (int k = 0; k < n; ++k) {
    double q0 = z[(size_t)0 * n + k];
    cplx ph{std::cos(delta * d[k]) * q0, -std::sin(delta * d[k]) * q0};
    for (int i = 0; i < n; ++i) w[i] = w[i] + z[(size_t)i * n + k] * ph;
  }
  return w;
}

}  // namespace pd
