// TEST INFRASTRUCTURE ONLY: a plain-C++ stand-in for the CUDA kernels so the host-side
// orchestration of pulser_diff_b200/csrc/engine.hpp (step controller, tape, discrete adjoint,
// gradient distribution, Krylov) and the Python layer above the C ABI can be exercised on a
// box without a GPU.  It is compiled into tests/emu/libpd_emu.so, which the product package
// never loads on its own (pulser_diff_b200/_cabi.py only accepts it when a test passes the
// path explicitly).  "Device" pointers are host pointers here; streams are ignored.
#pragma once
#include <chrono>
#include <cstdlib>

#include "../../pulser_diff_b200/csrc/pd_common.hpp"

namespace pd {

class HostBackend {
 public:
  static constexpr bool is_cuda = false;
  int device = 0;
  static int push_device(int) { return 0; }
  static void pop_device(int) {}
  static bool on_device(const void*) { return false; }
  static void transfer_counters(long long* h2d, long long* d2h, bool) { if (h2d) *h2d = 0; if (d2h) *d2h = 0; }
  int path = 0;
  explicit HostBackend(int) {}
  void* alloc(size_t bytes) {
    void* p = nullptr;
    if (posix_memalign(&p, 64, std::max<size_t>(bytes, 64)) != 0) throw Error(PD_ERR_STATE, "alloc");
    return p;
  }
  void free(void* p) { std::free(p); }
  void zero(void* p, size_t b, void*) { std::memset(p, 0, b); }
  void d2d(void* d, const void* s, size_t b, void*) { std::memmove(d, s, b); }
  void d2h(void* d, const void* s, size_t b, void*) { std::memmove(d, s, b); }
  void sync(void*) {}
  std::chrono::steady_clock::time_point t0_;
  void timer_start(void*) { t0_ = std::chrono::steady_clock::now(); }
  double timer_stop_ms(void*) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count();
  }
  size_t reduce_scratch_bytes(const Geometry&) { return 64; }
  // the small-register cooperative kernels exist only in the CUDA build
  bool small_supported(const Geometry&, const Program&) { return false; }
  bool small_units_supported(const Geometry&, const Program&) { return false; }
  int small_lanczos(const Geometry&, const Program&, const amp_t*, amp_t*, int, double, int, int, double, double*,
                    double*, double*, void*) {
    throw Error(PD_ERR_STATE, "small-register kernels need the CUDA build");
  }
  int small_forward(const Geometry&, const Program&, const Tableau&, const pd_options&, int, const amp_t*,
                    const double*, const double*, const double*, int, amp_t*,
                    std::vector<std::vector<pd_step_record>>&, bool, uint64_t*, void*) {
    throw Error(PD_ERR_STATE, "small-register kernels need the CUDA build");
  }
  int small_backward_units(const Geometry&, const Program&, const Tableau&, const std::vector<double>&, int,
                           const double*, const double*, uint64_t, const amp_t*, amp_t*, double*, double*, void*) {
    return 0;
  }
  void small_unit_counts(uint64_t, int, int* acc, int* att) { *acc = *att = -1; }
  int small_backward(const Geometry&, const Program&, const Tableau&, const std::vector<double>&, int,
                     const double*, const double*, const std::vector<std::vector<SkStepHost>>&, uint64_t,
                     const amp_t*, bool, double*, amp_t*, std::vector<std::vector<double>>&, void*) {
    return 0;
  }
  size_t segment_budget_bytes() { return segment_budget; }
  size_t segment_budget = (size_t)1 << 28;

  void build_diag_parts(double* parts, int nq, const double* diag, void*) {
    const size_t T = (size_t)1 << (nq - 12);
    for (size_t l = 0; l < 4096; ++l) parts[l] = diag[(((size_t)1 << nq) - 4096) | l];
    for (size_t h = 0; h < T; ++h) {
      parts[4096 + h] = diag[(h << 12) | 0xFFF];
      for (int p = 0; p < 12; ++p)
        parts[4096 + T + h * 12 + p] = diag[(h << 12) | (0xFFF & ~(1u << p))] - diag[(h << 12) | 0xFFF];
    }
  }
  void build_diag(double* diag, int nq, const double* u, void*) {
    size_t dim = (size_t)1 << nq;
    for (size_t s = 0; s < dim; ++s) {
      double acc = 0.0;
      for (int i = 0; i < nq; ++i) {
        if (s >> (nq - 1 - i) & 1) continue;
        for (int j = i + 1; j < nq; ++j)
          if (!(s >> (nq - 1 - j) & 1)) acc += u[i * nq + j];
      }
      diag[s] = acc;
    }
  }
  int lincomb(const Geometry& g, amp_t* out, int n_in, const amp_t* const* ins, const double* w, void*) {
    size_t n = g.dim * g.batch;
    for (size_t i = 0; i < n; ++i) {
      double re = 0, im = 0;
      for (int j = 0; j < n_in; ++j) { re += w[j] * ins[j][i].re; im += w[j] * ins[j][i].im; }
      out[i] = cplx{re, im};
    }
    return 1;
  }
  int lincomb_c(const Geometry& g, amp_t* out, int m, const amp_t* basis, size_t stride, const cplx* ws, void*) {
    size_t n = g.dim * g.batch;
    for (size_t i = 0; i < n; ++i) {
      cplx acc{0, 0};
      for (int j = 0; j < m; ++j) acc = acc + ws[j] * basis[(size_t)j * stride + i];
      out[i] = acc;
    }
    return 1;
  }
  const amp_t* combine(const Geometry& g, amp_t* comb, int n_in, const amp_t* const* ins, const double* w,
                      amp_t* scratch) {
    if (n_in > 1 || w[0] != 1.0) {
      amp_t* dst = comb ? comb : scratch;
      lincomb(g, dst, n_in, ins, w, nullptr);
      return dst;
    }
    if (comb) std::memmove(comb, ins[0], sizeof(amp_t) * g.dim * g.batch);
    return ins[0];
  }
  int stage_ket(const Geometry& g, amp_t* out, amp_t* comb, int n_in, const amp_t* const* ins,
                const double* w, const SiteOps& so, amp_t* scratch, void*) {
    const amp_t* in = combine(g, comb, n_in, ins, w, scratch);
    int nq = g.nq;
    for (size_t idx = 0; idx < g.dim * g.batch; ++idx) {
      size_t s = idx & (g.dim - 1);
      cplx acc = (so.kappa * cplx{g.diag[s], 0}) * in[idx];
      for (int q = 0; q < nq; ++q) {
        size_t m = (size_t)1 << (nq - 1 - q);
        int a = (s & m) ? 1 : 0;
        acc = acc + so.T[q * 4 + a * 2 + a] * in[idx] + so.T[q * 4 + a * 2 + (1 - a)] * in[idx ^ m];
      }
      out[idx] = acc;
    }
    return 1;
  }
  int dp5_step_ket(const Geometry&, const amp_t*, amp_t* const*, amp_t*, const SiteOps*, const Tableau&,
                   const double*, double, double, double, amp_t*, amp_t*, double*, double*, void*) {
    return 0;   // no fused path in the stand-in: the engine falls back to stage-by-stage
  }
  int stage_density(const Geometry& g, amp_t* out, amp_t* comb, int n_in, const amp_t* const* ins,
                    const double* w, const SiteOpsDensity& so, amp_t* scratch, void*) {
    const amp_t* in = combine(g, comb, n_in, ins, w, scratch);
    int nq = g.nq;
    size_t S = (size_t)1 << nq;
    for (size_t idx = 0; idx < g.dim * g.batch; ++idx) {
      size_t e = idx & (g.dim - 1), r = e >> nq, c = e & (S - 1);
      cplx acc = (so.kappa * cplx{g.diag[r] - g.diag[c], 0}) * in[idx];
      for (int q = 0; q < nq; ++q) {
        size_t mc = (size_t)1 << (nq - 1 - q), mr = mc << nq;
        int p = ((e & mr) ? 2 : 0) | ((e & mc) ? 1 : 0);
        const cplx* T = &so.T[q * 16 + p * 4];
        acc = acc + T[p] * in[idx] + T[p ^ 2] * in[idx ^ mr] + T[p ^ 1] * in[idx ^ mc] +
              T[p ^ 3] * in[idx ^ mr ^ mc];
      }
      out[idx] = acc;
    }
    return 1;
  }
  int scaled_sumsq(const Geometry& g, double* out, const amp_t* x, const amp_t* xsub, const amp_t* ref,
                   double atol, double rtol, double*, void*) {
    for (int b = 0; b < g.batch; ++b) {
      double acc = 0;
      for (size_t i = 0; i < g.dim; ++i) {
        size_t k = (size_t)b * g.dim + i;
        cplx v = xsub ? cplx(x[k]) - cplx(xsub[k]) : cplx(x[k]);
        double sc = atol + rtol * std::hypot(ref[k].re, ref[k].im);
        acc += (v.re / sc) * (v.re / sc) + (v.im / sc) * (v.im / sc);
      }
      out[b] = acc;
    }
    return 1;
  }
  int err_sumsq(const Geometry& g, double* out, const amp_t* const* k, const double* ew,
                const amp_t* y0, const amp_t* y1, double atol, double rtol, double*, void*) {
    for (int b = 0; b < g.batch; ++b) {
      double acc = 0;
      for (size_t i = 0; i < g.dim; ++i) {
        size_t x = (size_t)b * g.dim + i;
        double er = 0, ei = 0;
        for (int j = 0; j < 7; ++j)
          if (ew[j] != 0.0) { er += ew[j] * k[j][x].re; ei += ew[j] * k[j][x].im; }
        double sc = atol + rtol * std::max(std::hypot(y0[x].re, y0[x].im), std::hypot(y1[x].re, y1[x].im));
        acc += (er / sc) * (er / sc) + (ei / sc) * (ei / sc);
      }
      out[b] = acc;
    }
    return 1;
  }
  int corr(const Geometry& g, cplx* d_corr, double* d_wacc, double wscale, const amp_t* kbar,
           const amp_t* y, double*, void*) {
    int nq = g.nq;
    size_t S = (size_t)1 << nq;
    int per = g.kind == PD_KET ? 4 : 16;
    if (d_corr) for (int i = 0; i < nq * per; ++i) d_corr[i] = {0, 0};
    for (size_t idx = 0; idx < g.dim * g.batch; ++idx) {
      size_t e = idx & (g.dim - 1);
      cplx kb = conj(kbar[idx]);
      cplx self = kb * y[idx];
      if (d_wacc) {
        if (g.kind == PD_KET) d_wacc[e] += wscale * self.im;
        else { d_wacc[e >> nq] += wscale * self.im; d_wacc[e & (S - 1)] -= wscale * self.im; }
      }
      if (!d_corr) continue;
      for (int q = 0; q < nq; ++q) {
        if (g.kind == PD_KET) {
          size_t m = (size_t)1 << (nq - 1 - q);
          int a = (e & m) ? 1 : 0;
          d_corr[q * 4 + a * 2 + a] = d_corr[q * 4 + a * 2 + a] + self;
          d_corr[q * 4 + a * 2 + (1 - a)] = d_corr[q * 4 + a * 2 + (1 - a)] + kb * y[idx ^ m];
        } else {
          size_t mc = (size_t)1 << (nq - 1 - q), mr = mc << nq;
          int p = ((e & mr) ? 2 : 0) | ((e & mc) ? 1 : 0);
          cplx* C = &d_corr[q * 16 + p * 4];
          C[p] = C[p] + self;
          C[p ^ 2] = C[p ^ 2] + kb * y[idx ^ mr];
          C[p ^ 1] = C[p ^ 1] + kb * y[idx ^ mc];
          C[p ^ 3] = C[p ^ 3] + kb * y[idx ^ mr ^ mc];
        }
      }
    }
    return 1;
  }
  int corr_combo(const Geometry& g, cplx* d_corr, double* d_wacc, double wscale, const amp_t* kbar, int n_in,
                 const amp_t* const* ins, const double* w, amp_t* ybuf, double* scratch, void* s) {
    int n = 0;
    const amp_t* ysrc = ins[0];
    if (n_in > 1 || w[0] != 1.0) {
      n += lincomb(g, ybuf, n_in, ins, w, s);
      ysrc = ybuf;
    }
    return n + corr(g, d_corr, d_wacc, wscale, kbar, ysrc, scratch, s);
  }
  int re_dot(const Geometry& g, double* out, const amp_t* a, const amp_t* b, double*, void*) {
    double acc = 0;
    for (size_t i = 0; i < g.dim * g.batch; ++i) acc += a[i].re * b[i].re + a[i].im * b[i].im;
    *out = acc;
    return 1;
  }
  int pair_reduce(const Geometry& g, double* d_pair, const double* d_wacc, void*) {
    int nq = g.nq;
    size_t S = (size_t)1 << nq;
    for (int i = 0; i < nq; ++i)
      for (int j = 0; j < nq; ++j) {
        double acc = 0;
        if (i < j)
          for (size_t s = 0; s < S; ++s)
            if (!(s >> (nq - 1 - i) & 1) && !(s >> (nq - 1 - j) & 1)) acc += d_wacc[s];
        d_pair[i * nq + j] = acc;
      }
    return 1;
  }
  int sharded_accumulate(const Geometry& g, amp_t* out, const amp_t* psi, double shift, int n_peers,
                         const amp_t* const* peers, const cplx* coef, void*) {
    size_t L = g.dim * g.batch;
    for (size_t i = 0; i < L; ++i) {
      double re = out[i].re + shift * psi[i].re, im = out[i].im + shift * psi[i].im;
      for (int k = 0; k < n_peers; ++k) {
        re += coef[k].re * peers[k][i].re - coef[k].im * peers[k][i].im;
        im += coef[k].re * peers[k][i].im + coef[k].im * peers[k][i].re;
      }
      out[i] = cplx{re, im};
    }
    return 1;
  }
  int expect_diag(const Geometry& g, const amp_t* states, int n_t, const double* obs, cplx* out,
                  double*, void*) {
    size_t S = (size_t)1 << g.nq;
    for (int t = 0; t < n_t; ++t) {
      cplx acc{0, 0};
      const amp_t* st = states + (size_t)t * g.dim * g.batch;
      for (int b = 0; b < g.batch; ++b)
        if (g.kind == PD_KET)
          for (size_t s = 0; s < g.dim; ++s) {
            cplx v = st[(size_t)b * g.dim + s];
            acc.re += obs[s] * (v.re * v.re + v.im * v.im);
          }
        else
          for (size_t r = 0; r < S; ++r) acc = acc + obs[r] * st[(size_t)b * g.dim + r * S + r];
      out[t] = acc;
    }
    return 1;
  }
};

}  // namespace pd
