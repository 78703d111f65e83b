// CUDA backend (sm_100a) for the evolution engine: memory plumbing + kernel launchers.
//
// Kernel families:
//   * "gather" (gather_kernels.cu): one thread per amplitude, bit-flip partners fetched through
//     L1/L2.  Any N, any addressing, ket and density.  Correctness baseline; the path while the
//     working set is cache resident, and the Lindblad path.
//   * "small" (small_ket*.cu): kets of N <= 14 -- the whole adaptive evolution / adjoint sweep as one
//     cooperative kernel with register-resident lanes and a flag-in-data exchange through L2.
//   * "tiled" (tiled_ket.cu): kets of N = 18..23 on request (path 2) -- two tile types, fused finalise/start
//     launches; the default until the stream family overtook it in round 2.
//   * "stream" (stream_ket.cu): kets of N >= 19 -- one bit-group of H per tile type, >= 256 B pieces, A tiles +
//     first group as one L2-blocked dataflow launch; its tiled correlation kernels serve the adjoint sweep.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>

#include "pd_common.hpp"

namespace pd {

#define PD_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      throw Error(PD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

// Host<->device copies of the library go through these two wrappers, which count the bytes
// (C ABI pd_transfer_counters; bench.py reports them as h2d/d2h bytes per step).
struct XferCounters {
  std::atomic<long long> h2d{0}, d2h{0};
};
inline XferCounters& xfer_counters() {
  static XferCounters c;
  return c;
}
inline cudaError_t copy_h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  xfer_counters().h2d += (long long)bytes;
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
}
inline cudaError_t copy_d2h(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  xfer_counters().d2h += (long long)bytes;
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
}

// Buffers that may live on either side (the per-unit coefficient tables and their gradients of the
// *_units entry points): only host<->device traffic is counted.
inline bool is_device_pointer(const void* p) {
  cudaPointerAttributes a;
  if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
inline cudaError_t copy_in(void* dst_dev, const void* src, size_t bytes, cudaStream_t s) {
  if (is_device_pointer(src)) return cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyDeviceToDevice, s);
  return copy_h2d(dst_dev, src, bytes, s);
}
inline cudaError_t copy_out(void* dst, const void* src_dev, size_t bytes, cudaStream_t s) {
  if (is_device_pointer(dst)) return cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToDevice, s);
  return copy_d2h(dst, src_dev, bytes, s);
}

constexpr int kThreads = 256;
constexpr int kMaxReduceBlocks = 148 * 4;
constexpr int kMaxR = 64;  // max doubles reduced per block in one pass

struct DensityT {  // SiteOpsDensity without the by-value size limit worries: lives in a param
  SiteOpsDensity so;
};

int launch_lincomb(const Geometry& g, amp_t* out, int n_in, const amp_t* const* ins, const double* w,
                   cudaStream_t s);
int launch_lincomb_c(size_t n, amp_t* out, int m, const amp_t* basis, size_t stride, const cplx* ws,
                     cudaStream_t s);
int launch_apply_ket(const Geometry& g, amp_t* out, const amp_t* in, const SiteOps& so, cudaStream_t s);
int launch_apply_density(const Geometry& g, amp_t* out, const amp_t* in, const SiteOpsDensity& so,
                         cudaStream_t s);
int launch_build_diag(double* diag, int nq, const double* d_pair_u, cudaStream_t s);
int launch_build_diag_parts(double* parts, int nq, const double* diag, cudaStream_t s);
int launch_scaled_sumsq(const Geometry& g, double* out, const amp_t* x, const amp_t* xsub,
                        const amp_t* ref, double atol, double rtol, double* scratch, cudaStream_t s);
int launch_err_sumsq(const Geometry& g, double* out, const amp_t* const* k, const double* ew,
                     const amp_t* y0, const amp_t* y1, double atol, double rtol, double* scratch,
                     cudaStream_t s);
int launch_corr(const Geometry& g, cplx* d_corr, double* d_wacc, double wscale, const amp_t* kbar,
                const amp_t* y, double* scratch, cudaStream_t s);
int launch_re_dot(const Geometry& g, double* out, const amp_t* a, const amp_t* b, double* scratch,
                  cudaStream_t s);
int launch_pair_reduce(const Geometry& g, double* d_pair, const double* d_wacc, cudaStream_t s);
int launch_expect_diag(const Geometry& g, const amp_t* states, int n_t, const double* obs, cplx* out,
                       double* scratch, cudaStream_t s);
// sharded register (sharded_ket.cu): out += shift*psi + sum_k coef_k * peer_k, partner slices read
// in place from peer memory
int launch_sharded_accumulate(size_t n_amp, amp_t* out, const amp_t* psi, double shift, int n_peers,
                              const amp_t* const* peers, const cplx* coef, cudaStream_t s);
#if !defined(PD_C64)
// tiled family (tiled_ket.cu); returns 0 launches if the shape is not supported
bool tiled_ket_supported(const Geometry& g);
int launch_tiled_stage_ket(const Geometry& g, amp_t* out, amp_t* comb, int n_in,
                           const amp_t* const* ins, const double* w, const SiteOps& so, amp_t* tmp,
                           cudaStream_t s);

int launch_tiled_dp5_step(const Geometry& g, const amp_t* y, amp_t* const* k, amp_t* ynew,
                          const SiteOps* stage_ops, const double* beta, const double* b5,
                          const double* ew, double dt, double atol, double rtol, amp_t* tmp_a,
                          amp_t* tmp_b, double* err_partial, double* err_out, cudaStream_t s);
size_t tiled_err_partial_count(const Geometry& g);
#else
// The complex64 build (libpulser_diff_b200_c64.so) carries the bandwidth-bound families only: gather (any shape,
// ket and density), stream (kets of N >= 19) and the sharded accumulate.  The tiled kets, the density tiles and
// the small-register cooperative kernels exist in complex128; the Python layer serves registers of N <= 14 in
// complex64 mode through the complex128 kernels (nothing there is bandwidth-bound; ops.py).
inline bool tiled_ket_supported(const Geometry&) { return false; }
inline int launch_tiled_stage_ket(const Geometry&, amp_t*, amp_t*, int, const amp_t* const*, const double*,
                                  const SiteOps&, amp_t*, cudaStream_t) {
  throw Error(PD_ERR_STATE, "tiled ket kernels are not part of the complex64 build");
}
inline int launch_tiled_dp5_step(const Geometry&, const amp_t*, amp_t* const*, amp_t*, const SiteOps*, const double*,
                                 const double*, const double*, double, double, double, amp_t*, amp_t*, double*,
                                 double*, cudaStream_t) {
  throw Error(PD_ERR_STATE, "tiled ket kernels are not part of the complex64 build");
}
inline size_t tiled_err_partial_count(const Geometry&) { return 0; }
#endif
constexpr int kAutoTiledMinQubits = 19;   // up to N = 18 the working set is L2 resident and the gather kernels
                                          // win (N = 18: 0.119 ms per DP5 step against 0.135 stream, 0.159 tiled)
// stream family (stream_ket.cu): one bit-group of H per launch, >= 256 B pieces, any N >= 16
bool stream_ket_supported(const Geometry& g);
int launch_stream_stage_ket(const Geometry& g, amp_t* out, amp_t* ymat, int n_in, const amp_t* const* ins,
                            const double* w, const SiteOps& so, cudaStream_t s);
int launch_stream_dp5_step(const Geometry& g, const amp_t* y, amp_t* const* k, amp_t* ynew, amp_t* ymat, amp_t* aux,
                           const SiteOps* stage_ops, const double* beta, const double* ew, double dt,
                           double atol, double rtol, double* err_partial, double* err_out, cudaStream_t s);
size_t stream_err_partial_count(const Geometry& g);
// density tiles (dens_tile.cu): both bits of up to six sites closed per launch, N = 8..13
#if !defined(PD_C64)
bool dens_tile_supported(const Geometry& g);
int launch_dens_stage(const Geometry& g, amp_t* out, amp_t* ymat, int n_in, const amp_t* const* ins, const double* w,
                      const SiteOpsDensity& so, cudaStream_t s);
#else
inline bool dens_tile_supported(const Geometry&) { return false; }
inline int launch_dens_stage(const Geometry&, amp_t*, amp_t*, int, const amp_t* const*, const double*,
                             const SiteOpsDensity&, cudaStream_t) {
  throw Error(PD_ERR_STATE, "density tile kernels are not part of the complex64 build");
}
#endif
int launch_stream_corr(const Geometry& g, cplx* d_corr, double* d_wacc, double wscale, const amp_t* kbar, int n_in,
                       const amp_t* const* ins, const double* w, amp_t* ymat, cudaStream_t s);
#if !defined(PD_C64)
// small-register family (small_ket*.cu): whole forward / adjoint sweep in one cooperative kernel
struct SmallKetState;
SmallKetState* small_ket_create();
void small_ket_destroy(SmallKetState*);
bool small_ket_supported(const Geometry& g, const Program& prog);
bool small_ket_units_supported(const Geometry& g, const Program& prog);
int small_ket_lanczos(SmallKetState& S, const Geometry& g1, const Program& prog, const amp_t* v0, amp_t* basis,
                      int max_m, double t_eval, int j0, int j1, double beta_prev, double* alpha_host,
                      double* beta_host, double* nrm, cudaStream_t st);
void small_ket_unit_counts(SmallKetState& S, uint64_t tape_gen, int unit, int* accepted, int* attempts);
int small_ket_backward_units(SmallKetState& S, const Geometry& g, const Program& prog, const Tableau& tab,
                             const std::vector<double>& tsave, int n_units, const double* dv, const double* av,
                             uint64_t tape_gen, const amp_t* gstates, amp_t* lam_out, double* g_det, double* g_amp,
                             cudaStream_t st);
int small_ket_forward(SmallKetState& S, const Geometry& g, const Program& prog, const Tableau& tab,
                      const pd_options& o, int n_units, const amp_t* y0, const double* dv, const double* av,
                      const double* tsave, int n_t, amp_t* states,
                      std::vector<std::vector<pd_step_record>>& records, bool want_tape,
                      uint64_t* tape_gen_out, cudaStream_t st);
int small_ket_backward(SmallKetState& S, const Geometry& g, const Program& prog, const Tableau& tab,
                       const std::vector<double>& tsave, int n_units, const double* dv, const double* av,
                       const std::vector<std::vector<SkStepHost>>& steps, uint64_t tape_gen,
                       const amp_t* gstates, bool want_coef, double* d_wacc, amp_t* lam_out,
                       std::vector<std::vector<double>>& slot_sums, cudaStream_t st);

#else
struct SmallKetState {};
inline SmallKetState* small_ket_create() { return new SmallKetState(); }
inline void small_ket_destroy(SmallKetState* p) { delete p; }
inline bool small_ket_supported(const Geometry&, const Program&) { return false; }
inline bool small_ket_units_supported(const Geometry&, const Program&) { return false; }
inline int small_ket_lanczos(SmallKetState&, const Geometry&, const Program&, const amp_t*, amp_t*, int, double, int,
                             int, double, double*, double*, double*, cudaStream_t) {
  throw Error(PD_ERR_STATE, "small-register kernels are not part of the complex64 build");
}
inline void small_ket_unit_counts(SmallKetState&, uint64_t, int, int* a, int* b) { *a = *b = -1; }
inline int small_ket_backward_units(SmallKetState&, const Geometry&, const Program&, const Tableau&,
                                    const std::vector<double>&, int, const double*, const double*, uint64_t,
                                    const amp_t*, amp_t*, double*, double*, cudaStream_t) {
  return 0;
}
inline int small_ket_forward(SmallKetState&, const Geometry&, const Program&, const Tableau&, const pd_options&, int,
                             const amp_t*, const double*, const double*, const double*, int, amp_t*,
                             std::vector<std::vector<pd_step_record>>&, bool, uint64_t*, cudaStream_t) {
  throw Error(PD_ERR_STATE, "small-register kernels are not part of the complex64 build");
}
inline int small_ket_backward(SmallKetState&, const Geometry&, const Program&, const Tableau&,
                              const std::vector<double>&, int, const double*, const double*,
                              const std::vector<std::vector<SkStepHost>>&, uint64_t, const amp_t*, bool, double*,
                              amp_t*, std::vector<std::vector<double>>&, cudaStream_t) {
  return 0;
}
#endif

class CudaBackend {
 public:
  static constexpr bool is_cuda = true;
  static bool on_device(const void* p) { return is_device_pointer(p); }
  static void transfer_counters(long long* h2d, long long* d2h, bool reset) {
    XferCounters& c = xfer_counters();
    if (h2d) *h2d = c.h2d.load();
    if (d2h) *d2h = c.d2h.load();
    if (reset) { c.h2d = 0; c.d2h = 0; }
  }
  // device guard of the C ABI (cabi_impl.hpp DeviceScope): make `dev` current, return the previous device
  static int push_device(int dev) {
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    return prev;
  }
  static void pop_device(int prev) {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
  int device;
  int path = 0;  // 0 auto, 1 gather, 2 tiled
  explicit CudaBackend(int dev) : device(dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      throw Error(PD_ERR_CUDA, "pulser_diff_b200 needs a CUDA device (no CPU fallback): " +
                                   std::string(cudaGetErrorString(e)));
    if (dev < 0 || dev >= n) throw Error(PD_ERR_INVALID, "bad CUDA device ordinal");
    PD_CUDA_CHECK(cudaMalloc(&d_pair_u_, sizeof(double) * kMaxQubits * kMaxQubits));
  }
  ~CudaBackend() {
    cudaFree(d_pair_u_);
    if (small_) small_ket_destroy(small_);
  }
  bool small_supported(const Geometry& g, const Program& prog) { return small_ket_supported(g, prog); }
  bool small_units_supported(const Geometry& g, const Program& prog) { return small_ket_units_supported(g, prog); }
  int small_lanczos(const Geometry& g1, const Program& prog, const amp_t* v0, amp_t* basis, int max_m, double t_eval,
                    int j0, int j1, double beta_prev, double* alpha_host, double* beta_host, double* nrm, void* s) {
    if (!small_) small_ = small_ket_create();
    return small_ket_lanczos(*small_, g1, prog, v0, basis, max_m, t_eval, j0, j1, beta_prev, alpha_host, beta_host,
                             nrm, st(s));
  }
  int small_forward(const Geometry& g, const Program& prog, const Tableau& tab, const pd_options& o,
                    int n_units, const amp_t* y0, const double* dv, const double* av, const double* tsave,
                    int n_t, amp_t* states, std::vector<std::vector<pd_step_record>>& recs, bool want_tape,
                    uint64_t* gen, void* s) {
    if (!small_) small_ = small_ket_create();
    return small_ket_forward(*small_, g, prog, tab, o, n_units, y0, dv, av, tsave, n_t, states, recs,
                             want_tape, gen, st(s));
  }
  int small_backward_units(const Geometry& g, const Program& prog, const Tableau& tab,
                           const std::vector<double>& tsave, int n_units, const double* dv, const double* av,
                           uint64_t tape_gen, const amp_t* gstates, amp_t* lam_out, double* g_det, double* g_amp,
                           void* s) {
    if (!small_) small_ = small_ket_create();
    return small_ket_backward_units(*small_, g, prog, tab, tsave, n_units, dv, av, tape_gen, gstates, lam_out,
                                    g_det, g_amp, st(s));
  }
  void small_unit_counts(uint64_t tape_gen, int unit, int* acc, int* att) {
    *acc = *att = -1;
    if (small_) small_ket_unit_counts(*small_, tape_gen, unit, acc, att);
  }
  int small_backward(const Geometry& g, const Program& prog, const Tableau& tab,
                     const std::vector<double>& tsave, int n_units, const double* dv, const double* av,
                     const std::vector<std::vector<SkStepHost>>& steps, uint64_t tape_gen,
                     const amp_t* gstates, bool want_coef, double* d_wacc, amp_t* lam_out,
                     std::vector<std::vector<double>>& sums, void* s) {
    if (!small_) small_ = small_ket_create();
    return small_ket_backward(*small_, g, prog, tab, tsave, n_units, dv, av, steps, tape_gen, gstates,
                              want_coef, d_wacc, lam_out, sums, st(s));
  }
  void* alloc(size_t bytes) {
    void* p = nullptr;   // the C ABI entry point has made `device` current (DeviceScope)
    PD_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    return p;
  }
  void free(void* p) { if (p) cudaFree(p); }
  static cudaStream_t st(void* s) { return (cudaStream_t)s; }
  void zero(void* p, size_t bytes, void* s) { PD_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st(s))); }
  void d2d(void* d, const void* sr, size_t b, void* s) {
    PD_CUDA_CHECK(cudaMemcpyAsync(d, sr, b, cudaMemcpyDeviceToDevice, st(s)));
  }
  void d2h(void* d, const void* sr, size_t b, void* s) {
    PD_CUDA_CHECK(copy_d2h(d, sr, b, st(s)));
  }
  void sync(void* s) { PD_CUDA_CHECK(cudaStreamSynchronize(st(s))); }
  void timer_start(void* s) {
    if (!ev0_) { PD_CUDA_CHECK(cudaEventCreate(&ev0_)); PD_CUDA_CHECK(cudaEventCreate(&ev1_)); }
    PD_CUDA_CHECK(cudaEventRecord(ev0_, st(s)));
  }
  double timer_stop_ms(void* s) {
    PD_CUDA_CHECK(cudaEventRecord(ev1_, st(s)));
    PD_CUDA_CHECK(cudaEventSynchronize(ev1_));
    float ms = 0.f;
    PD_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    return (double)ms;
  }
  size_t reduce_scratch_bytes(const Geometry&) { return sizeof(double) * kMaxReduceBlocks * kMaxR * 2; }
  size_t segment_budget_bytes() {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return (size_t)1 << 30;
    return std::max<size_t>(fr / 3, (size_t)64 << 20);
  }

  void build_diag(double* diag, int nq, const double* pair_u_host, void* s) {
    PD_CUDA_CHECK(copy_h2d(d_pair_u_, pair_u_host, sizeof(double) * nq * nq, st(s)));
    launch_build_diag(diag, nq, d_pair_u_, st(s));
  }
  void build_diag_parts(double* parts, int nq, const double* diag, void* s) {
    launch_build_diag_parts(parts, nq, diag, st(s));
  }
  int lincomb(const Geometry& g, amp_t* out, int n_in, const amp_t* const* ins, const double* w, void* s) {
    return launch_lincomb(g, out, n_in, ins, w, st(s));
  }
  int lincomb_c(const Geometry& g, amp_t* out, int m, const amp_t* basis, size_t stride,
                const cplx* ws_host, void* s) {
    return launch_lincomb_c(g.dim * g.batch, out, m, basis, stride, ws_host, st(s));
  }
  int stage_ket(const Geometry& g, amp_t* out, amp_t* comb, int n_in, const amp_t* const* ins,
                const double* w, const SiteOps& so, amp_t* scratch, void* s) {
    if (use_stream(g)) {
      const bool plain = n_in == 1 && w[0] == 1.0 && comb == nullptr;
      return launch_stream_stage_ket(g, out, plain ? nullptr : (comb ? comb : scratch), n_in, ins, w, so, st(s));
    }
    if (use_tiled(g))
      return launch_tiled_stage_ket(g, out, comb, n_in, ins, w, so, scratch, st(s));
    int n = 0;
    const amp_t* src = ins[0];
    if (n_in > 1 || w[0] != 1.0) {
      amp_t* dst = comb ? comb : scratch;
      n += launch_lincomb(g, dst, n_in, ins, w, st(s));
      src = dst;
    } else if (comb) {
      d2d(comb, ins[0], sizeof(amp_t) * g.dim * g.batch, s);
    }
    return n + launch_apply_ket(g, out, src, so, st(s));
  }
  // Automatic family for large kets, from the measured DP5 step times: since round 2 the stream family wins
  // at every N >= 19 (N = 21 / 22 / 23: 0.554 / 0.927 / 1.667 ms per step against 0.587 / 1.090 / 2.080 ms
  // tiled; profiles/r02_stream_n26.md).  path 2 / 4 force one of them.
  static bool stream_preferred(int nq) { return nq >= kAutoTiledMinQubits; }
  bool use_tiled(const Geometry& g) const {
    if (path == 1 || path == 4 || !tiled_ket_supported(g)) return false;
    if (path == 2) return true;
    return g.nq >= kAutoTiledMinQubits && !(stream_preferred(g.nq) && stream_ket_supported(g));
  }
  bool use_stream(const Geometry& g) const {
    if (!stream_ket_supported(g)) return false;
    if (path == 4) return true;
    return path == 0 && g.nq >= kAutoTiledMinQubits && (stream_preferred(g.nq) || !tiled_ket_supported(g));
  }
  // One full Dormand-Prince step with the alternating tiled kernels; 0 = not handled here.
  int dp5_step_ket(const Geometry& g, const amp_t* y, amp_t* const* k, amp_t* ynew,
                   const SiteOps* stage_ops, const Tableau& tab, const double* ew, double dt,
                   double atol, double rtol, amp_t* tmp_a, amp_t* tmp_b, double* red_scratch,
                   double* err_out, void* s) {
    if (use_stream(g)) {
      // ew holds dt*(b5-b4); the stream step wants the weights of the slopes directly
      if (stream_err_partial_count(g) > (size_t)kMaxReduceBlocks * kMaxR * 2) return 0;
      return launch_stream_dp5_step(g, y, k, ynew, tmp_a, tmp_b, stage_ops, &tab.beta[0][0], ew, dt, atol, rtol,
                                    red_scratch, err_out, st(s));
    }
    if (!use_tiled(g)) return 0;
    if (tiled_err_partial_count(g) > (size_t)kMaxReduceBlocks * kMaxR * 2) return 0;
    return launch_tiled_dp5_step(g, y, k, ynew, stage_ops, &tab.beta[0][0], tab.b5, ew, dt, atol,
                                 rtol, tmp_a, tmp_b, red_scratch, err_out, st(s));
  }
  // The density tiles are opt-in (path 5): measured at N = 12 they are SLOWER than the gather kernel (DP5_ME
  // step 6.3-6.7 ms with 2-3 tile launches per application, 7.5 ms with one launch that gathers the sites
  // outside the tile, against 5.6 ms; profiles/r02_lindblad.md) -- the gather kernel runs at the L2 gather
  // rate with full occupancy, the tile kernels serialise a load phase and an instruction-heavy compute phase
  // at two CTAs per SM.
  bool use_dens_tiles(const Geometry& g) const { return path == 5 && dens_tile_supported(g); }
  int stage_density(const Geometry& g, amp_t* out, amp_t* comb, int n_in, const amp_t* const* ins,
                    const double* w, const SiteOpsDensity& so, amp_t* scratch, void* s) {
    if (use_dens_tiles(g)) {
      const bool plain = n_in == 1 && w[0] == 1.0 && comb == nullptr;
      return launch_dens_stage(g, out, plain ? nullptr : (comb ? comb : scratch), n_in, ins, w, so, st(s));
    }
    int n = 0;
    const amp_t* src = ins[0];
    if (n_in > 1 || w[0] != 1.0) {
      amp_t* dst = comb ? comb : scratch;
      n += launch_lincomb(g, dst, n_in, ins, w, st(s));
      src = dst;
    } else if (comb) {
      d2d(comb, ins[0], sizeof(amp_t) * g.dim * g.batch, s);
    }
    return n + launch_apply_density(g, out, src, so, st(s));
  }
  int scaled_sumsq(const Geometry& g, double* out, const amp_t* x, const amp_t* xsub, const amp_t* ref,
                   double atol, double rtol, double* scratch, void* s) {
    return launch_scaled_sumsq(g, out, x, xsub, ref, atol, rtol, scratch, st(s));
  }
  int err_sumsq(const Geometry& g, double* out, const amp_t* const* k, const double* ew,
                const amp_t* y0, const amp_t* y1, double atol, double rtol, double* scratch, void* s) {
    return launch_err_sumsq(g, out, k, ew, y0, y1, atol, rtol, scratch, st(s));
  }
  int corr(const Geometry& g, cplx* d_corr, double* d_wacc, double wscale, const amp_t* kbar,
           const amp_t* y, double* scratch, void* s) {
    return launch_corr(g, d_corr, d_wacc, wscale, kbar, y, scratch, st(s));
  }
  // correlations of kbar with the stage input y = sum_j w_j in_j; ybuf receives y when it has to be formed
  int corr_combo(const Geometry& g, cplx* d_corr, double* d_wacc, double wscale, const amp_t* kbar, int n_in,
                 const amp_t* const* ins, const double* w, amp_t* ybuf, double* scratch, void* s) {
    // the tiled correlation kernels serve both large-register families (any N >= 16)
    if ((use_stream(g) || use_tiled(g)) && stream_ket_supported(g))
      return launch_stream_corr(g, d_corr, d_wacc, wscale, kbar, n_in, ins, w, ybuf, st(s));
    int n = 0;
    const amp_t* ysrc = ins[0];
    if (n_in > 1 || w[0] != 1.0) {
      n += launch_lincomb(g, ybuf, n_in, ins, w, st(s));
      ysrc = ybuf;
    }
    return n + launch_corr(g, d_corr, d_wacc, wscale, kbar, ysrc, scratch, st(s));
  }
  int re_dot(const Geometry& g, double* out, const amp_t* a, const amp_t* b, double* scratch, void* s) {
    return launch_re_dot(g, out, a, b, scratch, st(s));
  }
  int pair_reduce(const Geometry& g, double* d_pair, const double* d_wacc, void* s) {
    return launch_pair_reduce(g, d_pair, d_wacc, st(s));
  }
  int expect_diag(const Geometry& g, const amp_t* states, int n_t, const double* obs, cplx* out,
                  double* scratch, void* s) {
    return launch_expect_diag(g, states, n_t, obs, out, scratch, st(s));
  }
  int sharded_accumulate(const Geometry& g, amp_t* out, const amp_t* psi, double shift, int n_peers,
                         const amp_t* const* peers, const cplx* coef, void* s) {
    return launch_sharded_accumulate(g.dim * g.batch, out, psi, shift, n_peers, peers, coef, st(s));
  }

 private:
  double* d_pair_u_ = nullptr;
  SmallKetState* small_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
};

}  // namespace pd
