// Small-register ket kernels: the launch-bound regime (N <= 14, state resident in registers/L2).
//
// For a 12-qubit register one state vector is 64 KiB: stage-by-stage launches spend their time in
// launch latency and host round trips (98k launches / 90 ms for one forward+gradient pass of the
// C2 workload, profiles/r01_launch_shares_bench.md).  Here the WHOLE adaptive Dormand-Prince
// evolution (reference call: pyqtorch.sesolve behind backend.py:488-494; controller of SURVEY.md
// Appendix A.3) and the WHOLE discrete-adjoint sweep (reference: the autograd tape,
// derivative.py:40,76) each run as ONE cooperative kernel launch:
//
//   * a vector of L complex amplitudes is 2L real "lanes"; thread r owns double r of every vector
//     (y, k1..k7 forward; lambda and the six stage adjoints backward) in REGISTERS across stages
//     and steps, so every Runge-Kutta combination is a scalar FMA chain on private data;
//   * a stage input is published to an L2-resident exchange buffer as flag-in-data lines and the
//     N bit-flip partners are polled from it: no kernel boundary, no barrier and no memory fence
//     between stages (a release fence alone costs ~2 us here); one exchange = one L2 round trip;
//   * many small CTAs (128 threads, one warp per scheduler) spread the lanes over the chip: the
//     per-stage critical path is one warp's ~100 instructions plus the L2 round trip;
//   * the pulse coefficients d_q(t), g_q(t) are interpolated on the device at every stage time
//     with the reference's rule (hamiltonian.py:532-542, quirk included);
//   * the step controller runs redundantly in every thread on the same reduced error norm, so
//     accept/reject decisions are uniform without a broadcast; the attempted-step log is written
//     by one thread and becomes the host tape (pd_step_record);
//   * the forward sweep records, per accepted step, the six stage inputs and slopes on a device
//     tape (what the reference's autograd tape stores); the adjoint sweep reads it back instead of
//     recomputing, and reduces per stage the per-term gradient sums for the host to scatter onto
//     the sample arrays.
#pragma once
#include "cuda_backend.cuh"

namespace pd {
namespace sk {

constexpr int SK_T = 128;          // threads per CTA: one warp per scheduler (64 and 256 measured slower)
constexpr int SK_MAXQ = 16;        // qubits handled by this family
constexpr int SK_MAXTERMS = 24;    // n_det, n_amp each
constexpr int SK_MAXB = 32;        // batch columns
constexpr int SK_MAXC = 296;       // co-resident CTAs of one cooperative launch (two per SM)

struct SkProg {
  int nq, n_samples, n_det, n_amp;
  double dt;
  const unsigned long long* det_masks;
  const double* det_values;   // [n_det][n_samples]
  const unsigned long long* amp_masks;
  const double* amp_values;   // [n_amp][n_samples][2]
  const double* diag;         // Dint[2^nq]
};

struct SkTab {
  double alpha[6];
  double beta[6][6];
  double b5[7];
  double eb[7];   // b5 - b4
};

// ---- flag-in-data exchange ("LL" lines) -----------------------------------------------------
// Every published double travels as one 16-byte line {lo, seq, hi, seq}: the 32-bit sequence number
// of the exchange phase is interleaved with the payload, each (payload, flag) pair is an aligned
// 8-byte unit and therefore never torn, and a reader simply re-reads until both flags carry the
// phase it waits for.  Lines are double-buffered by phase parity: a lane's slot is only read by
// lanes of its partner elements, and when the element has received phase n+1 from all of them
// they have all consumed its phase n line, so the slot can take phase n+2.
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned seq) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(seq), "r"(hi), "r"(seq)
               : "memory");
}
__device__ __forceinline__ uint4 ll_load(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ll_bad(const uint4& v, unsigned seq) { return (v.y ^ seq) | (v.w ^ seq); }
__device__ __forceinline__ double ll_value(const uint4& v) { return __hiloint2double((int)v.z, (int)v.x); }

// Bounded polling: a lane that waits longer than ~seconds raises the launch-wide abort flag, and
// every poll loop leaves as soon as it sees the flag, so a protocol error surfaces as an error code
// instead of a hung device.
struct SkPoll {
  int* abort_flag;
  unsigned spins = 0;
  __device__ __forceinline__ bool give_up() {
    if ((++spins & 1023u) != 0) return false;
    if (*(volatile int*)abort_flag) return true;
    if (spins > (1u << 24)) { *(volatile int*)abort_flag = 1; return true; }
    return false;
  }
};

// per-bit coefficients of one stage time in shared memory (bit position p = nq-1-q)
struct SkCoef {
  double d[SK_MAXQ];
  double gre[SK_MAXQ];
  double gim[SK_MAXQ];
  int uniform;      // every qubit carries the same d and g (global channels only)
  int pad;
};

// Thread (stage i, qubit q) evaluates the reference interpolation rule (hamiltonian.py:532-542)
// at that stage's time; all stage times of a step are known when the step starts.
__device__ __forceinline__ void sk_eval_one(const SkProg& P, double t, SkCoef* c, int q) {
  const int ns = P.n_samples;
  double d = 0.0, gre = 0.0, gim = 0.0;
  if (ns >= 2) {
    const double fl = floor(t / P.dt);
    long i1 = (long)fmin(fl, (double)(ns - 2));
    if (i1 < 0) i1 = 0;
    long i2 = i1 + 1 < (long)(ns - 2) ? i1 + 1 : (long)(ns - 2);
    if (i2 < 0) i2 = 0;
    const double x = (t - i1 * P.dt) / P.dt;
    for (int k = 0; k < P.n_det; ++k)
      if (P.det_masks[k] >> q & 1ull) {
        const double* v = P.det_values + (size_t)k * ns;
        const double cc = v[i1] + (v[i2] - v[i1]) * x;
        d += cc + cc;
      }
    for (int k = 0; k < P.n_amp; ++k)
      if (P.amp_masks[k] >> q & 1ull) {
        const double* v = P.amp_values + (size_t)k * ns * 2;
        gre += v[2 * i1] + (v[2 * i2] - v[2 * i1]) * x;
        gim += v[2 * i1 + 1] + (v[2 * i2 + 1] - v[2 * i1 + 1]) * x;
      }
  }
  const int p = P.nq - 1 - q;
  c->d[p] = d;
  c->gre[p] = gre;
  c->gim[p] = gim;
}
// times[i] for i in [0, 6): coefficient set i.  Two __syncthreads inside.
__device__ __forceinline__ void sk_eval_stages(const SkProg& P, const double* times, SkCoef* c, int tid) {
  const int nthr = blockDim.x;
  for (int w = tid; w < 6 * P.nq; w += nthr) sk_eval_one(P, times[w / P.nq], &c[w / P.nq], w % P.nq);
  __syncthreads();
  if (tid < 6) {
    int u = 1;
    for (int p = 1; p < P.nq; ++p)
      u &= (c[tid].d[p] == c[tid].d[0]) & (c[tid].gre[p] == c[tid].gre[0]) & (c[tid].gim[p] == c[tid].gim[0]);
    c[tid].uniform = u;
  }
  __syncthreads();
}

// Copies the pulse tables into shared memory when they fit (the per-stage interpolation then costs
// shared-memory latency instead of dependent L2 round trips) and repoints the program at them.
constexpr int SK_TABLE_BYTES = 32 * 1024;
// The table space is the launch's dynamic shared memory (sized by the host from the same formula, 0 when the
// tables do not fit): a unit of a batch then costs a few KiB of shared memory instead of a fixed 32 KiB.
__device__ __forceinline__ void sk_cache_tables(SkProg& P, unsigned char* tab_smem, int tid) {
  const size_t ndv = (size_t)P.n_det * P.n_samples, nav = (size_t)P.n_amp * P.n_samples * 2;
  const size_t need = (ndv + nav + P.n_det + P.n_amp) * 8;
  unsigned cap;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(cap));
  if (need > SK_TABLE_BYTES || need > cap) return;    // uniform
  const int SK_TN = blockDim.x;
  double* dv = reinterpret_cast<double*>(tab_smem);
  double* av = dv + ndv;
  unsigned long long* dm = reinterpret_cast<unsigned long long*>(av + nav);
  unsigned long long* am = dm + P.n_det;
  for (size_t i = tid; i < ndv; i += SK_TN) dv[i] = P.det_values[i];
  for (size_t i = tid; i < nav; i += SK_TN) av[i] = P.amp_values[i];
  for (int i = tid; i < P.n_det; i += SK_TN) dm[i] = P.det_masks[i];
  for (int i = tid; i < P.n_amp; i += SK_TN) am[i] = P.amp_masks[i];
  __syncthreads();
  P.det_values = dv; P.amp_values = av; P.det_masks = dm; P.amp_masks = am;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- one lane of one application of H ---------------------------------------------------------
// Lane r = 2*element + part (part 0 = real, 1 = imaginary).  The two lanes of an element split the
// partners by parity of the bit position (NH = NQG/2 partners each): a lane polls its partner
// lines, accumulates BOTH parts of H Y over them and the halves are combined with one shuffle.
//
//   part 0 gets (-i H Y).re =  (H Y).im ;  part 1 gets (-i H Y).im = -(H Y).re
//   (+i H Y, the adjoint generator, is the negative of both.)
//
// Thread constants (SkLane): byte offsets of the partner lines and the bit pattern of the element.
template <int NH>
struct SkLane {
  unsigned eoff;          // byte offset of the element's pair of lines (32 B per element)
  unsigned xm[NH];        // XOR masks on eoff selecting partner j (bit position 2j + part); 0 = no such qubit
  double wv[NH];          // 1 for a real partner, 0 for a padding slot (which re-reads the own lines)
  double sg[NH];          // +1 if the element's bit is set (ground), -1 if clear (Rydberg), 0 for padding
  unsigned bits;          // basis-state bits of the element
  int part;
  int nzero;              // number of zero (Rydberg) bits: multiplies a uniform detuning
  double dint;
};
template <int NH>
__device__ __forceinline__ void sk_lane_init(SkLane<NH>& ln, size_t e, size_t dim, int nq, int part, double dint) {
  ln.eoff = (unsigned)e * 32u;
  ln.bits = (unsigned)(e & (dim - 1));
  ln.part = part;
#pragma unroll
  for (int j = 0; j < NH; ++j) {
    const int p = 2 * j + part;
    const bool real = p < nq;
    ln.xm[j] = real ? 32u << p : 0u;
    ln.wv[j] = real ? 1.0 : 0.0;
    ln.sg[j] = real ? (((ln.bits >> p) & 1) ? 1.0 : -1.0) : 0.0;
  }
  ln.nzero = nq - __popc(ln.bits);
  ln.dint = dint;
}
// partner-line pointers of one exchange buffer (thread constants; the loads then need no address math)
template <int NH>
struct SkPtrs {
  const char* p[NH];
  uint4* own;
};
template <int NH>
__device__ __forceinline__ void sk_ptrs_init(SkPtrs<NH>& o, const SkLane<NH>& ln, uint4* buf, size_t r) {
  const char* base = reinterpret_cast<const char*>(buf);
#pragma unroll
  for (int j = 0; j < NH; ++j) o.p[j] = base + (ln.eoff ^ ln.xm[j]);
  o.own = buf + r;
}
template <int NH>
struct SkStageCoef {
  double dsum;
  int uniform;
  double gre_u, gim_u;        // uniform drive: one coefficient for every qubit
  double gre[NH], gim[NH];    // general: per partner, gim carries the sign of the element's bit
};
template <int NH>
__device__ __forceinline__ void sk_stage_coef(SkStageCoef<NH>& o, const SkLane<NH>& ln, const SkCoef& c, int nq) {
  o.uniform = c.uniform;
  if (c.uniform) {
    o.gre_u = c.gre[0];
    o.gim_u = c.gim[0];
    o.dsum = fma(c.d[0], (double)ln.nzero, ln.dint);
  } else {
    double dsum = ln.dint;
    for (int p = 0; p < nq; ++p) dsum += ((ln.bits >> p) & 1) ? 0.0 : c.d[p];
    o.dsum = dsum;
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      const int p = 2 * j + ln.part;
      o.gre[j] = p < nq ? c.gre[p] : 0.0;
      o.gim[j] = p < nq ? ln.sg[j] * c.gim[p] : 0.0;
    }
  }
}
// value of the lane for -iH (negate for +iH); `other` = the element's other part of Y.
// pp: LL lines of the published vector (two consecutive lines re, im per element).
template <int NH>
__device__ __forceinline__ double sk_apply_lane(const SkStageCoef<NH>& sc, const SkLane<NH>& ln, double other,
                                                const SkPtrs<NH>& pp, unsigned seq, bool active, int* abort_flag) {
  double sre = 0.0, sim = 0.0;
  if (active) {
    uint4 lr[NH], li[NH];
    unsigned bad;
    SkPoll poll{abort_flag};
    do {
#pragma unroll
      for (int j = 0; j < NH; ++j) {
        lr[j] = ll_load(pp.p[j]);
        li[j] = ll_load(pp.p[j] + 16);
      }
      bad = 0;
#pragma unroll
      for (int j = 0; j < NH; ++j) bad |= ll_bad(lr[j], seq) | ll_bad(li[j], seq);
    } while (bad != 0 && !poll.give_up());
    // (H Y).re += g.re*pv.re - gim*pv.im ;  (H Y).im += g.re*pv.im + gim*pv.re   (gim signed by the own bit)
    if (sc.uniform) {
      double a0 = 0.0, c0 = 0.0, b0 = 0.0, d0 = 0.0;
#pragma unroll
      for (int j = 0; j < NH; ++j) {
        const double pre = ll_value(lr[j]), pim = ll_value(li[j]);
        a0 = fma(ln.wv[j], pre, a0); c0 = fma(ln.wv[j], pim, c0);
        b0 = fma(ln.sg[j], pim, b0); d0 = fma(ln.sg[j], pre, d0);
      }
      sre = fma(sc.gre_u, a0, -sc.gim_u * b0);
      sim = fma(sc.gre_u, c0, sc.gim_u * d0);
    } else {
      double sre0 = 0.0, sim0 = 0.0, sre1 = 0.0, sim1 = 0.0;
#pragma unroll
      for (int j = 0; j < NH; ++j) {
        const double pre = ll_value(lr[j]), pim = ll_value(li[j]);
        sre0 = fma(sc.gre[j], pre, sre0); sre1 = fma(-sc.gim[j], pim, sre1);
        sim0 = fma(sc.gre[j], pim, sim0); sim1 = fma(sc.gim[j], pre, sim1);
      }
      sre = sre0 + sre1; sim = sim0 + sim1;
    }
  }
  // part 0 needs (H Y).im, part 1 needs (H Y).re: hand the other lane the half it is missing
  const double give = ln.part == 0 ? sre : sim;
  const double recv = __shfl_xor_sync(0xffffffffu, give, 1);
  const double mine = (ln.part == 0 ? sim : sre) + recv;
  const double h = fma(sc.dsum, other, mine);
  return ln.part == 0 ? h : -h;
}

// ------------------------------------------------------------------------------------------
// forward: whole adaptive evolution
// ------------------------------------------------------------------------------------------
struct SkResume {
  double t, dt, error, cache_dt, cache_err;
  long long steps_in_interval, pos;
  int kk, in_interval, n_rec, status;   // status: 0 done, 1 log full (relaunch), 2 max_steps, 3 non-finite,
                                        //         4 replay short, 5 exchange poll timed out
  int n_acc, tape_ok;                   // accepted steps so far; 0 once the stage tape overflowed
};

struct SkFwd {
  SkProg prog;
  SkTab tab;
  int batch, nC;
  size_t dim, L;
  double atol, rtol, safety, minf, maxf;
  long long max_steps;
  int n_replay;
  const double* replay_dt;
  const unsigned char* replay_clipped;
  const double* tsave;
  int n_t;
  cplx* y_io;       // [L]   resume state
  cplx* k0_io;      // [L]
  cplx* states;     // [n_t][L]
  uint4* YS;        // [2][2L] LL lines: stage exchange (zeroed by the host before each launch)
  uint4* red;       // [2][nC][batch] LL lines: error-norm partial sums
  pd_step_record* log;
  int log_cap;
  SkResume* resume;
  // stage tape for the adjoint sweep (null = not recorded): per ACCEPTED step the six stage
  // inputs Y_1..Y_6 (Y_1 = y_n) and the six slopes k_1..k_6, [step][6][2L] doubles each
  double* tapeY;
  double* tapeK;
  int tape_cap;     // steps
  int* abort_flag;  // zeroed by the host; set by a lane whose poll timed out
  // independent parameter sets ("units", blockIdx.y): same register and masks, own coefficient
  // values, own controller, own buffers.  n_units > 1 requires nC == 1 (a unit never waits for
  // another CTA, so the launch needs no co-residency).
  int n_units;
  size_t det_stride, amp_stride;   // doubles per unit in prog.det_values / prog.amp_values
};

// NQG = 4 serves batches of tiny registers (one warp per unit): capped at 128 registers so that 16 units fit an SM
template <int NQG>
__global__ void __launch_bounds__(SK_T) __maxnreg__(NQG <= 4 ? 128 : 255) k_small_forward(const __grid_constant__ SkFwd P) {
  constexpr int NH = NQG / 2;
  __shared__ SkCoef coef[6];
  __shared__ double s_red[SK_T / 32][SK_MAXB];
  __shared__ double s_err[SK_MAXB];
  __shared__ double s_times[6];
  __shared__ double s_fac;
  extern __shared__ __align__(16) unsigned char s_tab[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = tid & 1;
  const unsigned cta = blockIdx.x;
  const int nq = P.prog.nq;
  const size_t dim = P.dim, L = P.L, L2 = 2 * P.L;
  SkProg prog = P.prog;
  const size_t unit = blockIdx.y;
  prog.det_values += unit * P.det_stride;
  prog.amp_values += unit * P.amp_stride;
  sk_cache_tables(prog, s_tab, tid);
  // this unit's buffers
  cplx* const y_io = P.y_io + unit * L;
  cplx* const k0_io = P.k0_io + unit * L;
  cplx* const states = P.states + unit * (size_t)P.n_t * L;
  uint4* const YS = P.YS + unit * 2 * L2;
  uint4* const red = P.red + unit * 2 * (size_t)P.nC * P.batch;
  pd_step_record* const log = P.log + unit * (size_t)P.log_cap;
  SkResume* const resume = P.resume + unit;
  double* const tapeY = P.tapeY ? P.tapeY + unit * (size_t)P.tape_cap * 6 * L2 : nullptr;
  double* const tapeK = P.tapeK ? P.tapeK + unit * (size_t)P.tape_cap * 6 * L2 : nullptr;

  const size_t r = (size_t)cta * blockDim.x + tid;
  const bool on = r < L2;
  const size_t rr = on ? r : (size_t)part;
  const size_t e = rr >> 1;
  const int col = (int)(e / dim);
  SkLane<NH> ln;
  sk_lane_init<NH>(ln, e, dim, nq, part, prog.diag[e & (dim - 1)]);
  double y, k[7], ynew = 0.0;
  y = on ? reinterpret_cast<const double*>(y_io)[rr] : 0.0;
  k[0] = on ? reinterpret_cast<const double*>(k0_io)[rr] : 0.0;

  SkResume R = *resume;
  if (R.status == 0 && R.kk >= P.n_t) {               // this unit finished in an earlier launch (uniform per CTA)
    if (cta == 0 && tid == 0) resume->n_rec = 0;
    return;
  }
  double t = R.t, dt = R.dt, error = R.error, cache_dt = R.cache_dt, cache_err = R.cache_err;
  long long steps = R.steps_in_interval, pos = R.pos;
  int n_rec = 0, status = 0, par = 0, rpar = 0;
  unsigned seq = 1, rseq = 1;
  int n_acc = R.n_acc;
  bool tape_ok = R.tape_ok != 0 && tapeY != nullptr;
  bool in_interval = R.in_interval != 0;
  const bool replay = P.n_replay > 0;
  int kk = R.kk;

  // max over batch columns of sqrt(mean(sq)) (Hairer norm), identical in every thread of the unit
  auto col_norm = [&](double sq) -> double {
    for (int b = 0; b < P.batch; ++b) {
      const double v = warp_sum_d(col == b ? sq : 0.0);
      if (lane == 0) s_red[warp][b] = v;
    }
    __syncthreads();
    if (tid < P.batch) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w][tid];
      ll_store(red + ((size_t)rpar * P.nC + cta) * P.batch + tid, tot, rseq);
    }
    // every CTA collects all partial sums: warp b (strided) polls the lines of column b, lane j
    // those of CTAs j, j+32, ... and the warp sums them in a fixed order
    for (int b = warp; b < P.batch; b += (int)(blockDim.x >> 5)) {
      double tot = 0.0;
      for (int c = lane; c < P.nC; c += 32) {
        const uint4* src = red + ((size_t)rpar * P.nC + c) * P.batch + b;
        uint4 v;
        SkPoll poll{P.abort_flag};
        do { v = ll_load(src); } while (ll_bad(v, rseq) != 0 && !poll.give_up());
        tot += ll_value(v);
      }
      tot = warp_sum_d(tot);
      if (lane == 0) s_err[b] = tot;
    }
    __syncthreads();
    double nrm = 0.0;
    for (int b = 0; b < P.batch; ++b) nrm = fmax(nrm, sqrt(s_err[b] / (double)dim));
    __syncthreads();          // s_red / s_err are reused by the next reduction
    ++rseq;
    rpar ^= 1;
    return nrm;
  };
  // one application of -iH(t_idx) on the lane value v (publishes it, polls the partners)
  // partner / own line pointers of the two exchange buffers
  SkPtrs<NH> pp[2];
  sk_ptrs_init<NH>(pp[0], ln, YS, rr);
  sk_ptrs_init<NH>(pp[1], ln, YS + L2, rr);
  auto apply_at = [&](int cidx, double v) -> double {
    const SkPtrs<NH>& q = par ? pp[1] : pp[0];
    if (on) ll_store(q.own, v, seq);
    const double vo = __shfl_xor_sync(0xffffffffu, v, 1);
    SkStageCoef<NH> sc;
    sk_stage_coef<NH>(sc, ln, coef[cidx], nq);
    double out = sk_apply_lane<NH>(sc, ln, vo, q, seq, on, P.abort_flag);
    par ^= 1; ++seq;
    return on ? out : 0.0;
  };

  if (kk < 0) {
    // ---- fresh start: k1 = f(t0, y0) and Hairer's initial step ----------
    t = P.tsave[0];
    if (tid < 6) s_times[tid] = t;
    __syncthreads();
    sk_eval_stages(prog, s_times, coef, tid);
    k[0] = apply_at(0, y);
    dt = 0.0;
    if (!replay) {
      const double yo = __shfl_xor_sync(0xffffffffu, y, 1);
      const double sc0 = P.atol + P.rtol * hypot(y, yo);
      const double a0 = y / sc0, a1 = k[0] / sc0;
      const double d0 = col_norm(on ? a0 * a0 : 0.0);
      const double d1 = col_norm(on ? a1 * a1 : 0.0);
      const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      __syncthreads();
      if (tid < 6) s_times[tid] = t + h0;
      __syncthreads();
      sk_eval_stages(prog, s_times, coef, tid);
      const double f1 = apply_at(0, fma(h0, k[0], y));
      const double a2 = (f1 - k[0]) / sc0;
      const double d2 = col_norm(on ? a2 * a2 : 0.0) / h0;
      const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 6.0);
      dt = fmin(100 * h0, h1);
    }
    error = 1.0; cache_dt = dt; cache_err = 1.0;
    kk = 0;
    // the start-up used one (replay) or two exchanges; the stage loop below indexes the buffers
    // statically, so hand it the buffer whose turn it is as pp[0]
    if (par) { const SkPtrs<NH> tmp = pp[0]; pp[0] = pp[1]; pp[1] = tmp; par = 0; }
    __syncthreads();
  }

  for (; kk < P.n_t && status == 0; ++kk) {
    const double t_next = P.tsave[kk];
    if (!in_interval) { cache_dt = dt; cache_err = error; steps = 0; }
    in_interval = true;
    while (t < t_next) {
      if (n_rec >= P.log_cap) { status = 1; break; }
      bool clipped;
      if (!replay) {
        // update_tstep; the pow is evaluated by one thread
        if (tid == 0) s_fac = error == 0.0 ? 0.0 : P.safety * pow(error, -0.2);
        __syncthreads();
        const double fac = s_fac;
        if (error == 0.0) dt = dt * P.maxf;
        else dt = error <= 1.0 ? dt * fmax(1.0, fmin(P.maxf, fac)) : dt * fmin(0.9, fmax(P.minf, fac));
        clipped = t + dt >= t_next;
      } else {
        if (pos >= P.n_replay) { status = 4; break; }
        dt = P.replay_dt[pos];
        clipped = P.replay_clipped[pos] != 0;
        ++pos;
      }
      if (clipped) { cache_dt = dt; cache_err = error; dt = t_next - t; }
      // coefficients of the six stage times
      if (tid < 6) s_times[tid] = t + dt * P.tab.alpha[tid];
      __syncthreads();
      sk_eval_stages(prog, s_times, coef, tid);
      if (tape_ok && n_acc >= P.tape_cap) tape_ok = false;
      double* tY = tape_ok ? tapeY + (size_t)n_acc * 6 * L2 : nullptr;
      double* tK = tape_ok ? tapeK + (size_t)n_acc * 6 * L2 : nullptr;
      if (tY && on) { tY[r] = y; tK[r] = k[0]; }
      // ---- stages 2..7 -------------------------------------------------------------------
#pragma unroll
      for (int i = 1; i < 7; ++i) {
        // six exchanges per attempt (and two in the start-up): the buffer parity of stage i is static
        const SkPtrs<NH>& q = pp[(i - 1) & 1];
        double v = y;
#pragma unroll
        for (int j = 0; j < i; ++j) v = fma(dt * P.tab.beta[i - 1][j], k[j], v);
        if (on) ll_store(q.own, v, seq);
        if (i == 6) ynew = v;
        if (tY && i < 6 && on) tY[(size_t)i * L2 + r] = v;
        const double vo = __shfl_xor_sync(0xffffffffu, v, 1);
        SkStageCoef<NH> sc;
        sk_stage_coef<NH>(sc, ln, coef[i - 1], nq);
        k[i] = sk_apply_lane<NH>(sc, ln, vo, q, seq, on, P.abort_flag);
        if (!on) k[i] = 0.0;
        if (tK && i < 6 && on) tK[(size_t)i * L2 + r] = k[i];
        ++seq;
      }
      // ---- error norm ----------------------------------------------------------------------
      double er = 0.0;
#pragma unroll
      for (int j = 0; j < 7; ++j) er = fma(dt * P.tab.eb[j], k[j], er);
      {
        const double yo = __shfl_xor_sync(0xffffffffu, y, 1);
        const double yno = __shfl_xor_sync(0xffffffffu, ynew, 1);
        const double m0 = fma(y, y, yo * yo), m1 = fma(ynew, ynew, yno * yno);
        const double sc = P.atol + P.rtol * sqrt(fmax(m0, m1));
        er /= sc;
      }
      error = col_norm(on ? er * er : 0.0);
      if (*(volatile int*)P.abort_flag) { status = 5; break; }
      if (!(error == error)) { status = 3; break; }
      const bool accepted = replay ? true : error <= 1.0;
      if (cta == 0 && tid == 0) {
        pd_step_record rec;
        rec.t = t; rec.dt = dt; rec.error = error; rec.accepted = accepted ? 1 : 0;
        rec.clipped = clipped ? 1 : 0; rec.interval = kk; rec._pad = 0;
        log[n_rec] = rec;
      }
      ++n_rec;
      if (accepted) {
        t = clipped ? t_next : t + dt;
        ++n_acc;
        y = ynew; k[0] = k[6];
      }
      if (++steps >= P.max_steps) { status = 2; break; }
    }
    if (status != 0) break;
    dt = cache_dt; error = cache_err;
    in_interval = false;
    if (on) reinterpret_cast<double*>(states + (size_t)kk * L)[r] = y;
  }
  if (on) {
    reinterpret_cast<double*>(y_io)[r] = y;
    reinterpret_cast<double*>(k0_io)[r] = k[0];
  }
  if (cta == 0 && tid == 0) {
    SkResume o;
    o.t = t; o.dt = dt; o.error = error; o.cache_dt = cache_dt; o.cache_err = cache_err;
    o.steps_in_interval = steps; o.pos = pos; o.kk = kk; o.in_interval = in_interval ? 1 : 0;
    o.n_rec = n_rec; o.status = status; o.n_acc = n_acc; o.tape_ok = tape_ok ? 1 : 0;
    *resume = o;
  }
}

// ------------------------------------------------------------------------------------------
// backward: whole discrete-adjoint sweep over the accepted-step sequence, reading the stage tape
// ------------------------------------------------------------------------------------------
struct SkStep {
  double t, dt;
  int interval, clipped;
};

struct SkBwd {
  SkProg prog;
  SkTab tab;
  int batch, nC;
  size_t dim, L;
  int n_t, n_steps;
  const SkStep* steps;
  const cplx* gstates;    // [n_t][L] or null
  const double* tapeY;    // [n_steps][6][2L]
  const double* tapeK;    // [n_steps][6][2L]
  uint4* KB;              // [2][2L] LL lines: exchange of the adjoint stage inputs (zeroed by the host)
  double* slotpart;       // [n_steps*6][nC][nred]  nred = n_det + 2 n_amp + 1
  double* wacc_elem;      // [L] or null
  cplx* lam_out;          // [L]
  int want_coef;
  int* abort_flag;
  // independent units (blockIdx.y), see SkFwd: per-unit step lists (stride max_steps), tapes
  // (stride tape_cap steps), cotangents, slot sums (stride max_steps*6*nC*nred) and outputs
  int n_units, max_steps, tape_cap;
  const int* unit_steps;      // [n_units] accepted steps of each unit (null: n_steps for all)
  size_t det_stride, amp_stride;
};

template <int NQG>
__global__ void __launch_bounds__(SK_T) __maxnreg__(NQG <= 4 ? 128 : 255) k_small_backward(const __grid_constant__ SkBwd P) {
  constexpr int NH = NQG / 2;
  __shared__ SkCoef coef[6];
  __shared__ double s_red[SK_T / 32][2 * SK_MAXTERMS + SK_MAXTERMS + 1];
  __shared__ unsigned long long s_dm[SK_MAXTERMS], s_am[SK_MAXTERMS];
  __shared__ double s_times[6];
  extern __shared__ __align__(16) unsigned char s_tab[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = tid & 1;
  const unsigned cta = blockIdx.x;
  const int nq = P.prog.nq, n_det = P.prog.n_det, n_amp = P.prog.n_amp;
  const int nred = n_det + 2 * n_amp + 1;
  const size_t L = P.L, dim = P.dim, L2 = 2 * P.L;
  if (tid < n_det) s_dm[tid] = P.prog.det_masks[tid];
  if (tid < n_amp) s_am[tid] = P.prog.amp_masks[tid];
  SkProg prog = P.prog;
  const size_t unit = blockIdx.y;
  prog.det_values += unit * P.det_stride;
  prog.amp_values += unit * P.amp_stride;
  sk_cache_tables(prog, s_tab, tid);
  __syncthreads();
  const int n_steps = P.unit_steps ? P.unit_steps[unit] : P.n_steps;
  const SkStep* const steps = P.steps + unit * (size_t)P.max_steps;
  const double* const tapeY = P.tapeY + unit * (size_t)P.tape_cap * 6 * L2;
  const double* const tapeK = P.tapeK + unit * (size_t)P.tape_cap * 6 * L2;
  uint4* const KB = P.KB + unit * 2 * L2;
  double* const slotpart = P.slotpart + unit * (size_t)P.max_steps * 6 * P.nC * nred;

  const size_t r = (size_t)cta * blockDim.x + tid;
  const bool on = r < L2;
  const size_t rr = on ? r : (size_t)part;
  const size_t e = rr >> 1;
  SkLane<NH> ln;
  sk_lane_init<NH>(ln, e, dim, nq, part, prog.diag[e & (dim - 1)]);
  const double* gst = P.gstates ? reinterpret_cast<const double*>(P.gstates + unit * (size_t)P.n_t * L) : nullptr;
  double wacc = 0.0;
  double lam = (on && gst) ? gst[(size_t)(P.n_t - 1) * L2 + rr] : 0.0;
  SkPtrs<NH> pp[2];
  sk_ptrs_init<NH>(pp[0], ln, KB, rr);
  sk_ptrs_init<NH>(pp[1], ln, KB + L2, rr);
  unsigned seq = 1;
  int hi = n_steps;
  for (int kk = P.n_t - 1; kk >= 1; --kk) {
    int lo = hi;
    while (lo > 0 && steps[lo - 1].interval == kk) --lo;
    for (int gi = hi - 1; gi >= lo; --gi) {
      const SkStep st = steps[gi];
      const double h = st.dt;
      __syncthreads();
      if (tid < 6) s_times[tid] = tid == 0 ? st.t : st.t + h * P.tab.alpha[tid - 1];
      __syncthreads();
      sk_eval_stages(prog, s_times, coef, tid);
      const char* tY = reinterpret_cast<const char*>(tapeY + (size_t)gi * 6 * L2);
      const double* tK = tapeK + (size_t)gi * 6 * L2;
      double yb[6];
#pragma unroll
      for (int i = 5; i >= 0; --i) {
        const SkPtrs<NH>& q = pp[(5 - i) & 1];      // six exchanges per step: static buffer parity
        double u = h * P.tab.b5[i] * lam;
#pragma unroll
        for (int j = i + 1; j < 6; ++j) u = fma(h * P.tab.beta[j - 1][i], yb[j], u);
        if (on) ll_store(q.own, u, seq);
        const double uo = __shfl_xor_sync(0xffffffffu, u, 1);
        SkStageCoef<NH> sc;
        sk_stage_coef<NH>(sc, ln, coef[i], nq);
        // tape values of this slot: element e of stage input i sits at byte offset e*16 (re, im)
        const char* ysrc = tY + (size_t)i * L2 * 8;
        const double kt = on ? tK[(size_t)i * L2 + r] : 0.0;
        const double2 own = *reinterpret_cast<const double2*>(ysrc + (ln.eoff >> 1));
        double2 pv[NH];
        if (P.want_coef) {
#pragma unroll
          for (int j = 0; j < NH; ++j)
            pv[j] = *reinterpret_cast<const double2*>(ysrc + ((ln.eoff ^ ln.xm[j]) >> 1));
        }
        yb[i] = -sk_apply_lane<NH>(sc, ln, uo, q, seq, on, P.abort_flag);
        if (!on) yb[i] = 0.0;
        ++seq;
        // ---- per-term gradient sums of this slot -----------------------------------------------
        // kb = conj(u_e); self = kb*Y_e; fl_j = kb*Y_partner(j).  Everything stays in registers.
        if (P.want_coef || P.wacc_elem) {
          const double ure = part == 0 ? u : uo, uim = part == 0 ? uo : u;
          const double self_im = on ? ure * own.y - uim * own.x : 0.0;
          if (part == 0) wacc += self_im;
          double hd = on ? u * kt : 0.0;
          if (P.want_coef) {
            double gd[NH], ga[NH], gb[NH];
#pragma unroll
            for (int j = 0; j < NH; ++j) {
              gd[j] = 0.0; ga[j] = 0.0; gb[j] = 0.0;
              if (ln.wv[j] != 0.0 && on) {
                const bool a = (ln.bits >> (2 * j + part)) & 1;
                const double fl_re = ure * pv[j].x + uim * pv[j].y;
                const double fl_im = ure * pv[j].y - uim * pv[j].x;
                gd[j] = a ? 0.0 : self_im;
                ga[j] = fl_im;
                gb[j] = a ? fl_re : -fl_re;
              }
            }
            for (int kd = 0; kd < n_det; ++kd) {
              const unsigned long long m = s_dm[kd];
              double v = 0.0;
#pragma unroll
              for (int j = 0; j < NH; ++j)
                if (ln.wv[j] != 0.0 && (m >> (nq - 1 - (2 * j + part)) & 1ull)) v += gd[j];
              v = warp_sum_d(v);
              if (lane == 0) s_red[warp][kd] = v;
            }
            for (int ka = 0; ka < n_amp; ++ka) {
              const unsigned long long m = s_am[ka];
              double va = 0.0, vb = 0.0;
#pragma unroll
              for (int j = 0; j < NH; ++j)
                if (ln.wv[j] != 0.0 && (m >> (nq - 1 - (2 * j + part)) & 1ull)) { va += ga[j]; vb += gb[j]; }
              va = warp_sum_d(va);
              vb = warp_sum_d(vb);
              if (lane == 0) { s_red[warp][n_det + 2 * ka] = va; s_red[warp][n_det + 2 * ka + 1] = vb; }
            }
          } else if (lane == 0) {
            for (int q = 0; q < nred - 1; ++q) s_red[warp][q] = 0.0;
          }
          hd = warp_sum_d(hd);
          if (lane == 0) s_red[warp][nred - 1] = hd;
          __syncthreads();
          for (int q = tid; q < nred; q += (int)blockDim.x) {   // nred can exceed a 32-thread CTA
            double tot = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w][q];
            slotpart[((size_t)gi * 6 + i) * P.nC * nred + (size_t)cta * nred + q] = tot;
          }
          __syncthreads();
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) lam += yb[i];
    }
    hi = lo;
    if (gst && on) lam += gst[(size_t)(kk - 1) * L2 + r];
  }
  if (on) {
    reinterpret_cast<double*>(P.lam_out + unit * L)[r] = lam;
    if (P.wacc_elem && part == 0) P.wacc_elem[unit * L + e] = wacc;
  }
}

// ------------------------------------------------------------------------------------------
// Lanczos recurrence of KRYLOV_SE for one column vector: iterations
// [j0, j1) of   r = H v_j;  alpha_j = <v_j, r>;  r -= alpha_j v_j + beta_{j-1} v_{j-1};
// beta_j = |r|;  v_{j+1} = r / beta_j   in one cooperative launch.  The recurrence does not depend
// on where the host stops it (exponential-error / breakdown tests on the tridiagonal matrix), so
// the host applies the reference's stopping rule a posteriori to the returned (alpha, beta) and
// uses the first m basis vectors.  H is frozen at t_eval (the interval end).
// ------------------------------------------------------------------------------------------
struct SkLanczos {
  SkProg prog;
  int nC;
  size_t dim;
  double t_eval;
  int j0, j1, max_m;   // iterations of this launch; basis[j0] (and basis[j0-1]) are already in place
  double beta_prev;    // beta_{j0-1} (0 for j0 = 0)
  cplx* basis;         // [max_m][dim]; j0 = 0: basis[0] receives v0 / |v0|
  const cplx* v0;      // j0 = 0 only
  double* alpha;       // [max_m] device
  double* beta;        // [max_m] device
  double* nrm_out;     // |v0| (j0 = 0)
  uint4* YS;           // [2][2 dim] LL lines (zeroed by the host)
  uint4* red;          // [2][nC] LL lines
  int* abort_flag;
};

template <int NQG>
__global__ void __launch_bounds__(SK_T) k_small_lanczos(const __grid_constant__ SkLanczos P) {
  constexpr int NH = NQG / 2;
  __shared__ SkCoef coef[6];
  __shared__ double s_red[SK_T / 32];
  __shared__ double s_tot;
  __shared__ double s_times[6];
  extern __shared__ __align__(16) unsigned char s_tab[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = tid & 1;
  const unsigned cta = blockIdx.x;
  const int nq = P.prog.nq;
  const size_t dim = P.dim, L2 = 2 * P.dim;
  SkProg prog = P.prog;
  sk_cache_tables(prog, s_tab, tid);
  const size_t r = (size_t)cta * blockDim.x + tid;
  const bool on = r < L2;
  const size_t rr = on ? r : (size_t)part;
  const size_t e = rr >> 1;
  SkLane<NH> ln;
  sk_lane_init<NH>(ln, e, dim, nq, part, prog.diag[e & (dim - 1)]);
  if (tid < 6) s_times[tid] = P.t_eval;
  __syncthreads();
  sk_eval_stages(prog, s_times, coef, tid);
  SkStageCoef<NH> sc;
  sk_stage_coef<NH>(sc, ln, coef[0], nq);
  int par = 0, rpar = 0;
  unsigned seq = 1, rseq = 1;
  SkPtrs<NH> pp[2];
  sk_ptrs_init<NH>(pp[0], ln, P.YS, rr);
  sk_ptrs_init<NH>(pp[1], ln, P.YS + L2, rr);

  // sum over all lanes of the unit, identical in every thread (fixed summation order)
  auto all_sum = [&](double x) -> double {
    const double v = warp_sum_d(x);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
      ll_store(P.red + (size_t)rpar * P.nC + cta, tot, rseq);
    }
    if (warp == 0) {
      double tot = 0.0;
      for (int c = lane; c < P.nC; c += 32) {
        uint4 v4;
        SkPoll poll{P.abort_flag};
        do { v4 = ll_load(P.red + (size_t)rpar * P.nC + c); } while (ll_bad(v4, rseq) != 0 && !poll.give_up());
        tot += ll_value(v4);
      }
      tot = warp_sum_d(tot);
      if (lane == 0) s_tot = tot;
    }
    __syncthreads();
    const double out = s_tot;
    __syncthreads();
    ++rseq; rpar ^= 1;
    return out;
  };

  double* basis = reinterpret_cast<double*>(P.basis);
  double vj, vprev = 0.0, beta_prev = P.beta_prev;
  if (P.j0 == 0) {
    const double x = on ? reinterpret_cast<const double*>(P.v0)[rr] : 0.0;
    const double nrm = sqrt(all_sum(x * x));
    if (cta == 0 && tid == 0) *P.nrm_out = nrm;
    vj = nrm > 0.0 ? x / nrm : 0.0;
    if (on) basis[rr] = vj;
  } else {
    vj = on ? basis[(size_t)P.j0 * L2 + rr] : 0.0;
    vprev = on ? basis[(size_t)(P.j0 - 1) * L2 + rr] : 0.0;
  }
  for (int j = P.j0; j < P.j1; ++j) {
    // r = H v_j: publish the lane, poll the partners; plain H (no factor -i):
    //   part 0 needs (H v).re, part 1 needs (H v).im; sk_apply_lane returns -iH v, i.e.
    //   part 0 -> (H v).im, part 1 -> -(H v).re, so swap the parts back through the sibling lane
    const SkPtrs<NH>& q = par ? pp[1] : pp[0];
    if (on) ll_store(q.own, vj, seq);
    const double vo = __shfl_xor_sync(0xffffffffu, vj, 1);
    const double mih = sk_apply_lane<NH>(sc, ln, vo, q, seq, on, P.abort_flag);   // lane's part of -i H v
    par ^= 1; ++seq;
    const double other = __shfl_xor_sync(0xffffffffu, mih, 1);
    // (-iHv).re = (Hv).im (held by part 0), (-iHv).im = -(Hv).re (held by part 1)
    double rl = part == 0 ? -other : other;      // part 0: (Hv).re = -(-iHv).im ; part 1: (Hv).im = (-iHv).re
    if (!on) rl = 0.0;
    const double a = all_sum(vj * rl);
    rl = rl - a * vj - beta_prev * vprev;
    const double b = sqrt(all_sum(rl * rl));
    if (cta == 0 && tid == 0) { P.alpha[j] = a; P.beta[j] = b; }
    const double vnext = b > 0.0 ? rl / b : 0.0;
    if (on && j + 1 < P.max_m) basis[(size_t)(j + 1) * L2 + rr] = vnext;
    vprev = vj; vj = vnext; beta_prev = b;
  }
}

static __global__ void k_fold_columns(const double* __restrict__ in, double* __restrict__ out, size_t dim, int batch) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= dim) return;
  double v = 0.0;
  for (int b = 0; b < batch; ++b) v += in[(size_t)b * dim + s];
  out[s] += v;
}

// out[slot][r] = sum over CTAs (fixed order) of the per-CTA partial sums of the adjoint sweep
static __global__ void k_sum_slots(const double* __restrict__ part, double* __restrict__ out, size_t n_slots, int nC,
                                   int nred) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots * nred) return;
  const size_t sl = i / nred;
  const int r = (int)(i % nred);
  double v = 0.0;
  for (int c = 0; c < nC; ++c) v += part[(sl * nC + c) * nred + r];
  out[i] = v;
}

// ---- batches of units: step lists and gradient scatter stay on the device --------------------------
// appends the accepted records of the last launch of every unit to its device step list
static __global__ void k_compact_log(const pd_step_record* __restrict__ log, int log_cap, const SkResume* __restrict__ resume,
                                     SkStep* __restrict__ steps, int max_steps, int* __restrict__ n_steps,
                                     int* __restrict__ n_attempts, int n_units) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_units) return;
  int n = n_steps[u];
  const int nr = resume[u].n_rec;
  for (int r = 0; r < nr; ++r) {
    const pd_step_record rec = log[(size_t)u * log_cap + r];
    if (rec.accepted && n < max_steps) {
      steps[(size_t)u * max_steps + n] = SkStep{rec.t, rec.dt, rec.interval, rec.clipped};
      ++n;
    }
  }
  n_steps[u] = n;
  n_attempts[u] += nr;
}
// scatters the per-slot term sums of every unit onto its sample-gradient arrays with the
// reference interpolation weights (engine.hpp::distribute_terms); one CTA per unit
static __global__ void k_distribute_units(const double* __restrict__ slot, const SkStep* __restrict__ steps,
                                          const int* __restrict__ n_steps, int max_steps, int nparts, int nred,
                                          SkTab tab, double dt, int ns, int n_det, int n_amp,
                                          double* __restrict__ g_det, double* __restrict__ g_amp) {
  extern __shared__ double acc[];      // [n_det][ns] then [n_amp][ns][2]
  const int u = blockIdx.x;
  const int n_acc = (n_det + 2 * n_amp) * ns;
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  const int nslots = n_steps[u] * 6;
  for (int sl = threadIdx.x; sl < nslots; sl += blockDim.x) {
    const SkStep st = steps[(size_t)u * max_steps + sl / 6];
    const int i = sl % 6;
    const double ts = st.t + st.dt * (i == 0 ? 0.0 : tab.alpha[i - 1]);
    const double fl = floor(ts / dt);
    long i1 = (long)fmin(fl, (double)(ns - 2));
    if (i1 < 0) i1 = 0;
    long i2 = i1 + 1 < (long)(ns - 2) ? i1 + 1 : (long)(ns - 2);
    if (i2 < 0) i2 = 0;
    const double x = (ts - i1 * dt) / dt;
    double sums[2 * SK_MAXTERMS + SK_MAXTERMS];
    for (int r = 0; r < nred - 1; ++r) {
      double v = 0.0;
      for (int c = 0; c < nparts; ++c) v += slot[(((size_t)u * max_steps * 6 + sl) * nparts + c) * nred + r];
      sums[r] = v;
    }
    for (int kd = 0; kd < n_det; ++kd) {
      const double g = 2.0 * sums[kd];
      atomicAdd(&acc[kd * ns + i1], g * (1.0 - x));
      atomicAdd(&acc[kd * ns + i2], g * x);
    }
    double* ga = acc + n_det * ns;
    for (int ka = 0; ka < n_amp; ++ka) {
      const double gre = sums[n_det + 2 * ka], gim = sums[n_det + 2 * ka + 1];
      atomicAdd(&ga[(ka * ns + i1) * 2], gre * (1.0 - x));
      atomicAdd(&ga[(ka * ns + i1) * 2 + 1], gim * (1.0 - x));
      atomicAdd(&ga[(ka * ns + i2) * 2], gre * x);
      atomicAdd(&ga[(ka * ns + i2) * 2 + 1], gim * x);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_det * ns; i += blockDim.x) g_det[(size_t)u * n_det * ns + i] = acc[i];
  for (int i = threadIdx.x; i < 2 * n_amp * ns; i += blockDim.x) g_amp[(size_t)u * 2 * n_amp * ns + i] = acc[n_det * ns + i];
}

inline void fill_tab(const Tableau& t, SkTab& o) {
  for (int i = 0; i < 6; ++i) o.alpha[i] = t.alpha[i];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) o.beta[i][j] = t.beta[i][j];
  for (int j = 0; j < 7; ++j) { o.b5[j] = t.b5[j]; o.eb[j] = t.b5[j] - t.b4[j]; }
}

// A unit that spans several CTAs polls lines written by its other CTAs: cooperative launch (all
// CTAs co-resident).  Units of one CTA never wait for another CTA: plain launch, any grid size.
// `lanes` = real lanes of one unit (2 x amplitudes x batch columns).  A single-CTA unit gets a CTA of just
// enough warps (a 2-qubit unit with four initial states is ONE warp: up to ~12 units per SM are then co-resident
// instead of 3, which is what bounds a batch of thousands of small parameter sets) and only the shared memory
// its tables need.
template <class K, class PT>
void launch_units(K kern, const PT& P, int nC, int n_units, size_t lanes, cudaStream_t s) {
  const size_t need = ((size_t)P.prog.n_det * P.prog.n_samples + (size_t)P.prog.n_amp * P.prog.n_samples * 2 +
                       P.prog.n_det + P.prog.n_amp) * 8;
  const unsigned dyn = need <= (size_t)SK_TABLE_BYTES ? (unsigned)((need + 15) & ~(size_t)15) : 0u;
  if (nC > 1) {
    if (n_units != 1) throw Error(PD_ERR_STATE, "small_ket: multi-unit launches need single-CTA units");
    void* args[1] = {const_cast<PT*>(&P)};
    PD_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)nC), dim3(SK_T), args, dyn, s));
  } else {
    const unsigned nthr = n_units > 1 ? (unsigned)std::min<size_t>(SK_T, std::max<size_t>(32, (lanes + 31) & ~(size_t)31))
                                      : (unsigned)SK_T;
    kern<<<dim3(1, (unsigned)n_units), nthr, dyn, s>>>(P);
    PD_CUDA_CHECK(cudaGetLastError());
  }
}

}  // namespace sk
}  // namespace pd
