// Tiled ket kernels: the HBM-bound path for large registers.
//
// H(t) = H_A(t) + H_B(t):
//   H_A = static diagonal + detuning diagonal + sigma-x flips on the LOW  bits [0, LA)
//   H_B =                                       sigma-x flips on the HIGH bits [LA, N)
// A type-A tile is 2^TB contiguous amplitudes (closes over the low bits); a type-B tile is
// 2^(N-LA) rows (stride 2^LA amplitudes) x 2^C contiguous columns (closes over the high bits).
// Either tile lives in shared memory once; every bit flip is a shared-memory partner read.
//
// One kernel does TWO things on its tile, so that no stage input is ever materialised:
//   finalise:  out_prev     = partial_prev + H_tau(t_prev) * Yprev,   Yprev = sum_j wprev_j v_j
//   start:     partial_next =                H_tau(t_next) * Ynext,   Ynext = sum_j wnext_j v_j
//                                                                            + wnext_out * out_prev
// Alternating tile types A,B,A,B,... a Dormand-Prince step is 7 launches (K0 start-only ...
// K6 finalise-only with the error norm in its epilogue) reading 34 and writing 13 vectors:
// 784 B per amplitude against the 576 B algorithmic figure.
#include "cuda_backend.cuh"

namespace pd {

namespace {

constexpr int TB = 12;            // tile = 4096 amplitudes = 64 KiB of shared memory
constexpr int TILE = 1 << TB;
#ifndef PD_TILED_LNT
#define PD_TILED_LNT 9
#endif
constexpr int LNT = PD_TILED_LNT;  // log2(threads per CTA)
constexpr int NT = 1 << LNT;      // threads per CTA; 2 CTAs per SM
constexpr int EPT = TILE / NT;    // 16 amplitudes per thread
constexpr int kMaxIn = 8;
constexpr int kMinTiledQubits = 16;
constexpr int kMaxTiledQubits = 2 * TB - 1;   // C >= 1 (32 B pieces); C >= 3 up to N = 21

struct BitCoef {   // coefficients of one application for the bits this tile type handles
  cplx kappa;                 // scale of the static diagonal (type A)
  cplx t00[kMaxQubits];       // per GLOBAL bit position p: T[a=0][a=0]  (diagonal, a = 0)
  cplx t11[kMaxQubits];
  cplx t01[kMaxQubits];       // row a=0 <- a'=1
  cplx t10[kMaxQubits];       // row a=1 <- a'=0
};

struct TiledParams {
  int nq, type, C, n_in, do_prev, do_next, do_err;
  size_t dim;
  const cplx* v[kMaxIn];
  double wprev[kMaxIn], wnext[kMaxIn], werr[kMaxIn];
  double wnext_out, werr_out;
  const cplx* partial_prev;
  cplx* out_prev;
  cplx* partial_next;
  cplx* ynext_out;
  const double* diag;
  double atol, rtol;
  double* err_partial;   // [gridDim.x] per-CTA sums of |err/scale|^2
};

__device__ __forceinline__ cplx ldg(const cplx* p) {
  double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return {v.x, v.y};
}

// global amplitude index of local element e of tile `tile`
__device__ __forceinline__ size_t gindex(int type, int C, size_t tile, int e) {
  if (type == 0) return (tile << TB) + (size_t)e;
  size_t row = (size_t)(e >> C), col = (size_t)(e & ((1 << C) - 1));
  return (row << TB) + (tile << C) + col;
}

// ---- TMEM as a software-managed accumulator store ------------------------------------------
// Each thread parks its 16 complex "z" accumulators (the v-part of the next stage input) in
// tensor memory between phases: 64 x 32-bit columns per thread, 128 columns per CTA (two warps
// share one 32-lane quarter), so two CTAs per SM use half of the 512 columns.  This frees 64
// registers per thread for deeper global-load pipelining; no tensor-core math is involved.
constexpr int kTmemCols = 128;
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem) {
  uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot_smem);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst),
               "r"(kTmemCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kTmemCols)
               : "memory");
}
__device__ __forceinline__ void tmem_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// 4 complex doubles (16 x b32) per call
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const cplx (&v)[4]) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r[4 * i + 0] = (uint32_t)__double2loint(v[i].re);
    r[4 * i + 1] = (uint32_t)__double2hiint(v[i].re);
    r[4 * i + 2] = (uint32_t)__double2loint(v[i].im);
    r[4 * i + 3] = (uint32_t)__double2hiint(v[i].im);
  }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, cplx (&v)[4]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i].re = __hiloint2double((int)r[4 * i + 1], (int)r[4 * i + 0]);
    v[i].im = __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]);
  }
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

constexpr int GP = 4;            // elements per TMEM transfer (16 x b32)
constexpr int NG = EPT / GP;     // 4 groups per thread

// Own part of H applied to the tile in shared memory for the GP elements e_i = t + NT*(GP*G + i).
// Everything about the element index that is known at compile time after unrolling (its bits
// 8..11 come from G and i) is resolved by the compiler: partner offsets and, for those bits, the
// coefficient choice.  Bits 0..7 come from the thread id: one select per bit and group.
//   UNI:  all qubits share one drive coefficient (global channel): out = c10*L + c01*(S-L) with
//         S = sum of all partners, L = sum of partners seen from a set bit -- 2-4 DADD per flip.
constexpr int NE = 2;   // elements per apply call (register pressure: S, L, out per element)
template <int TYPE, bool UNI, int G, int I0>
__device__ __forceinline__ void apply_group(const cplx* __restrict__ T, int t, const BitCoef& bc,
                                            int lb0, const cplx* dsum, cplx* out) {
  cplx S[NE], L[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    S[i] = {0.0, 0.0};
    L[i] = {0.0, 0.0};
    out[i] = {0.0, 0.0};
  }
#pragma unroll
  for (int lb = 0; lb < TB; ++lb) {
    if (TYPE == 1 && lb < lb0) continue;          // uniform: type-B tiles flip local bits [C, TB)
    const int gb = TYPE == 0 ? lb : TB + lb - lb0;
    if (lb < LNT) {
      const bool a = (t >> lb) & 1;
      const int tp = t ^ (1 << lb);
      cplx c{0.0, 0.0};
      if (!UNI) c = a ? bc.t10[gb] : bc.t01[gb];
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const cplx pv = T[tp + NT * (GP * G + I0 + i)];
        if (UNI) {
          S[i].re += pv.re; S[i].im += pv.im;
          L[i].re += a ? pv.re : 0.0; L[i].im += a ? pv.im : 0.0;
        } else {
          fma_acc(out[i], c, pv);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        constexpr int dummy = 0; (void)dummy;
        const int idx = GP * G + I0 + i;
        const bool a = (idx >> (lb - LNT)) & 1;               // compile-time after unrolling
        const cplx pv = T[t + NT * (idx ^ (1 << (lb - LNT)))];
        if (UNI) {
          S[i].re += pv.re; S[i].im += pv.im;
          if (a) { L[i].re += pv.re; L[i].im += pv.im; }
        } else {
          fma_acc(out[i], a ? bc.t10[gb] : bc.t01[gb], pv);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    if (UNI) {
      const cplx H{S[i].re - L[i].re, S[i].im - L[i].im};
      fma_acc(out[i], bc.t10[TYPE == 0 ? 0 : TB], L[i]);
      fma_acc(out[i], bc.t01[TYPE == 0 ? 0 : TB], H);
    }
    (void)dsum;
  }
}

// Compile-time recursion over the NG groups of a thread (group index must be a constant so that
// apply_group resolves element bits 8..11 at compile time).
template <int TYPE>
struct Pref {          // partial + static diagonal of one group, fetched one group ahead
  cplx pp[GP];
  double dg[GP];
};
template <int TYPE>
__device__ __forceinline__ void prefetch_group(const TiledParams& P, int t, size_t tile, size_t boff,
                                               int C, int G, Pref<TYPE>& pf) {
#pragma unroll
  for (int i = 0; i < GP; ++i) {
    const size_t gi = gindex(TYPE, C, tile, t + NT * (G * GP + i));
    pf.pp[i] = P.partial_prev ? ldg(P.partial_prev + boff + gi) : cplx{0.0, 0.0};
    pf.dg[i] = TYPE == 0 ? __ldg(P.diag + gi) : 0.0;
  }
}
template <int TYPE, bool UNI, int G>
struct Phase1b {
  static __device__ __forceinline__ void run(const TiledParams& P, const BitCoef& cprev,
                                             const cplx* __restrict__ T, int t, size_t tile,
                                             size_t boff, int C, int lb0, cplx hi_prev,
                                             const cplx (*tab)[64], uint32_t tbase, double& err_acc,
                                             const Pref<TYPE>& cur) {
    Pref<TYPE> nxt;
    if (G + 1 < NG) prefetch_group<TYPE>(P, t, tile, boff, C, G + 1, nxt);
    cplx dsum[GP], o[GP];
#pragma unroll
    for (int i = 0; i < GP; ++i) dsum[i] = {0.0, 0.0};
    apply_group<TYPE, UNI, G, 0>(T, t, cprev, lb0, dsum, o);
    apply_group<TYPE, UNI, G, 2>(T, t, cprev, lb0, dsum + 2, o + 2);
    cplx zq[GP];
    tmem_ld4(tbase + (uint32_t)(G * 16), zq);
#pragma unroll
    for (int i = 0; i < GP; ++i) {
      const int e = t + NT * (G * GP + i);
      const size_t gi = boff + gindex(TYPE, C, tile, e);
      if (TYPE == 0) {
        const cplx ds = cplx{cprev.kappa.re * cur.dg[i], cprev.kappa.im * cur.dg[i]} + hi_prev +
                        tab[0][e & 63] + tab[1][(e >> 6) & 63];
        fma_acc(o[i], ds, T[e]);
      }
      o[i] = o[i] + cur.pp[i];
      P.out_prev[gi] = o[i];
      zq[i].re = fma(P.wnext_out, o[i].re, zq[i].re);
      zq[i].im = fma(P.wnext_out, o[i].im, zq[i].im);
      if (P.do_err) {
        // here z accumulates sum_j werr_j v_j (wnext := werr, wnext_out := werr_out)
        const cplx y1 = T[e];
        const cplx y0 = ldg(P.v[0] + gi);
        const double sc = P.atol + P.rtol * fmax(hypot(y0.re, y0.im), hypot(y1.re, y1.im));
        const double er = zq[i].re / sc, ei = zq[i].im / sc;
        err_acc += er * er + ei * ei;
      }
    }
    if (P.do_next) tmem_st4(tbase + (uint32_t)(G * 16), zq);
    Phase1b<TYPE, UNI, G + 1>::run(P, cprev, T, t, tile, boff, C, lb0, hi_prev, tab, tbase, err_acc, nxt);
  }
};
template <int TYPE, bool UNI>
struct Phase1b<TYPE, UNI, NG> {
  static __device__ __forceinline__ void run(const TiledParams&, const BitCoef&, const cplx*, int,
                                             size_t, size_t, int, int, cplx, const cplx (*)[64],
                                             uint32_t, double&, const Pref<TYPE>&) {}
};
template <int TYPE, bool UNI, int G>
struct Phase2b {
  static __device__ __forceinline__ void run(const TiledParams& P, const BitCoef& cnext,
                                             const cplx* __restrict__ T, int t, size_t tile,
                                             size_t boff, int C, int lb0, cplx hi_next,
                                             const cplx (*tab)[64]) {
    cplx dsum[GP], o[GP];
    double dg[GP];
#pragma unroll
    for (int i = 0; i < GP; ++i) {
      dsum[i] = {0.0, 0.0};
      dg[i] = TYPE == 0 ? __ldg(P.diag + gindex(TYPE, C, tile, t + NT * (G * GP + i))) : 0.0;
    }
    apply_group<TYPE, UNI, G, 0>(T, t, cnext, lb0, dsum, o);
    apply_group<TYPE, UNI, G, 2>(T, t, cnext, lb0, dsum + 2, o + 2);
#pragma unroll
    for (int i = 0; i < GP; ++i) {
      const int e = t + NT * (G * GP + i);
      if (TYPE == 0) {
        const cplx ds = cplx{cnext.kappa.re * dg[i], cnext.kappa.im * dg[i]} + hi_next +
                        tab[0][e & 63] + tab[1][(e >> 6) & 63];
        fma_acc(o[i], ds, T[e]);
      }
      P.partial_next[boff + gindex(TYPE, C, tile, e)] = o[i];
    }
    Phase2b<TYPE, UNI, G + 1>::run(P, cnext, T, t, tile, boff, C, lb0, hi_next, tab);
  }
};
template <int TYPE, bool UNI>
struct Phase2b<TYPE, UNI, NG> {
  static __device__ __forceinline__ void run(const TiledParams&, const BitCoef&, const cplx*, int,
                                             size_t, size_t, int, int, cplx, const cplx (*)[64]) {}
};

#ifndef PD_TILED_MINB
#define PD_TILED_MINB 2
#endif
template <int TYPE, bool UNI>
__global__ void __launch_bounds__(NT, PD_TILED_MINB)
k_tiled(const __grid_constant__ TiledParams P, const __grid_constant__ BitCoef cprev,
        const __grid_constant__ BitCoef cnext) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* T = reinterpret_cast<cplx*>(smem_raw);
  __shared__ cplx tab_prev[2][64], tab_next[2][64];   // detuning diagonal over local bits 0-5 / 6-11
  __shared__ double red[NT / 32];
  __shared__ uint32_t tmem_slot;

  const int t = threadIdx.x;
  const int warp = t >> 5;
  const int nq = P.nq, C = P.C;
  const size_t tiles_per_vec = P.dim >> TB;
  const size_t tile = blockIdx.x % tiles_per_vec;
  const size_t boff = (blockIdx.x / tiles_per_vec) * P.dim;     // batch column offset
  // bits this tile closes over: type A local [0,TB) = global [0,TB); type B local [C,TB) = global [TB,nq)
  const int lb0 = TYPE == 0 ? 0 : C;

  if (warp == 0) tmem_alloc(&tmem_slot);

  // detuning-diagonal tables (type A): sum over local bits of (a ? t11 : t00), split 6 + 6 bits;
  // the bits above the tile are constant per tile.
  cplx hi_prev{0, 0}, hi_next{0, 0};
  if (TYPE == 0) {
    if (t < 128) {
      int half = t >> 6, x = t & 63;
      cplx sp{0, 0}, sn{0, 0};
      for (int b = 0; b < 6; ++b) {
        int gb = half * 6 + b;
        if (gb < nq) {
          bool a = (x >> b) & 1;
          sp = sp + (a ? cprev.t11[gb] : cprev.t00[gb]);
          sn = sn + (a ? cnext.t11[gb] : cnext.t00[gb]);
        }
      }
      tab_prev[half][x] = sp;
      tab_next[half][x] = sn;
    }
    for (int gb = TB; gb < nq; ++gb) {
      bool a = (tile >> (gb - TB)) & 1;
      hi_prev = hi_prev + (a ? cprev.t11[gb] : cprev.t00[gb]);
      hi_next = hi_next + (a ? cnext.t11[gb] : cnext.t00[gb]);
    }
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  // this thread's TMEM window: lane quarter of its warp, 64 columns
  const uint32_t tbase = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * (EPT * 4));

  // ---- phase 1a: one pass over the input vectors builds Yprev (-> smem) and the v-part of
  //      Ynext (-> TMEM).  8 elements per sub-pass, 8 x 16 B loads in flight per thread and input.
  constexpr int QP = 4;   // elements per sub-pass; four input vectors per iteration: 16 loads in flight
  constexpr int JU = 4;
#pragma unroll
  for (int q0 = 0; q0 < EPT; q0 += QP) {
    cplx yp[QP], zq[QP];
    size_t gi[QP];
#pragma unroll
    for (int i = 0; i < QP; ++i) {
      yp[i] = {0.0, 0.0};
      zq[i] = {0.0, 0.0};
      gi[i] = boff + gindex(TYPE, C, tile, t + NT * (q0 + i));
    }
    for (int j = 0; j < P.n_in; j += JU) {
      cplx x[JU][QP];
      double wp[JU], wn[JU];
#pragma unroll
      for (int u = 0; u < JU; ++u) {
        const bool on = j + u < P.n_in;                 // uniform
        const cplx* vu = P.v[on ? j + u : j];
        wp[u] = on ? P.wprev[j + u] : 0.0;
        wn[u] = on ? P.wnext[j + u] : 0.0;
#pragma unroll
        for (int i = 0; i < QP; ++i) x[u][i] = ldg(vu + gi[i]);
      }
#pragma unroll
      for (int u = 0; u < JU; ++u)
#pragma unroll
        for (int i = 0; i < QP; ++i) {
          yp[i].re = fma(wp[u], x[u][i].re, yp[i].re); yp[i].im = fma(wp[u], x[u][i].im, yp[i].im);
          zq[i].re = fma(wn[u], x[u][i].re, zq[i].re); zq[i].im = fma(wn[u], x[u][i].im, zq[i].im);
        }
    }
    if (P.do_prev) {
#pragma unroll
      for (int i = 0; i < QP; ++i) T[t + NT * (q0 + i)] = yp[i];
    }
    tmem_st4(tbase + (uint32_t)((q0 / GP) * 16), zq);
  }
  Pref<TYPE> pf0;
  if (P.do_prev) prefetch_group<TYPE>(P, t, tile, boff, C, 0, pf0);
  tmem_wait_st();
  __syncthreads();

  // ---- phase 1b: finalise the previous application
  double err_acc = 0.0;
  if (P.do_prev) {
    Phase1b<TYPE, UNI, 0>::run(P, cprev, T, t, tile, boff, C, lb0, hi_prev, tab_prev, tbase, err_acc, pf0);
    tmem_wait_st();
  }
  if (P.do_err) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err_acc += __shfl_xor_sync(0xffffffffu, err_acc, o);
    if ((t & 31) == 0) red[warp] = err_acc;
    __syncthreads();
    if (t == 0) {
      double s = 0.0;
      for (int w = 0; w < NT / 32; ++w) s += red[w];
      P.err_partial[blockIdx.x] = s;
    }
  }
  if (P.do_next) {
    // ---- phase 2a: Ynext -> smem (and optionally to global: y_{n+1})
    __syncthreads();
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      cplx zq[GP];
      tmem_ld4(tbase + (uint32_t)(g * 16), zq);
#pragma unroll
      for (int i = 0; i < GP; ++i) {
        const int e = t + NT * (g * GP + i);
        T[e] = zq[i];
        if (P.ynext_out) P.ynext_out[boff + gindex(TYPE, C, tile, e)] = zq[i];
      }
    }
    __syncthreads();
    // ---- phase 2b: start the next application
    Phase2b<TYPE, UNI, 0>::run(P, cnext, T, t, tile, boff, C, lb0, hi_next, tab_next);
  }
  // ---- release tensor memory (same warp that allocated it)
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot);
}

void fill_coef(const SiteOps& so, int nq, BitCoef& bc) {
  bc.kappa = so.kappa;
  for (int q = 0; q < nq; ++q) {
    int p = nq - 1 - q;   // global bit position of qubit q
    bc.t00[p] = so.T[q * 4 + 0];
    bc.t01[p] = so.T[q * 4 + 1];
    bc.t10[p] = so.T[q * 4 + 2];
    bc.t11[p] = so.T[q * 4 + 3];
  }
}

bool g_attr_set[64] = {};     // per device: function attributes belong to the device's context
bool uniform_drive(const BitCoef& a, int nq) {
  for (int p = 1; p < nq; ++p)
    if (a.t01[p].re != a.t01[0].re || a.t01[p].im != a.t01[0].im || a.t10[p].re != a.t10[0].re ||
        a.t10[p].im != a.t10[0].im)
      return false;
  return true;
}
void launch(const TiledParams& P, const BitCoef& cp, const BitCoef& cn, int batch, cudaStream_t s) {
  int dev = 0;
  PD_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) throw Error(PD_ERR_STATE, "device index out of range");
  if (!g_attr_set[dev]) {
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_tiled<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 16));
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_tiled<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 16));
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_tiled<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 16));
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_tiled<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * 16));
    g_attr_set[dev] = true;
  }
  unsigned grid = (unsigned)((P.dim >> TB) * (size_t)batch);
  const bool uni = uniform_drive(cp, P.nq) && uniform_drive(cn, P.nq);
  if (P.type == 0) {
    if (uni) k_tiled<0, true><<<grid, NT, TILE * 16, s>>>(P, cp, cn);
    else k_tiled<0, false><<<grid, NT, TILE * 16, s>>>(P, cp, cn);
  } else {
    if (uni) k_tiled<1, true><<<grid, NT, TILE * 16, s>>>(P, cp, cn);
    else k_tiled<1, false><<<grid, NT, TILE * 16, s>>>(P, cp, cn);
  }
  PD_CUDA_CHECK(cudaGetLastError());
}

__global__ void k_sum_partials(const double* __restrict__ partial, int per_col, double* out) {
  __shared__ double sh[32];
  int b = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < per_col; i += blockDim.x) s += partial[(size_t)b * per_col + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sh[w];
    out[b] = tot;
  }
}

TiledParams base_params(const Geometry& g) {
  TiledParams P{};
  P.nq = g.nq;
  P.dim = g.dim;
  P.C = 2 * TB - g.nq;   // type-B tile: 2^(nq-TB) rows x 2^C columns
  P.diag = g.diag;
  return P;
}

}  // namespace

bool tiled_ket_supported(const Geometry& g) {
  return g.kind == PD_KET && g.nq >= kMinTiledQubits && g.nq <= kMaxTiledQubits;
}

// out = G (sum_j w_j in_j); comb (nullable) = the combined input.  Two launches: type A starts
// (diagonal + low bits), type B finalises (high bits).  `tmp` holds the partial in between.
int launch_tiled_stage_ket(const Geometry& g, cplx* out, cplx* comb, int n_in,
                           const cplx* const* ins, const double* w, const SiteOps& so, cplx* tmp,
                           cudaStream_t s) {
  if (n_in > kMaxIn) throw Error(PD_ERR_INVALID, "tiled stage takes at most 8 inputs");
  BitCoef bc;
  fill_coef(so, g.nq, bc);
  TiledParams A = base_params(g);
  A.type = 0; A.n_in = n_in; A.do_next = 1;
  for (int j = 0; j < n_in; ++j) { A.v[j] = ins[j]; A.wnext[j] = w[j]; }
  A.partial_next = tmp;
  A.ynext_out = comb;
  launch(A, bc, bc, g.batch, s);
  TiledParams B = base_params(g);
  B.type = 1; B.n_in = n_in; B.do_prev = 1;
  for (int j = 0; j < n_in; ++j) { B.v[j] = ins[j]; B.wprev[j] = w[j]; }
  B.partial_prev = tmp;
  B.out_prev = out;
  launch(B, bc, bc, g.batch, s);
  return 2;
}

// One Dormand-Prince step as 7 alternating launches.  k[0] holds f(t, y) on entry (FSAL);
// on exit k[1..6] are filled, ynew = y_{n+1}, err_out[b] = sum |err/scale|^2 per batch column.
int launch_tiled_dp5_step(const Geometry& g, const cplx* y, cplx* const* k, cplx* ynew,
                          const SiteOps* stage_ops /* [7], index i = stage i+1 */, const double* beta,
                          const double* b5, const double* ew, double dt, double atol, double rtol,
                          cplx* tmp_a, cplx* tmp_b, double* err_partial, double* err_out,
                          cudaStream_t s) {
  BitCoef bc[7];
  for (int i = 1; i < 7; ++i) fill_coef(stage_ops[i], g.nq, bc[i]);
  // launch m (0..6): tile type m%2; finalises stage m+1 (if m >= 1), starts stage m+2 (if m <= 5)
  cplx* partial[2] = {tmp_a, tmp_b};
  for (int m = 0; m < 7; ++m) {
    TiledParams P = base_params(g);
    P.type = m % 2;
    int fin = m;        // k index finalised: k[fin] (stage fin+1), needs fin >= 1
    int sta = m + 1;    // k index started:   k[sta] (stage sta+1), needs sta <= 6
    P.do_prev = m >= 1;
    P.do_next = m <= 5;
    // inputs: y, k[0..m-1]
    P.n_in = 1 + m;
    if (P.n_in > kMaxIn) throw Error(PD_ERR_STATE, "dp5 tiled: too many inputs");
    P.v[0] = y;
    for (int j = 0; j < m; ++j) P.v[1 + j] = k[j];
    if (P.do_prev) {
      // Yprev = input of stage fin+1 = y + dt * sum_{j<fin} beta[fin-1][j] k[j]
      P.wprev[0] = 1.0;
      for (int j = 0; j < fin; ++j) P.wprev[1 + j] = dt * beta[(fin - 1) * 6 + j];
      P.partial_prev = partial[(m + 1) % 2];
      P.out_prev = k[fin];
    }
    if (P.do_next) {
      // Ynext = input of stage sta+1 = y + dt * sum_{j<sta} beta[sta-1][j] k[j]; k[sta-1] = out_prev
      P.wnext[0] = 1.0;
      for (int j = 0; j < m; ++j) P.wnext[1 + j] = dt * beta[(sta - 1) * 6 + j];
      P.wnext_out = m >= 1 ? dt * beta[(sta - 1) * 6 + (sta - 1)] : 0.0;
      if (m == 0) { /* stage 2 input: y + dt*beta[0][0]*k[0]; k[0] is not an input yet */ }
      P.partial_next = partial[m % 2];
      if (sta == 6) P.ynext_out = ynew;
    }
    if (m == 0) {
      // k[0] (FSAL) must be an explicit input of the first launch
      P.n_in = 2;
      P.v[1] = k[0];
      P.wnext[1] = dt * beta[0];
    } else {
      // for m >= 1 the inputs are y, k[0..m-1]; out_prev = k[m] enters Ynext through wnext_out
      P.wnext_out = dt * beta[(sta - 1) * 6 + m];
      if (!P.do_next) P.wnext_out = 0.0;
    }
    if (m == 6) {
      // error norm in the epilogue: err = sum_j ew_j k_j (k[6] = out_prev), y0 = v[0], y1 = Yprev
      P.do_err = 1;
      P.werr_out = ew[6];
      for (int j = 0; j < 8; ++j) P.wnext[j] = 0.0;
      for (int j = 0; j < 6; ++j) P.wnext[1 + j] = ew[j];
      P.wnext_out = ew[6];
      P.atol = atol; P.rtol = rtol;
      P.err_partial = err_partial;
    }
    const BitCoef& cp = bc[P.do_prev ? fin : 1];
    const BitCoef& cn = bc[P.do_next ? sta : 6];
    launch(P, cp, cn, g.batch, s);
  }
  int per_col = (int)(g.dim >> TB);
  k_sum_partials<<<g.batch, 256, 0, s>>>(err_partial, per_col, err_out);
  (void)b5;
  return 8;
}

size_t tiled_err_partial_count(const Geometry& g) { return (g.dim >> TB) * (size_t)g.batch; }

}  // namespace pd
