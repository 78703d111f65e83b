// Streaming tile kernels for large kets ("stream" family): one bit-group of H per launch.
//
// H(t) = H_A(t) + sum_g H_g(t):
//   H_A = static diagonal + detuning diagonal + sigma-x flips on the LOW 12 bits (contiguous tile)
//   H_g = sigma-x flips on a group of <= 8 higher bits (strided tile: 2^nb rows x 2^(12-nb) columns,
//         so every global access is a contiguous piece of >= 256 B -- the size below which strided
//         pieces lose bandwidth on B200, profiles/r01_membench.md)
//
// One launch streams one 4096-amplitude tile per CTA through 64 KiB of shared memory:
//   A launch :  Y = sum_j w_j v_j  (stage combination on the fly, written once as Ymat)
//               out = H_A Y
//   g launch :  out += H_g Ymat      (in place)
// A Dormand-Prince stage is 1 + G launches (G = ceil((N-12)/8) groups; G = 2 at N = 26).  Each
// launch is a pure stream (a few coalesced vectors in, one or two out) with a single shared-memory
// pass, small enough in registers and shared memory for three CTAs per SM, which is what keeps HBM
// busy while other CTAs are in their shared-memory phase.
//
// L2 blocking (round 2).  B200's L2 serves hits at 15-19 TB/s against 6.5 TB/s from HBM
// (profiles/r02_l2bench.md), so the A launch and the FIRST group launch of a stage run as ONE dataflow
// launch, k_stream_ag: work items are handed out in ticket order, chunk by chunk (256 tiles = 16 MiB per
// vector; 64 tiles = 4 MiB by default): A tiles of chunk c alternate, ticket by ticket, with the group tiles of
// chunk c - lag, which wait on a per-chunk completion counter and find Ymat and the partial result in L2 (the A
// tiles write them with an evict_last policy, inputs and final results stream with evict_first).  Per stage the HBM traffic falls from
// n + 8.5 to n + 5.5 vector passes (n = vectors combined); a DP5 step at N = 26 moves 62 passes instead
// of 80.
#include <cuda.h>

#include <cstdlib>
#include <map>
#include <mutex>

#include "cuda_backend.cuh"

namespace pd {

namespace {

// Tile / register complex type: the storage precision of the build (complex128, or {float, float} in the
// complex64 build, where the whole tile arithmetic runs in float and only reductions stay double).
#if defined(PD_C64)
using treal = float;
struct alignas(8) tcplx { float re, im; };
__host__ __device__ inline tcplx operator*(tcplx a, tcplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__host__ __device__ inline void fma_acc(tcplx& acc, tcplx a, tcplx b) {
  acc.re = fmaf(a.re, b.re, fmaf(-a.im, b.im, acc.re));
  acc.im = fmaf(a.re, b.im, fmaf(a.im, b.re, acc.im));
}
#else
using treal = double;
using tcplx = cplx;
#endif
static_assert(sizeof(tcplx) == sizeof(amp_t), "tile type and storage type share one layout");
__host__ __device__ inline const tcplx* tc(const amp_t* p) { return reinterpret_cast<const tcplx*>(p); }
__host__ __device__ inline tcplx* tc(amp_t* p) { return reinterpret_cast<tcplx*>(p); }
constexpr int kAmpBytes = (int)sizeof(tcplx);
#if defined(PD_MINCTAS)
constexpr int kMinCtas = PD_MINCTAS;     // (A/B builds)
#else
constexpr int kMinCtas = kC64 ? 4 : 3;   // resident CTAs per SM the REAL-drive kernels are compiled for
#endif

constexpr int TB = 12;
constexpr int TILE = 1 << TB;
constexpr int NT = 256;
constexpr int EPT = TILE / NT;   // 16
constexpr int kMaxIn = 8;
constexpr int kMaxGroupBits = 8;

// Every generator the engine applies is kappa * H with H Hermitian (kappa = 1, -i or +i; pd_common.hpp
// site_ops_ket), so the tiles accumulate h = H Y with a REAL diagonal and conjugate-paired flips and scale
// once: per flip 2 FMAs when the drive has no phase (REAL), 4 otherwise.
struct StreamCoef {
  tcplx kappa;
  treal d[kMaxQubits];        // per GLOBAL bit position p: diagonal entry of H for bit value 0 (|r>)
  treal gre[kMaxQubits];      // flip entry of H, row bit 1 <- partner bit 0: g = gre + i gim;
  treal gim[kMaxQubits];      //                  row bit 0 <- partner bit 1: conj(g)
};

struct StreamParams {
  int nq;
  int lo, nb, C;              // strided groups: row bits [lo, lo+nb), C = TB - nb column bits
  int n_in;
  size_t dim;
  const tcplx* v[kMaxIn];
  treal w[kMaxIn];
  tcplx* ymat;                 // A launch: combined input written here (nullable)
  tcplx* out;
  const double* diag;         // Dint split for tiles (Geometry::diag_parts): no 8 B/amplitude stream from HBM
  // embedded error estimate of the Dormand-Prince step (last stage only):
  //   A launch:      aux = sum_j w2_j v_j                  (partial error vector, w2 = dt*(b5-b4))
  //   last g launch: err = aux + werr*out;  sum |err / (atol + rtol*max(|y0|,|Ymat|))|^2 per CTA
  treal w2[kMaxIn];
  tcplx* aux;                  // A: written (nullable);  g: read when err_partial != null
  const tcplx* y0;
  double werr, atol, rtol;
  double* err_partial;        // [n_tiles] (g launch, nullable)
  int pol_st, pol_ld;         // L2 policy kinds (l2_policy) of this launch's result stores / tile loads
  int pol_in;                 // A launch: L2 policy of the input loads (a plain application re-reads its input
                              // in the group tiles, a combination is read once)
  int via_ring;               // pipelined group items: tile data arrives as TMA boxes through the ring (else: loads)
};

#if defined(PD_C64)
__device__ __forceinline__ tcplx ldcs(const tcplx* p) {
  float2 v = __ldcs(reinterpret_cast<const float2*>(p));
  return {v.x, v.y};
}
#else
__device__ __forceinline__ tcplx ldcs(const tcplx* p) {
  double2 v = __ldcs(reinterpret_cast<const double2*>(p));
  return {v.x, v.y};
}
#endif

// L2 residency control of the dataflow launch (k_stream_ag): what an A tile writes (Ymat, partial result) is
// re-read by a group tile a few chunks later and should outlive the input vectors streaming through L2; what
// a group tile writes is not touched again before the next launch.
__device__ __forceinline__ unsigned long long l2_policy(int kind) {   // 0 normal, 1 evict_last, 2 evict_first
  unsigned long long pol;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#if defined(PD_C64)
__device__ __forceinline__ void st_pol(tcplx* p, tcplx v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.re), "f"(v.im), "l"(pol) : "memory");
}
__device__ __forceinline__ tcplx ld_pol(const tcplx* p, unsigned long long pol) {
  tcplx v;
  asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.re), "=f"(v.im) : "l"(p), "l"(pol));
  return v;
}
// one amplitude, global -> shared, asynchronously (LDGSTS): 8-byte copies exist only with .ca
__device__ __forceinline__ void cp_async_amp(tcplx* smem_dst, const tcplx* src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_amp_pol(tcplx* smem_dst, const tcplx* src, unsigned long long pol) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
#else
__device__ __forceinline__ void st_pol(tcplx* p, tcplx v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.re), "d"(v.im), "l"(pol) : "memory");
}
__device__ __forceinline__ tcplx ld_pol(const tcplx* p, unsigned long long pol) {
  tcplx v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.re), "=d"(v.im) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void cp_async_amp(tcplx* smem_dst, const tcplx* src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_amp_pol(tcplx* smem_dst, const tcplx* src, unsigned long long pol) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
#endif

// CTA barrier of a tile routine: every thread of the CTA (NBAR = 0), or the NBAR consumer threads of the
// persistent kernel (named barrier 1; its producer warp does not take part).
template <int NBAR>
__device__ __forceinline__ void tile_sync() {
  if (NBAR == 0) __syncthreads();
  else asm volatile("bar.sync 1, %0;" ::"n"(NBAR) : "memory");
}

// h += (gre + i*s) * pv  (REAL: s == 0)
template <bool REAL>
__device__ __forceinline__ void flip_acc(tcplx& h, treal gre, treal s, tcplx pv) {
  h.re = fma(gre, pv.re, h.re);
  h.im = fma(gre, pv.im, h.im);
  if (!REAL) {
    h.re = fma(-s, pv.im, h.re);
    h.im = fma(s, pv.re, h.im);
  }
}

// ---- A tile: contiguous 4096 amplitudes; combination + diagonal + flips of the low 12 bits -----------
// Thread t owns elements e = t + 256 i (i < 16): tile bits 0-7 come from t, bits 8-11 from i.  The partner
// of a flip on bit lb < 8 lives at a per-thread base pointer + a compile-time offset; on bit lb >= 8 it is
// one of the thread's own elements; the sign of the flip's imaginary part (bit value of e) is per thread
// (lb < 8) or a compile-time constant (lb >= 8): no per-element index arithmetic or selects.
// second phase of an A tile: T holds the combined input Y of the tile (written by this CTA, not yet synced)
template <bool REAL, int NBAR>
__device__ __forceinline__ void a_tile_flips(const StreamParams& P, const StreamCoef& cf, tcplx* T, size_t tile,
                                             size_t base) {
  const int t = threadIdx.x;
  const int nq = P.nq;
  // per-thread constants of the second phase (computed while the loads above drain)
  const tcplx* Tt = T + t;
  const tcplx* Tp[8];
  treal sg[8];
  // Diagonal = detunings + interaction.  The interaction of an occupied low bit p with the occupied bits above
  // the tile, ch[tile][p], acts like one more detuning on p: 13 doubles per tile from a small table instead of
  // 8 B per amplitude from HBM; the pairs inside the low 12 bits come from a 32 KiB table that lives in L1.
  const size_t n_tiles = P.dim >> TB;
  const double* ch = P.diag + TILE + n_tiles + tile * 12;
  treal dthr = (treal)__ldg(P.diag + TILE + tile);   // bits this thread's elements share: tile bits 0-7, bits above the tile
#pragma unroll
  for (int lb = 0; lb < 8; ++lb) {
    const bool a = (t >> lb) & 1;
    Tp[lb] = T + (t ^ (1 << lb));
    sg[lb] = a ? cf.gim[lb] : -cf.gim[lb];
    dthr += a ? (treal)0 : (cf.d[lb] + (treal)__ldg(ch + lb));
  }
  treal d8[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) d8[k] = cf.d[8 + k] + (treal)__ldg(ch + 8 + k);
  for (int gb = TB; gb < nq; ++gb) dthr += ((tile >> (gb - TB)) & 1) ? (treal)0 : cf.d[gb];
  const double* dg_ptr = P.diag + t;
  const unsigned long long pol = l2_policy(P.pol_st);
  tile_sync<NBAR>();

  // ---- out = kappa * ( (Dint + detuning diagonal) Y + low-bit flips )
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const treal dg = (treal)__ldg(dg_ptr + NT * i);
    tcplx h{0.0, 0.0};
#pragma unroll
    for (int lb = 0; lb < 8; ++lb) flip_acc<REAL>(h, cf.gre[lb], sg[lb], Tp[lb][NT * i]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool a = (i >> k) & 1;
      flip_acc<REAL>(h, cf.gre[8 + k], a ? cf.gim[8 + k] : -cf.gim[8 + k], Tt[NT * (i ^ (1 << k))]);
    }
    treal dd = dthr + dg;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (!((i >> k) & 1)) dd += d8[k];
    const tcplx own = Tt[NT * i];
    h.re = fma(dd, own.re, h.re);
    h.im = fma(dd, own.im, h.im);
    st_pol(P.out + base + t + NT * i, cf.kappa * h, pol);
  }
}

template <bool REAL, bool AUX>
__device__ __forceinline__ void a_tile(const StreamParams& P, const StreamCoef& cf, tcplx* T, size_t lin_tile) {
  const int t = threadIdx.x;
  const size_t tiles_per_vec = P.dim >> TB;
  const size_t tile = lin_tile % tiles_per_vec;
  const size_t base = (lin_tile / tiles_per_vec) * P.dim + (tile << TB);   // batch column + tile

  // ---- plain application (one input, weight 1, nothing to materialise): the tile goes straight to shared
  // memory with 16 asynchronous 16-byte copies per thread in flight -- one memory latency instead of four
  // rounds of register loads
  if (!AUX && P.n_in == 1 && P.w[0] == (treal)1 && P.ymat == nullptr) {
    // (no L2 cache hint here: with the three-way policy select feeding these copies ptxas 12.9 emitted LDGSTS
    // reading uniform registers it had not set -- "illegal instruction" at run time; the input of a plain
    // application wants the normal policy anyway, the group tiles read it again)
    const tcplx* v0 = P.v[0] + base + t;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      cp_async_amp(T + t + NT * i, v0 + NT * i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    a_tile_flips<REAL, 0>(P, cf, T, tile, base);
    return;
  }
  // ---- combination: Y = sum_j w_j v_j, 4 elements x 2 vectors in flight per thread
  constexpr int QP = 4;
  constexpr bool want_aux = AUX;                   // the error-estimate vector of the last stage
  const unsigned long long pol = l2_policy(P.pol_st), pol_in = l2_policy(P.pol_in);
#pragma unroll
  for (int q0 = 0; q0 < EPT; q0 += QP) {
    tcplx y[QP], z[QP];
#pragma unroll
    for (int i = 0; i < QP; ++i) { y[i] = {0.0, 0.0}; z[i] = {0.0, 0.0}; }
    for (int j = 0; j < P.n_in; j += 2) {
      const bool two = j + 1 < P.n_in;               // uniform
      const tcplx* v0 = P.v[j] + base;
      const treal w0 = P.w[j];
      tcplx x0[QP], x1[QP];
#pragma unroll
      for (int i = 0; i < QP; ++i) x0[i] = ld_pol(v0 + t + NT * (q0 + i), pol_in);
      if (two) {
        const tcplx* v1 = P.v[j + 1] + base;
        const treal w1 = P.w[j + 1];
#pragma unroll
        for (int i = 0; i < QP; ++i) x1[i] = ld_pol(v1 + t + NT * (q0 + i), pol_in);
#pragma unroll
        for (int i = 0; i < QP; ++i) {
          y[i].re = fma(w1, x1[i].re, y[i].re); y[i].im = fma(w1, x1[i].im, y[i].im);
        }
        if (want_aux) {
          const treal u1 = P.w2[j + 1];
#pragma unroll
          for (int i = 0; i < QP; ++i) {
            z[i].re = fma(u1, x1[i].re, z[i].re); z[i].im = fma(u1, x1[i].im, z[i].im);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < QP; ++i) {
        y[i].re = fma(w0, x0[i].re, y[i].re); y[i].im = fma(w0, x0[i].im, y[i].im);
      }
      if (want_aux) {
        const treal u0 = P.w2[j];
#pragma unroll
        for (int i = 0; i < QP; ++i) {
          z[i].re = fma(u0, x0[i].re, z[i].re); z[i].im = fma(u0, x0[i].im, z[i].im);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < QP; ++i) {
      T[t + NT * (q0 + i)] = y[i];
      if (P.ymat) st_pol(P.ymat + base + t + NT * (q0 + i), y[i], pol);
      if (want_aux) P.aux[base + t + NT * (q0 + i)] = z[i];
    }
  }
  a_tile_flips<REAL, 0>(P, cf, T, tile, base);
}

template <bool REAL, bool AUX>
__global__ void __launch_bounds__(NT, REAL ? kMinCtas : 2)
k_stream_a(const __grid_constant__ StreamParams P, const __grid_constant__ StreamCoef cf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  a_tile<REAL, AUX>(P, cf, reinterpret_cast<tcplx*>(smem_raw), blockIdx.x);
}

// ---- group tile: strided tile (2^nb rows x 2^C columns), out += kappa * H_g Ymat ------------------------
// Tile element e = col | row << C; thread t owns e = t + 256 i, so consecutive elements of a thread are
// 2^(8-C) rows apart: their global indices are g0 + i * stride.  The row bits split into tile bits [C, 8)
// (partner = another thread's element, same i: per-thread pointer, runtime count 8 - C <= 4) and tile bits
// 8-11 (partner = the thread's own element i ^ 2^k).
__device__ __forceinline__ size_t gindex(const StreamParams& P, size_t tile, int e) {
  const size_t col = (size_t)(e & ((1 << P.C) - 1)), row = (size_t)(e >> P.C);
  const int lw = P.lo - P.C;                       // tile-index bits placed below the row bits
  const size_t ul = tile & (((size_t)1 << lw) - 1), uh = tile >> lw;
  return col | (ul << P.C) | (row << P.lo) | (uh << (P.lo + P.nb));
}

template <bool REAL, int NBAR>
__device__ __forceinline__ void g_tile(const StreamParams& P, const StreamCoef& cf, tcplx* T, size_t lin_tile) {
  const int t = threadIdx.x;
  const size_t tiles_per_vec = P.dim >> TB;
  const size_t tile = lin_tile % tiles_per_vec;
  const size_t boff = (lin_tile / tiles_per_vec) * P.dim;
  const int C = P.C;
  const size_t g0 = boff + gindex(P, tile, t);
  const size_t stride = (size_t)1 << (P.lo + 8 - C);
  const tcplx* ym = P.v[0] + g0;
  tcplx* out = P.out + g0;
  const unsigned long long pol_ld = l2_policy(P.pol_ld), pol_st = l2_policy(P.pol_st);

  // Ymat tile -> shared memory with 16-byte asynchronous copies (LDGSTS): 16 in flight per thread, no
  // registers; the first round of the partial result is fetched while they land
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    cp_async_amp_pol(T + t + NT * i, ym + (size_t)i * stride, pol_ld);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const tcplx* Tt = T + t;
  const int n_cross = 8 - C;                       // 0..4 row bits that live in t
  const int p8 = P.lo + n_cross;                   // global bit position of tile bit 8
  treal gre8[4], gim8[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { gre8[k] = cf.gre[p8 + k]; gim8[k] = cf.gim[p8 + k]; }
  tcplx nxt[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) nxt[j] = ld_pol(out + (size_t)j * stride, pol_ld);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  tile_sync<NBAR>();
  double err_acc = 0.0;
#pragma unroll
  for (int q0 = 0; q0 < EPT; q0 += 4) {
    tcplx acc[4], h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j] = nxt[j];
      h[j] = {0.0, 0.0};
    }
    if (q0 + 4 < EPT) {
#pragma unroll
      for (int j = 0; j < 4; ++j) nxt[j] = ld_pol(out + (size_t)(q0 + 4 + j) * stride, pol_ld);
    }
    for (int b = 0; b < n_cross; ++b) {
      const int lb = C + b;
      const bool a = (t >> lb) & 1;
      const tcplx* Tq = T + (t ^ (1 << lb));
      const treal gr = cf.gre[P.lo + b];
      const treal s = a ? cf.gim[P.lo + b] : -cf.gim[P.lo + b];
#pragma unroll
      for (int j = 0; j < 4; ++j) flip_acc<REAL>(h[j], gr, s, Tq[NT * (q0 + j)]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = q0 + j;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool a = (i >> k) & 1;
        flip_acc<REAL>(h[j], gre8[k], a ? gim8[k] : -gim8[k], Tt[NT * (i ^ (1 << k))]);
      }
      fma_acc(acc[j], cf.kappa, h[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) st_pol(out + (size_t)(q0 + j) * stride, acc[j], pol_st);
    if (P.err_partial) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t gi = g0 + (size_t)(q0 + j) * stride;
        const tcplx ep = ldcs(P.aux + gi);
        const tcplx y0 = ldcs(P.y0 + gi);
        const tcplx y1 = Tt[NT * (q0 + j)];
        // |err / (atol + rtol max(|y0|, |y1|))|^2 with one square root and one division (amplitudes are <= 1:
        // the squares neither overflow nor matter when they underflow)
        const double m2 = (double)fmax(fma(y0.re, y0.re, y0.im * y0.im), fma(y1.re, y1.re, y1.im * y1.im));
        const double sc = fma(P.rtol, sqrt(m2), P.atol);
        const double er = fma(P.werr, (double)acc[j].re, (double)ep.re), ei = fma(P.werr, (double)acc[j].im, (double)ep.im);
        err_acc += fma(er, er, ei * ei) / (sc * sc);
      }
    }
  }
  if (P.err_partial) {
    __shared__ double red[NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err_acc += __shfl_xor_sync(0xffffffffu, err_acc, o);
    if ((t & 31) == 0) red[t >> 5] = err_acc;
    tile_sync<NBAR>();
    if (t == 0) {
      double sacc = 0.0;
      for (int w = 0; w < NT / 32; ++w) sacc += red[w];
      P.err_partial[lin_tile] = sacc;
    }
  }
}

template <bool REAL>
__global__ void __launch_bounds__(NT, REAL ? kMinCtas : 2)
k_stream_g(const __grid_constant__ StreamParams P, const __grid_constant__ StreamCoef cf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  g_tile<REAL, 0>(P, cf, reinterpret_cast<tcplx*>(smem_raw), blockIdx.x);
}

// ---- dataflow launch: A tiles and the first group's tiles, chunk by chunk through L2 --------------------
// sync[0] = ticket dispenser, sync[1 + c] = finished A tiles of chunk c (zeroed by the host before the launch).
// Tickets make the order of work items the order in which CTAs START, so a group tile only ever waits for A
// tiles that are already running or done (no reliance on the block scheduler's dispatch order).
struct AgCtl {
  unsigned* sync;
  unsigned chunk_log2;     // tiles per chunk = 1 << chunk_log2; the first group's row bits lie inside a chunk
  unsigned n_chunks;
  unsigned lag;            // group tiles of chunk c are handed out after the A tiles of chunk c + lag
  unsigned mix;
  unsigned done_per_tile;  // arrivals on a chunk counter per finished A tile (1: per CTA, 16: per consumer warp)
  unsigned dbg;            // experiment switches (PD_STREAM_DBG)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <bool REAL, bool AUX>
__global__ void __launch_bounds__(NT, REAL ? kMinCtas : 2)
k_stream_ag(const __grid_constant__ StreamParams PA, const __grid_constant__ StreamParams PG,
            const __grid_constant__ StreamCoef cf, const AgCtl ctl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  tcplx* T = reinterpret_cast<tcplx*>(smem_raw);
  __shared__ unsigned s_item;
  if (threadIdx.x == 0) s_item = atomicAdd(ctl.sync, 1u);
  __syncthreads();
  const unsigned item = s_item;
  const unsigned CT = 1u << ctl.chunk_log2;
  const unsigned blk = item >> (ctl.chunk_log2 + 1);
  unsigned r = item & (2 * CT - 1);
  // ctl.mix: A and group items alternate ticket by ticket (a steady mix of HBM-bound and L2-bound CTAs on every
  // SM) instead of CT A items followed by CT group items
  bool is_a = r < CT;
  if (ctl.mix) { is_a = !(r & 1u); r = (r >> 1) + (is_a ? 0u : CT); }
  if (is_a) {
    if (blk >= ctl.n_chunks) return;
    a_tile<REAL, AUX>(PA, cf, T, (size_t)blk * CT + r);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(ctl.sync + 1 + blk, 1u);
  } else {
    if (blk < ctl.lag) return;
    const unsigned c = blk - ctl.lag;
    if (threadIdx.x == 0) {
      // bounded wait (~seconds): a lost producer must end as an error, never as a hung GPU
      unsigned spins = 0;
      while (ld_acquire_u32(ctl.sync + 1 + c) < CT) {
        __nanosleep(128);
        if (++spins > (1u << 24)) __trap();
      }
    }
    __syncthreads();
    g_tile<REAL, 0>(PG, cf, T, (size_t)c * CT + (r - CT));
  }
}

#if !defined(PD_C64)   // the experimental TMA pipeline exists for complex128 only
// ---- persistent dataflow launch with an asynchronous input pipeline (round 2) ----------------------------
// Same work items, ticket order and chunk counters as k_stream_ag, but the CTAs are persistent (one per SM, 16
// consumer warps + a producer warp) and warp-specialised:
//   * the last warp (one elected lane) is the PRODUCER: it draws the tickets, waits for a group item's chunk, hands
//     the item to the consumers through a small queue, and streams the input vectors of A items into a ring
//     of 16 KiB shared-memory slots with cp.async.bulk (TMA bulk copies, completion counted in bytes on an
//     mbarrier) -- it runs ahead of the consumers across item boundaries, so the HBM stream never waits for a
//     tile's shared-memory phase; group tiles (strided rows) travel as TMA tensor boxes (cp.async.bulk.tensor);
//   * warps 0-15 are CONSUMERS: they combine the staged inputs in registers (the thread keeps its 8 elements
//     of Y, so the three highest tile bits flip in registers), park Y in the 64 KiB tile for the nine
//     cross-thread bits, and write Ymat / the result with L2 policies.
// Ring slots are recycled through full/empty mbarriers; only the tile buffer uses a (consumer-only) named
// barrier.
constexpr int kPipeCons = 512;                       // consumer threads (16 warps)
constexpr int kPipeThreads = kPipeCons + 32;         // + the producer warp
constexpr int PEPT = TILE / kPipeCons;               // 8 elements per consumer thread: e = t + 512 i
constexpr int PXB = 9;                               // tile bits that live in the thread index (cross-thread flips)
constexpr int PRB = TB - PXB;                        // tile bits that live in the element index (flips in registers)
constexpr int kSub = 1024;                           // amplitudes per ring slot (16 KiB)
constexpr int SPT = kSub / kPipeCons;                // 2 elements per thread and slot
constexpr int NSUB = TILE / kSub;                    // 4 slots per tile and vector
constexpr int kPipeSlots = 10;
constexpr int kItemQ = 4;
constexpr unsigned kSlotBytes = kSub * 16;
struct PipeShared {
  tcplx T[TILE];
  tcplx ring[kPipeSlots][kSub];
  unsigned long long full[kPipeSlots], empty[kPipeSlots], item_full[kItemQ], item_empty[kItemQ];
  unsigned item_kind[kItemQ], item_tile[kItemQ];
  double red[kPipeCons / 32];
};
constexpr size_t kPipeSmem = sizeof(PipeShared) + 128;   // + slack to align the base to 128 B (TMA destinations)
static_assert(kPipeSmem <= 232448, "one pipeline CTA per SM");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// bounded wait (seconds): a broken pipeline must end as an error, never as a hung GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  const unsigned a = smem_u32(b);
  unsigned ok = 0, spins = 0;
  long long t0 = 0;
  for (;;) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                          unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// one box of a strided group tile (cuTensorMapEncodeTiled view, see make_group_map): rows x contiguous piece
__device__ __forceinline__ void tensor_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                               unsigned long long* bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
               " [%0], [%1, {%2, %3, %4}], [%5], %6;"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kPipeCons) : "memory"); }

// ticket -> work item (shared with k_stream_ag): kind 0 = A tile, 1 = group tile, 2 = nothing, 3 = end
struct PipeItem { unsigned kind, tile, chunk; };
__device__ __forceinline__ PipeItem decode_item(const AgCtl& ctl, unsigned item, unsigned total) {
  if (item >= total) return {3u, 0u, 0u};
  const unsigned CT = 1u << ctl.chunk_log2;
  const unsigned blk = item >> (ctl.chunk_log2 + 1);
  unsigned r = item & (2 * CT - 1);
  bool is_a = r < CT;
  if (ctl.mix) { is_a = !(r & 1u); r = (r >> 1) + (is_a ? 0u : CT); }
  if (is_a) {
    if (blk >= ctl.n_chunks) return {2u, 0u, 0u};
    return {0u, blk * CT + r, blk};
  }
  if (blk < ctl.lag) return {2u, 0u, 0u};
  const unsigned c = blk - ctl.lag;
  return {1u, c * CT + (r - CT), c};
}

// one ring slot -> this thread's SPT elements of sub-block `q` (waits for the copy, frees the slot)
__device__ __forceinline__ void ring_take(PipeShared& S, unsigned& step, tcplx (&x)[SPT]) {
  const unsigned sl = step % kPipeSlots;
  mbar_wait(&S.full[sl], (step / kPipeSlots) & 1u);
  const tcplx* src = S.ring[sl] + threadIdx.x;
#pragma unroll
  for (int ii = 0; ii < SPT; ++ii) x[ii] = src[kPipeCons * ii];
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&S.empty[sl]);
  ++step;
}

// A item: Y = sum_j w_j v_j from the staged inputs, Ymat, out = kappa (diagonal + flips of the low 12 bits) Y.
// The thread keeps its 8 elements of Y: tile bits 9-11 flip in registers, bits 0-8 through the tile buffer.
template <bool REAL, bool AUX>
__device__ __forceinline__ void pipe_a_item(const StreamParams& P, const StreamCoef& cf, PipeShared& S, size_t lin_tile,
                                            unsigned& step) {
  const int t = threadIdx.x;
  const size_t tiles_per_vec = P.dim >> TB;
  const size_t tile = lin_tile % tiles_per_vec;
  const size_t base = (lin_tile / tiles_per_vec) * P.dim + (tile << TB);
  const unsigned long long pol = l2_policy(P.pol_st);
  tcplx y[PEPT];
#pragma unroll
  for (int q = 0; q < NSUB; ++q) {
    tcplx z[SPT];
#pragma unroll
    for (int ii = 0; ii < SPT; ++ii) { y[SPT * q + ii] = {0.0, 0.0}; z[ii] = {0.0, 0.0}; }
    for (int j = 0; j < P.n_in; ++j) {
      tcplx x[SPT];
      ring_take(S, step, x);
      const double w = P.w[j];
#pragma unroll
      for (int ii = 0; ii < SPT; ++ii) {
        y[SPT * q + ii].re = fma(w, x[ii].re, y[SPT * q + ii].re);
        y[SPT * q + ii].im = fma(w, x[ii].im, y[SPT * q + ii].im);
      }
      if (AUX) {
        const double u = P.w2[j];
#pragma unroll
        for (int ii = 0; ii < SPT; ++ii) { z[ii].re = fma(u, x[ii].re, z[ii].re); z[ii].im = fma(u, x[ii].im, z[ii].im); }
      }
    }
    if (AUX) {
#pragma unroll
      for (int ii = 0; ii < SPT; ++ii) P.aux[base + t + kPipeCons * (SPT * q + ii)] = z[ii];
    }
  }
  tcplx* T = S.T;
  const tcplx* Tp[PXB];
  double sg[PXB];
  double dthr = 0.0;          // diagonal of the bits this thread's elements share: tile bits 0-8 and the bits above the tile
#pragma unroll
  for (int lb = 0; lb < PXB; ++lb) {
    const bool a = (t >> lb) & 1;
    Tp[lb] = T + (t ^ (1 << lb));
    sg[lb] = a ? cf.gim[lb] : -cf.gim[lb];
    dthr += a ? 0.0 : cf.d[lb];
  }
  for (int gb = TB; gb < P.nq; ++gb) dthr += ((tile >> (gb - TB)) & 1) ? 0.0 : cf.d[gb];
  // interaction part of the diagonal from the tile tables (see a_tile_flips)
  const double* ch = P.diag + TILE + (P.dim >> TB) + tile * 12;
  dthr += __ldg(P.diag + TILE + tile);
#pragma unroll
  for (int lb = 0; lb < PXB; ++lb) dthr += ((t >> lb) & 1) ? 0.0 : __ldg(ch + lb);
  double dr[PRB];
#pragma unroll
  for (int k = 0; k < PRB; ++k) dr[k] = cf.d[PXB + k] + __ldg(ch + PXB + k);
  const double* dg_ptr = P.diag + t;
  double dg[PEPT];
#pragma unroll
  for (int i = 0; i < PEPT; ++i) dg[i] = __ldg(dg_ptr + kPipeCons * i);
  cons_sync();                                       // every warp is done reading the previous item's tile
#pragma unroll
  for (int i = 0; i < PEPT; ++i) {
    T[t + kPipeCons * i] = y[i];
    if (P.ymat) st_pol(P.ymat + base + t + kPipeCons * i, y[i], pol);
  }
  cons_sync();
#pragma unroll
  for (int i = 0; i < PEPT; ++i) {
    tcplx h{0.0, 0.0};
#pragma unroll
    for (int lb = 0; lb < PXB; ++lb) flip_acc<REAL>(h, cf.gre[lb], sg[lb], Tp[lb][kPipeCons * i]);
#pragma unroll
    for (int k = 0; k < PRB; ++k) {
      const bool a = (i >> k) & 1;
      flip_acc<REAL>(h, cf.gre[PXB + k], a ? cf.gim[PXB + k] : -cf.gim[PXB + k], y[i ^ (1 << k)]);
    }
    double dd = dthr + dg[i];
#pragma unroll
    for (int k = 0; k < PRB; ++k)
      if (!((i >> k) & 1)) dd += dr[k];
    h.re = fma(dd, y[i].re, h.re);
    h.im = fma(dd, y[i].im, h.im);
    st_pol(P.out + base + t + kPipeCons * i, cf.kappa * h, pol);
  }
}

// group item: the Ymat tile and the partial result arrive through the ring as boxes of 2^(10-C) rows (TMA tensor
// copies issued by the producer).  Tile element e = col | row << C = t + 512 i: the row bits split into tile bits
// [C, 9) (partner = another thread's element: tile buffer) and tile bits 9-11 (the thread's own elements:
// registers).
template <bool REAL>
__device__ __forceinline__ void pipe_g_item(const StreamParams& P, const StreamCoef& cf, PipeShared& S, size_t lin_tile,
                                            unsigned& step) {
  const int t = threadIdx.x;
  const size_t tiles_per_vec = P.dim >> TB;
  const size_t tile = lin_tile % tiles_per_vec;
  const size_t boff = (lin_tile / tiles_per_vec) * P.dim;
  const int C = P.C;
  const size_t g0 = boff + gindex(P, tile, t);
  const size_t stride = (size_t)1 << (P.lo + PXB - C);
  tcplx* out = P.out + g0;
  const unsigned long long pol_st = l2_policy(P.pol_st);
  tcplx y[PEPT];
  const bool ring = P.via_ring != 0;                 // uniform
  if (ring) {
#pragma unroll
    for (int q = 0; q < NSUB; ++q) {
      tcplx x[SPT];
      ring_take(S, step, x);
#pragma unroll
      for (int ii = 0; ii < SPT; ++ii) y[SPT * q + ii] = x[ii];
    }
  } else {
    const unsigned long long pol_ld = l2_policy(P.pol_ld);
    const tcplx* ym = P.v[0] + g0;
#pragma unroll
    for (int i = 0; i < PEPT; ++i) y[i] = ld_pol(ym + (size_t)i * stride, pol_ld);
  }
  tcplx* T = S.T;
  const int n_cross = PXB - C;                       // row bits that live in t
  const int p9 = P.lo + n_cross;                     // global bit position of tile bit 9
  double gre9[PRB], gim9[PRB];
#pragma unroll
  for (int k = 0; k < PRB; ++k) { gre9[k] = cf.gre[p9 + k]; gim9[k] = cf.gim[p9 + k]; }
  // without the ring the partial result is fetched now, so that its latency hides behind the tile exchange
  tcplx accg[PEPT];
  if (!ring) {
    const unsigned long long pol_ld = l2_policy(P.pol_ld);
#pragma unroll
    for (int i = 0; i < PEPT; ++i) accg[i] = ld_pol(out + (size_t)i * stride, pol_ld);
  }
  cons_sync();
#pragma unroll
  for (int i = 0; i < PEPT; ++i) T[t + kPipeCons * i] = y[i];
  cons_sync();
  double err_acc = 0.0;
#pragma unroll
  for (int q = 0; q < NSUB; ++q) {
    tcplx acc[SPT], h[SPT];
    if (ring) ring_take(S, step, acc);
    else {
#pragma unroll
      for (int ii = 0; ii < SPT; ++ii) acc[ii] = accg[SPT * q + ii];
    }
#pragma unroll
    for (int ii = 0; ii < SPT; ++ii) h[ii] = {0.0, 0.0};
    for (int b = 0; b < n_cross; ++b) {
      const int lb = C + b;
      const bool a = (t >> lb) & 1;
      const tcplx* Tq = T + (t ^ (1 << lb));
      const double gr = cf.gre[P.lo + b];
      const double sgn = a ? cf.gim[P.lo + b] : -cf.gim[P.lo + b];
#pragma unroll
      for (int ii = 0; ii < SPT; ++ii) flip_acc<REAL>(h[ii], gr, sgn, Tq[kPipeCons * (SPT * q + ii)]);
    }
#pragma unroll
    for (int ii = 0; ii < SPT; ++ii) {
      const int i = SPT * q + ii;
#pragma unroll
      for (int k = 0; k < PRB; ++k) {
        const bool a = (i >> k) & 1;
        flip_acc<REAL>(h[ii], gre9[k], a ? gim9[k] : -gim9[k], y[i ^ (1 << k)]);
      }
      fma_acc(acc[ii], cf.kappa, h[ii]);
      st_pol(out + (size_t)i * stride, acc[ii], pol_st);
    }
    if (P.err_partial) {
#pragma unroll
      for (int ii = 0; ii < SPT; ++ii) {
        const int i = SPT * q + ii;
        const size_t gi = g0 + (size_t)i * stride;
        const tcplx ep = ldcs(P.aux + gi);
        const tcplx y0 = ldcs(P.y0 + gi);
        const double m2 = fmax(fma(y0.re, y0.re, y0.im * y0.im), fma(y[i].re, y[i].re, y[i].im * y[i].im));
        const double sc = fma(P.rtol, sqrt(m2), P.atol);
        const double er = fma(P.werr, acc[ii].re, ep.re), ei = fma(P.werr, acc[ii].im, ep.im);
        err_acc += fma(er, er, ei * ei) / (sc * sc);
      }
    }
  }
  if (P.err_partial) {
    double* red = S.red;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err_acc += __shfl_xor_sync(0xffffffffu, err_acc, o);
    cons_sync();
    if ((t & 31) == 0) red[t >> 5] = err_acc;
    cons_sync();
    if (t == 0) {
      double sacc = 0.0;
      for (int w = 0; w < kPipeCons / 32; ++w) sacc += red[w];
      P.err_partial[lin_tile] = sacc;
    }
  }
}

template <bool REAL, bool AUX>
__global__ void __launch_bounds__(kPipeThreads, 1)
k_stream_pipe(const __grid_constant__ StreamParams PA, const __grid_constant__ StreamParams PG,
              const __grid_constant__ StreamCoef cf, const AgCtl ctl, const unsigned total_items,
              const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PipeShared& S = *reinterpret_cast<PipeShared*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kPipeSlots; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], kPipeCons / 32); }
    for (int i = 0; i < kItemQ; ++i) { mbar_init(&S.item_full[i], 1); mbar_init(&S.item_empty[i], kPipeCons / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const unsigned CT = 1u << ctl.chunk_log2;
  if (warp == kPipeCons / 32) {
    // ===== producer: one lane =====
    if (lane != 0) return;
    const unsigned long long pol_in = l2_policy(2);  // inputs stream through L2 once
    const size_t tiles_per_vec = PA.dim >> TB;
    unsigned step = 0, it = 0;
    for (;;) {
      const unsigned ticket = atomicAdd(ctl.sync, 1u);
      const PipeItem w = decode_item(ctl, ticket, total_items);
      if (w.kind == 2u) continue;
      if (w.kind == 1u && ctl.done_per_tile) {
        unsigned spins = 0;
        while (ld_acquire_u32(ctl.sync + 1 + w.chunk) < CT * ctl.done_per_tile) {
          __nanosleep(64);
          if (++spins > (1u << 25)) __trap();
        }
        __threadfence();
        if (!(ctl.dbg & 2u)) asm volatile("fence.proxy.async.global;" ::: "memory");   // the chunk was written through the generic proxy, TMA reads it
      }
      const unsigned q = it % kItemQ;
      mbar_wait(&S.item_empty[q], ((it / kItemQ) & 1u) ^ 1u);
      S.item_kind[q] = w.kind;
      S.item_tile[q] = w.tile;
      mbar_arrive(&S.item_full[q]);
      ++it;
      if (w.kind == 3u) break;
      const size_t tile = (size_t)w.tile % tiles_per_vec;
      const size_t col = (size_t)w.tile / tiles_per_vec;          // batch column
      if (w.kind == 0u) {
        const size_t base = col * PA.dim + (tile << TB);
        for (int q4 = 0; q4 < NSUB; ++q4)
          for (int j = 0; j < PA.n_in; ++j) {
            const unsigned sl = step % kPipeSlots;
            mbar_wait(&S.empty[sl], ((step / kPipeSlots) & 1u) ^ 1u);
            mbar_expect_tx(&S.full[sl], kSlotBytes);
            bulk_load(S.ring[sl], PA.v[j] + base + (size_t)q4 * kSub, kSlotBytes, &S.full[sl], pol_in);
            ++step;
          }
      } else if (PG.via_ring) {
        // tile index -> box coordinates (gindex): contiguous doubles, row, slab
        const int C = PG.C, lw = PG.lo - C;
        const int c0 = (int)((tile & (((size_t)1 << lw) - 1)) << (C + 1));
        const int c2 = (int)((tile >> lw) + col * (PG.dim >> (PG.lo + PG.nb)));
        const int rows = kSub >> C;
        const unsigned long long pol_g = l2_policy(PG.pol_ld);
        for (int v = 0; v < 2; ++v)
          for (int q4 = 0; q4 < NSUB; ++q4) {
            const unsigned sl = step % kPipeSlots;
            mbar_wait(&S.empty[sl], ((step / kPipeSlots) & 1u) ^ 1u);
            mbar_expect_tx(&S.full[sl], kSlotBytes);
            if (v == 0) tensor_load_3d(S.ring[sl], &tm_y, c0, q4 * rows, c2, &S.full[sl], pol_g);
            else tensor_load_3d(S.ring[sl], &tm_out, c0, q4 * rows, c2, &S.full[sl], pol_g);
            ++step;
          }
      }
    }
    return;
  }
  // ===== consumers =====
  unsigned step = 0, it = 0;
  for (;;) {
    const unsigned q = it % kItemQ;
    mbar_wait(&S.item_full[q], (it / kItemQ) & 1u);
    const unsigned kind = S.item_kind[q], tile = S.item_tile[q];
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.item_empty[q]);
    ++it;
    if (kind == 3u) break;
    if (kind == 0u) {
      pipe_a_item<REAL, AUX>(PA, cf, S, tile, step);
      // this warp's share of the tile is written: publish it to the group tiles of the chunk
      // the group tiles read these lines through the async proxy (TMA): generic -> async proxy fence for
      // global memory (the unqualified fence.proxy.async only covers shared memory), then the usual release
      if (!(ctl.dbg & 1u)) asm volatile("fence.proxy.async.global;" ::: "memory");
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(ctl.sync + 1 + (tile >> ctl.chunk_log2), 1u);
    } else {
      pipe_g_item<REAL>(PG, cf, S, tile, step);
    }
  }
}

#endif  // !PD_C64

// SiteOps holds kappa * H per qubit (pd_common.hpp site_ops_ket): recover the Hermitian entries.  kappa is
// 1 or +-i, so multiplying by conj(kappa) is exact.
void fill_coef(const SiteOps& so, int nq, StreamCoef& c) {
  c.kappa = {(treal)so.kappa.re, (treal)so.kappa.im};
  const cplx kc = conj(so.kappa);
  const double k2 = so.kappa.re * so.kappa.re + so.kappa.im * so.kappa.im;
  if (k2 != 1.0) throw Error(PD_ERR_STATE, "stream kernels expect a unit-modulus generator scale");
  for (int q = 0; q < nq; ++q) {
    const int p = nq - 1 - q;   // global bit position of qubit q
    const cplx d = kc * so.T[q * 4 + 0], g = kc * so.T[q * 4 + 2], gc = kc * so.T[q * 4 + 1];
    if (d.im != 0.0 || g.re != gc.re || g.im != -gc.im || so.T[q * 4 + 3].re != 0.0 || so.T[q * 4 + 3].im != 0.0)
      throw Error(PD_ERR_STATE, "stream kernels expect kappa * (Hermitian site operators)");
    c.d[p] = (treal)d.re;
    c.gre[p] = (treal)g.re;
    c.gim[p] = (treal)g.im;
  }
}
bool real_drive(const StreamCoef& a, int nq) {
  for (int p = 0; p < nq; ++p)
    if (a.gim[p] != 0.0) return false;
  return true;
}

bool g_attr_set[64] = {};     // per device: function attributes belong to the device's context
int current_device() {
  int dev = 0;
  PD_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) throw Error(PD_ERR_STATE, "device index out of range");
  return dev;
}
#if !defined(PD_C64)
// 3-D view of a vector for the strided tiles of the group on bits [lo, lo + nb), in doubles (2 per amplitude):
// dim0 = 2^(lo+1) contiguous, dim1 = 2^nb rows (stride 2^lo amplitudes), dim2 = the slabs above (and the
// batch columns, which continue the same stride); box = one ring slot = 2^(10-C) rows x 2^C amplitudes.
CUtensorMap make_group_map(const tcplx* base, size_t dim, int batch, int lo, int nb) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) throw Error(PD_ERR_STATE, "cuTensorMapEncodeTiled is not available");
    return reinterpret_cast<EncodeFn>(fn);
  }();
  const int C = TB - nb;
  CUtensorMap tm;
  const cuuint64_t gdim[3] = {(cuuint64_t)2 << lo, (cuuint64_t)1 << nb, (cuuint64_t)(dim >> (lo + nb)) * (cuuint64_t)batch};
  const cuuint64_t gstride[2] = {(cuuint64_t)16 << lo, (cuuint64_t)16 << (lo + nb)};
  const cuuint32_t box[3] = {(cuuint32_t)2 << C, (cuuint32_t)1 << (10 - C), 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<tcplx*>(base), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(PD_ERR_STATE, "cuTensorMapEncodeTiled failed for a group tile view");
  return tm;
}

#endif  // !PD_C64
int sm_count() {
  static int n[64] = {};
  const int dev = current_device();
  if (!n[dev]) PD_CUDA_CHECK(cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev));
  return n[dev];
}
void set_attrs() {
  const int dev = current_device();
  if (g_attr_set[dev]) return;
  auto big = [](auto* f) { PD_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * kAmpBytes)); };
  big(k_stream_a<false, false>); big(k_stream_a<false, true>); big(k_stream_a<true, false>); big(k_stream_a<true, true>);
  big(k_stream_ag<false, false>); big(k_stream_ag<false, true>); big(k_stream_ag<true, false>); big(k_stream_ag<true, true>);
  big(k_stream_g<false>); big(k_stream_g<true>);
#if !defined(PD_C64)
  auto pipe_attr = [](auto* f) { PD_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem)); };
  pipe_attr(k_stream_pipe<false, false>); pipe_attr(k_stream_pipe<false, true>);
  pipe_attr(k_stream_pipe<true, false>); pipe_attr(k_stream_pipe<true, true>);
#endif
  g_attr_set[dev] = true;
}

// Ticket + per-chunk counters of the dataflow launches, one buffer per (device, stream): launches on one
// stream are ordered, launches on different streams must not share counters.
constexpr unsigned kMaxChunks = 1u << 16;
unsigned* ag_sync_buffer(cudaStream_t s) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, unsigned*> bufs;
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(mu);
  auto it = bufs.find({dev, s});
  if (it != bufs.end()) return it->second;
  unsigned* p = nullptr;
  PD_CUDA_CHECK(cudaMalloc(&p, sizeof(unsigned) * (kMaxChunks + 1)));
  bufs[{dev, s}] = p;
  return p;
}
unsigned env_unsigned(const char* name, unsigned dflt) {
  const char* e = std::getenv(name);
  return e ? (unsigned)std::strtoul(e, nullptr, 10) : dflt;
}
// PD_STREAM_FUSE=0 keeps the A launch and the first group launch separate (for A/B measurements)
int fuse_mode() {
  static const int mode = [] {
    const char* e = std::getenv("PD_STREAM_FUSE");
    const int m = e ? (e[0] - '0') : 1;             // 0: separate launches, 1: k_stream_ag, 2: k_stream_pipe
    return (kC64 && m == 2) ? 1 : m;                // (the pipeline is a complex128 experiment)
  }();
  return mode;
}

// the strided groups of the bits above the contiguous tile: group gi takes bits [lo, lo + nb)
struct Groups {
  int G;
  int lo[8], nb[8];
};
Groups make_groups(int nq) {
  Groups gr{};
  const int rest = nq - TB;
  gr.G = (rest + kMaxGroupBits - 1) / kMaxGroupBits;
  int lo = TB;
  for (int gi = 0; gi < gr.G; ++gi) gr.nb[gi] = rest / gr.G + (gi < rest % gr.G ? 1 : 0);
  // The FIRST group rides in the dataflow launch: its closure (2^nb tiles) is the L2 blocking unit.  With two
  // groups PD_STREAM_G1BITS moves bits between them (a group tile needs 4 <= nb <= 8: 8 - C cross bits >= 0 and
  // pieces of >= 256 B).
  // Measured at N = 24, 26 (profiles/r02_stream_knobs.md): a 6-bit first group (4 MiB closure per vector) beats
  // the even split by 3-4 %.
  static const unsigned g1_env = env_unsigned("PD_STREAM_G1BITS", 0);
  if (gr.G == 2) {
    // complex64: a 7-bit second group keeps its strided pieces at 256 B (32 amplitudes of 8 B); measured at N = 26:
    // 6 + 8 bits 6.48 ms, 7 + 7 bits 6.06 ms, 8 + 6 bits 7.07 ms per DP5 step; N = 27: 7 + 8 bits 12.9 ms, 8 + 7 bits
    // 14.3 ms (profiles/r02_c64.md)
    constexpr int kSecondGroupBits = kC64 ? 7 : kMaxGroupBits;
    int nb0 = rest >= 12 ? std::max(rest - kSecondGroupBits, 6) : gr.nb[0];
    if (nb0 > 7) nb0 = std::max(rest - kMaxGroupBits, 6);      // (an 8-bit FIRST group is the slower trade: N = 27)
    if (g1_env) nb0 = (int)g1_env;
    const int nb1 = rest - nb0;
    if (nb0 >= 4 && nb0 <= kMaxGroupBits && nb1 >= 4 && nb1 <= kMaxGroupBits) { gr.nb[0] = nb0; gr.nb[1] = nb1; }
  }
  for (int gi = 0; gi < gr.G; ++gi) { gr.lo[gi] = lo; lo += gr.nb[gi]; }
  return gr;
}

struct ErrTail {   // embedded error estimate, finished by the LAST group launch of stage 7
  tcplx* aux;
  const tcplx* y0;
  double werr, atol, rtol;
  double* err_partial;
};

// One stage: out = H_A Y (A launch; Y formed from A.v / written to A.ymat) then out += H_g ysrc per group.
// The A launch and the first group launch go out as one dataflow launch (see k_stream_ag).
int launch_stage(const Geometry& g, const StreamParams& A, const StreamCoef& cf, bool uni, const tcplx* ysrc,
                 const ErrTail* tail, cudaStream_t s) {
  const size_t n_tiles = (g.dim >> TB) * (size_t)g.batch;
  const unsigned grid = (unsigned)n_tiles;
  const Groups gr = make_groups(g.nq);
  StreamParams B[8];
  static const unsigned hints = env_unsigned("PD_STREAM_HINTS", 7);
  for (int gi = 0; gi < gr.G; ++gi) {
    StreamParams& b = B[gi];
    b = StreamParams{};
    b.nq = g.nq; b.dim = g.dim; b.n_in = 1; b.v[0] = ysrc; b.w[0] = 1.0; b.out = A.out;
    b.lo = gr.lo[gi]; b.nb = gr.nb[gi]; b.C = TB - gr.nb[gi];
    b.via_ring = 1;
    b.pol_st = hints & 2 ? 2 : 0;                    // group results: not re-read before the next launch
    b.pol_ld = (hints & 4) && gi == 0 ? 2 : 0;       // fused group: last use of the A tiles' lines
    if (tail && gi == gr.G - 1) {
      b.aux = tail->aux; b.y0 = tail->y0; b.werr = tail->werr; b.atol = tail->atol; b.rtol = tail->rtol;
      b.err_partial = tail->err_partial;
    }
  }
  int n = 0, first = 0;
  const unsigned tiles_per_vec = (unsigned)(g.dim >> TB);
  // defaults from the sweep in profiles/r02_stream_knobs.md (N=26: 14.25 ms with chunk 8 / lag 1 / no hints ->
  // 13.0 ms with chunk 6 / lag 6 / alternating tickets / L2 policies)
  static const unsigned max_chunk_log2 = env_unsigned("PD_STREAM_CHUNK", 6), lag = env_unsigned("PD_STREAM_LAG", 6);
  unsigned chunk_log2 = 0;
  while ((1u << (chunk_log2 + 1)) <= tiles_per_vec && chunk_log2 < std::max<unsigned>(max_chunk_log2, gr.nb[0])) ++chunk_log2;
  const size_t n_chunks = n_tiles >> chunk_log2;
  if (fuse_mode() != 0 && gr.G >= 1 && (unsigned)gr.nb[0] <= chunk_log2 && n_chunks >= 16 && n_chunks <= kMaxChunks) {
    static const unsigned mix = env_unsigned("PD_STREAM_MIX", 1);
    // the pipelined launch moves group tiles as TMA boxes: inner extent 2^(C+1) doubles <= 256, i.e. nb >= 5
    const bool pipe = fuse_mode() == 2 && gr.nb[0] >= 5;
    static const unsigned dbg = env_unsigned("PD_STREAM_DBG", 0);
    AgCtl ctl{ag_sync_buffer(s), chunk_log2, (unsigned)n_chunks, std::max(1u, lag), mix, pipe ? 16u : 1u, dbg};
    StreamParams Af = A;
    Af.pol_st = hints & 1 ? 1 : 0;                   // keep Ymat and the partial result in L2 for the group tiles
    static const unsigned plain_in = env_unsigned("PD_STREAM_PLAIN_IN", 0);
    Af.pol_in = (A.ymat == nullptr && A.n_in == 1) ? (int)plain_in : 2;
    PD_CUDA_CHECK(cudaMemsetAsync(ctl.sync, 0, sizeof(unsigned) * (n_chunks + 1), s));
    const unsigned ag_grid = (unsigned)(2 * (n_tiles + ((size_t)ctl.lag << chunk_log2)));
    const bool aux = A.aux != nullptr;
#if !defined(PD_C64)
    if (pipe) {
      const unsigned pgrid = (unsigned)std::min<size_t>((size_t)sm_count(), n_tiles);
      // Group tiles of the dataflow launch read what A tiles of the SAME launch wrote a moment earlier.  Moving
      // them as TMA boxes returned stale sectors now and then (N=22, batch 2, 8 chunks: 26 of 400 applications
      // wrong in a few 32 B sectors of one piece, with generic->async proxy fences on both sides; profiles/
      // r02_stream_pipe.md), so here they are fetched with ordinary loads; PD_STREAM_TMA_G=1 re-enables the boxes.
      static const unsigned tma_g = env_unsigned("PD_STREAM_TMA_G", 0);
      B[0].via_ring = tma_g ? 1 : 0;
      const CUtensorMap tm_y = make_group_map(ysrc, g.dim, g.batch, gr.lo[0], gr.nb[0]);
      const CUtensorMap tm_out = make_group_map(A.out, g.dim, g.batch, gr.lo[0], gr.nb[0]);
      auto go = [&](auto* f) { f<<<pgrid, kPipeThreads, kPipeSmem, s>>>(Af, B[0], cf, ctl, ag_grid, tm_y, tm_out); };
      if (uni) { if (aux) go(k_stream_pipe<true, true>); else go(k_stream_pipe<true, false>); }
      else { if (aux) go(k_stream_pipe<false, true>); else go(k_stream_pipe<false, false>); }
    } else
#endif
    {
      auto go = [&](auto* f) { f<<<ag_grid, NT, TILE * kAmpBytes, s>>>(Af, B[0], cf, ctl); };
      if (uni) { if (aux) go(k_stream_ag<true, true>); else go(k_stream_ag<true, false>); }
      else { if (aux) go(k_stream_ag<false, true>); else go(k_stream_ag<false, false>); }
    }
    n = 1;
    first = 1;
  } else {
    const bool aux = A.aux != nullptr;
    StreamParams An = A;
    An.pol_in = 2;
    auto go = [&](auto* f) { f<<<grid, NT, TILE * kAmpBytes, s>>>(An, cf); };
    if (uni) { if (aux) go(k_stream_a<true, true>); else go(k_stream_a<true, false>); }
    else { if (aux) go(k_stream_a<false, true>); else go(k_stream_a<false, false>); }
    n = 1;
  }
  // later groups: their inputs come from earlier LAUNCHES, so the tiles can travel as TMA boxes through the
  // persistent pipeline (every ticket is a group item; PD_STREAM_GPIPE=0 keeps the plain kernel)
  static const unsigned gpipe = env_unsigned("PD_STREAM_GPIPE", 1);
  for (int gi = first; gi < gr.G; ++gi) {
#if !defined(PD_C64)
    if (gpipe && fuse_mode() == 2 && gr.nb[gi] >= 5 && n_tiles >= 2 * (size_t)sm_count()) {
      AgCtl gctl{ag_sync_buffer(s), 0u, 0u, 0u, 0u, 0u, 0u};   // chunk = 1 tile, no A items, nothing to wait for
      PD_CUDA_CHECK(cudaMemsetAsync(gctl.sync, 0, sizeof(unsigned) * 2, s));
      const CUtensorMap tm_y = make_group_map(ysrc, g.dim, g.batch, gr.lo[gi], gr.nb[gi]);
      const CUtensorMap tm_out = make_group_map(A.out, g.dim, g.batch, gr.lo[gi], gr.nb[gi]);
      const unsigned pgrid = (unsigned)std::min<size_t>((size_t)sm_count(), n_tiles);
      const unsigned total = (unsigned)(2 * n_tiles);
      if (uni) k_stream_pipe<true, false><<<pgrid, kPipeThreads, kPipeSmem, s>>>(B[gi], B[gi], cf, gctl, total, tm_y, tm_out);
      else k_stream_pipe<false, false><<<pgrid, kPipeThreads, kPipeSmem, s>>>(B[gi], B[gi], cf, gctl, total, tm_y, tm_out);
    } else
#endif
    if (uni) k_stream_g<true><<<grid, NT, TILE * kAmpBytes, s>>>(B[gi], cf);
    else k_stream_g<false><<<grid, NT, TILE * kAmpBytes, s>>>(B[gi], cf);
    ++n;
  }
  return n;
}

}  // namespace

constexpr int kMinStreamQubits = 16;   // below this the gather kernels run out of L2 anyway

bool stream_ket_supported(const Geometry& g) {
  // the strided groups take the bits above 12 in chunks of <= 8; every chunk needs >= 1 bit
  return g.kind == PD_KET && g.nq >= kMinStreamQubits && g.nq <= kMaxQubits - 1 && g.diag_parts != nullptr &&
         (g.dim >> TB) * (size_t)g.batch < ((size_t)1 << 31);
}

// out = G (sum_j w_j in_j); ymat receives the combined input (required: the group launches read it).
int launch_stream_stage_ket(const Geometry& g, amp_t* out_, amp_t* ymat_, int n_in, const amp_t* const* ins_,
                            const double* w, const SiteOps& so, cudaStream_t s) {
  tcplx* out = tc(out_);
  tcplx* ymat = tc(ymat_);
  const tcplx* const* ins = reinterpret_cast<const tcplx* const*>(ins_);
  if (n_in > kMaxIn) throw Error(PD_ERR_INVALID, "stream stage takes at most 8 inputs");
  set_attrs();
  StreamCoef cf;
  fill_coef(so, g.nq, cf);
  const bool uni = real_drive(cf, g.nq);
  // a plain application (one input, weight 1) needs no materialised combination
  const bool plain = n_in == 1 && w[0] == 1.0 && ymat == nullptr;
  const tcplx* ysrc = plain ? ins[0] : ymat;
  if (!plain && ymat == nullptr) throw Error(PD_ERR_STATE, "stream stage needs a buffer for the combined input");
  StreamParams A{};
  A.nq = g.nq; A.dim = g.dim; A.n_in = n_in; A.diag = g.diag_parts; A.ymat = plain ? nullptr : ymat; A.out = out;
  for (int j = 0; j < n_in; ++j) { A.v[j] = ins[j]; A.w[j] = (treal)w[j]; }
  const int n = launch_stage(g, A, cf, uni, ysrc, nullptr, s);
  PD_CUDA_CHECK(cudaGetLastError());
  return n;
}


namespace {
__global__ void k_stream_sum_partials(const double* __restrict__ partial, int per_col, double* out) {
  __shared__ double sh[32];
  const int b = blockIdx.x;
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
}  // namespace

// ---- adjoint site correlations on tiles ---------------------------------------------------------------
// Per qubit q the gradient distribution (engine.hpp::distribute) needs three real sums over the state index,
//   gd_q = sum_{s: bit_q(s)=0} Im(conj(kbar[s]) y[s]),   ga_q = sum_s Im(conj(kbar[s]) y[s ^ m_q]),
//   gb_q = sum_s (bit_q(s) ? 1 : -1) Re(conj(kbar[s]) y[s ^ m_q]),
// with y the stage input (a combination of y_n and the slopes, formed on the fly like in k_stream_a and
// written once as Ymat for the group launches).  The contiguous launch covers the low 12 qubits' bits and
// every self term; each group launch its own bits.  Partners come from shared memory, not through L2 as in
// the gather kernel k_corr_ket.
namespace {
constexpr int kCA = 3 * TB + 1;          // per-CTA partial sums of the contiguous launch: 12 x (gd, ga, gb) + self
constexpr int kCG = 2 * kMaxGroupBits;   // per-CTA partial sums of a group launch: nb x (ga, gb)

struct CorrParams {
  int nq, lo, nb, C, n_in;
  size_t dim;
  const tcplx* v[kMaxIn];
  treal w[kMaxIn];
  tcplx* ymat;              // contiguous launch: combined stage input written here (nullable)
  const tcplx* ysrc;        // group launches
  const tcplx* kbar;
  double* wacc;            // [2^nq] += wscale * Im(conj(kbar) y)   (nullable; contiguous launch)
  double wscale;
  double* partial;         // [gridDim.x][kCA or kCG]
};

template <int R, class AccT>
__device__ __forceinline__ void corr_block_reduce(AccT (&acc)[R], double* partial, size_t block, size_t n_blocks) {
  __shared__ double sh[NT / 32][R];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double v = (double)acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][r] = v;
  }
  __syncthreads();
  if (threadIdx.x < R) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += sh[w][threadIdx.x];
    partial[(size_t)threadIdx.x * n_blocks + block] = s;   // [R][n_blocks]: the final pass reads rows
  }
}

// Same thread <-> element map as a_tile (e = t + 256 i): for tile bits lb < 8 the bit value of e is a property
// of the thread, so the sign of gb and the "bit clear" condition of gd are applied ONCE per thread after the
// element loop; for lb >= 8 they are compile-time constants.  Per flip and element: one LDS.128 and four FMAs
// straight into the two accumulators (ga += Im(conj(kb) y'), gb_raw += Re(conj(kb) y')).
// KEEP: the tile's lines stay in L2 (normal policy instead of streaming loads) for the first group's tile of
// the same chunk, which the interleaved launch k_stream_corr_ag schedules right behind it.
template <bool KEEP>
__device__ __forceinline__ tcplx corr_ld(const tcplx* p, unsigned long long pol) {
  if (KEEP) return ld_pol(p, pol);
  return ldcs(p);
}
template <bool KEEP>
__device__ __forceinline__ void corr_a_tile(const CorrParams& P, tcplx* T, size_t lin_tile, size_t n_blocks) {
  const int t = threadIdx.x;
  const size_t tiles_per_vec = P.dim >> TB;
  const size_t tile = lin_tile % tiles_per_vec;
  const size_t base = (lin_tile / tiles_per_vec) * P.dim + (tile << TB);
  const unsigned long long pol_keep = KEEP ? l2_policy(1) : 0ull;
  constexpr int QP = 4;
#pragma unroll
  for (int q0 = 0; q0 < EPT; q0 += QP) {
    tcplx y[QP];
#pragma unroll
    for (int i = 0; i < QP; ++i) y[i] = {0.0, 0.0};
    for (int j = 0; j < P.n_in; ++j) {
      const tcplx* vj = P.v[j] + base;
      const treal wj = P.w[j];
      tcplx x[QP];
#pragma unroll
      for (int i = 0; i < QP; ++i) x[i] = corr_ld<KEEP>(vj + t + NT * (q0 + i), pol_keep);
#pragma unroll
      for (int i = 0; i < QP; ++i) { y[i].re = fma(wj, x[i].re, y[i].re); y[i].im = fma(wj, x[i].im, y[i].im); }
    }
#pragma unroll
    for (int i = 0; i < QP; ++i) {
      T[t + NT * (q0 + i)] = y[i];
      if (P.ymat) P.ymat[base + t + NT * (q0 + i)] = y[i];
    }
  }
  const tcplx* Tt = T + t;
  const tcplx* Tp[8];
#pragma unroll
  for (int lb = 0; lb < 8; ++lb) Tp[lb] = T + (t ^ (1 << lb));
  __syncthreads();
  treal ga[TB], gb[TB], self_lo[4] = {0, 0, 0, 0}, self_all = 0;
#pragma unroll
  for (int lb = 0; lb < TB; ++lb) { ga[lb] = 0; gb[lb] = 0; }
  // four elements at a time (i = 4 io + ii): bounds the loads in flight / registers; tile bits 8, 9 are
  // compile-time inside the group, bits 10, 11 are uniform per group
#pragma unroll 1
  for (int io = 0; io < 4; ++io) {
    tcplx kb4[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) kb4[ii] = corr_ld<KEEP>(P.kbar + base + t + NT * (4 * io + ii), pol_keep);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int i = 4 * io + ii;
      const tcplx kbv = kb4[ii];                                // conj(kbar) = (kbv.re, -kbv.im)
      const tcplx own = Tt[NT * i];
      const treal self_im = fma(kbv.re, own.im, -kbv.im * own.re);
      self_all += self_im;
      if (P.wacc) atomicAdd(P.wacc + (tile << TB) + t + NT * i, P.wscale * (double)self_im);
#pragma unroll
      for (int lb = 0; lb < 8; ++lb) {
        const tcplx pv = Tp[lb][NT * i];
        ga[lb] = fma(kbv.re, pv.im, fma(-kbv.im, pv.re, ga[lb]));
        gb[lb] = fma(kbv.re, pv.re, fma(kbv.im, pv.im, gb[lb]));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool a = k < 2 ? ((ii >> k) & 1) : ((io >> (k - 2)) & 1);
        const tcplx pv = Tt[NT * (i ^ (1 << k))];
        const treal re = fma(kbv.re, pv.re, kbv.im * pv.im);
        ga[8 + k] = fma(kbv.re, pv.im, fma(-kbv.im, pv.re, ga[8 + k]));
        gb[8 + k] += a ? re : -re;
        self_lo[k] += a ? (treal)0 : self_im;
      }
    }
  }
  treal acc[kCA];
#pragma unroll
  for (int lb = 0; lb < 8; ++lb) {
    const bool a = (t >> lb) & 1;
    acc[lb * 3 + 0] = a ? (treal)0 : self_all;
    acc[lb * 3 + 1] = ga[lb];
    acc[lb * 3 + 2] = a ? gb[lb] : -gb[lb];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc[(8 + k) * 3 + 0] = self_lo[k];
    acc[(8 + k) * 3 + 1] = ga[8 + k];
    acc[(8 + k) * 3 + 2] = gb[8 + k];
  }
  acc[kCA - 1] = self_all;
  corr_block_reduce<kCA>(acc, P.partial, lin_tile, n_blocks);
}
__global__ void __launch_bounds__(NT, 2) k_stream_corr_a(const __grid_constant__ CorrParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  corr_a_tile<false>(P, reinterpret_cast<tcplx*>(smem_raw), blockIdx.x, gridDim.x);
}

__device__ __forceinline__ size_t cindex(const CorrParams& P, size_t tile, int e) {
  const size_t col = (size_t)(e & ((1 << P.C) - 1)), row = (size_t)(e >> P.C);
  const int lw = P.lo - P.C;
  const size_t ul = tile & (((size_t)1 << lw) - 1), uh = tile >> lw;
  return col | (ul << P.C) | (row << P.lo) | (uh << (P.lo + P.nb));
}

// group tile: element e = t + 256 i at global index g0 + i * stride (see g_tile); row bits [C, 8) live in t
// (runtime count 8 - C <= 4), row bits 8-11 in i.
__device__ __forceinline__ void corr_g_tile(const CorrParams& P, tcplx* T, size_t lin_tile, size_t n_blocks) {
  const int t = threadIdx.x;
  const size_t tiles_per_vec = P.dim >> TB;
  const size_t tile = lin_tile % tiles_per_vec;
  const size_t boff = (lin_tile / tiles_per_vec) * P.dim;
  const int C = P.C;
  const size_t g0 = boff + cindex(P, tile, t);
  const size_t stride = (size_t)1 << (P.lo + 8 - C);
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    cp_async_amp(T + t + NT * i, P.ysrc + g0 + (size_t)i * stride);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  tcplx kb[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) kb[i] = ldcs(P.kbar + g0 + (size_t)i * stride);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const tcplx* Tt = T + t;
  const int n_cross = 8 - C;
  treal acc[kCG];
#pragma unroll
  for (int r = 0; r < kCG; ++r) acc[r] = 0;
  for (int b = 0; b < n_cross; ++b) {
    const tcplx* Tq = T + (t ^ (1 << (C + b)));
    treal a_im = 0, a_re = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const tcplx pv = Tq[NT * i];
      a_im = fma(kb[i].re, pv.im, fma(-kb[i].im, pv.re, a_im));
      a_re = fma(kb[i].re, pv.re, fma(kb[i].im, pv.im, a_re));
    }
    const bool a = (t >> (C + b)) & 1;
    // static index into acc: b is a runtime loop counter, so scatter with a compile-time unrolled select
#pragma unroll
    for (int bb = 0; bb < 4; ++bb)
      if (bb == b) { acc[bb * 2 + 0] = a_im; acc[bb * 2 + 1] = a ? a_re : -a_re; }
  }
  // tile bits 8-11 are group bits n_cross .. n_cross + 3
  treal ga8[4] = {0, 0, 0, 0}, gb8[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const tcplx pv = Tt[NT * (i ^ (1 << k))];
      const treal re = fma(kb[i].re, pv.re, kb[i].im * pv.im);
      ga8[k] = fma(kb[i].re, pv.im, fma(-kb[i].im, pv.re, ga8[k]));
      gb8[k] += ((i >> k) & 1) ? re : -re;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int bb = 0; bb < kMaxGroupBits; ++bb)
      if (bb == n_cross + k) { acc[bb * 2 + 0] = ga8[k]; acc[bb * 2 + 1] = gb8[k]; }
  }
  corr_block_reduce<kCG>(acc, P.partial, lin_tile, n_blocks);
}
__global__ void __launch_bounds__(NT, 2) k_stream_corr_g(const __grid_constant__ CorrParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  corr_g_tile(P, reinterpret_cast<tcplx*>(smem_raw), blockIdx.x, gridDim.x);
}
// Contiguous tiles and the FIRST group's tiles of a plain stage input in one launch, chunk by chunk (a chunk =
// 2^chunk_log2 tiles, closed under the first group's bits): the group tiles of chunk c - lag alternate with the
// contiguous tiles of chunk c, so they find kbar and Y -- both only read here, no ordering needed -- in L2
// instead of fetching them from HBM a second time (6 -> 4 vector passes per stage with two groups).
__global__ void __launch_bounds__(NT, 2)
k_stream_corr_ag(const __grid_constant__ CorrParams PA, const __grid_constant__ CorrParams PG, const unsigned chunk_log2,
                 const unsigned n_chunks, const unsigned n_tiles, const unsigned lag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  tcplx* T = reinterpret_cast<tcplx*>(smem_raw);
  const unsigned CT = 1u << chunk_log2;
  const unsigned item = blockIdx.x;
  const unsigned blk = item >> (chunk_log2 + 1);
  const unsigned r = item & (2 * CT - 1);
  const bool is_a = !(r & 1u);
  if (is_a) {
    if (blk >= n_chunks) return;
    corr_a_tile<true>(PA, T, (size_t)blk * CT + (r >> 1), n_tiles);
  } else {
    if (blk < lag) return;
    corr_g_tile(PG, T, (size_t)(blk - lag) * CT + (r >> 1), n_tiles);
  }
}

// One warp per qubit bit position p: sums the per-CTA partials in a fixed order and writes the 2x2 block
// d_corr[q][a][a'] that engine.hpp::distribute reads: c[0] = (0, gd), c[2] = (gb, ga), c[1] = c[3] = 0.
struct CorrFinal {
  int nq, n_groups;
  int lo[8], nb[8];
  const double* part_a;
  const double* part_g[8];
  unsigned nblocks;
  unsigned tiles_per_vec;
  cplx* d_corr;
};
__global__ void __launch_bounds__(256) k_stream_corr_final(const __grid_constant__ CorrFinal F) {
  const int p = blockIdx.x, t = threadIdx.x;
  double gd = 0.0, ga = 0.0, gb = 0.0;
  const size_t nbk = F.nblocks;
  if (p < TB) {
    const double* r0 = F.part_a + (size_t)(p * 3) * nbk;
    for (unsigned b = t; b < F.nblocks; b += 256) {
      gd += r0[b]; ga += r0[nbk + b]; gb += r0[2 * nbk + b];
    }
  } else {
    // self terms: CTAs whose tile index has bit (p - TB) clear
    const double* rs = F.part_a + (size_t)(kCA - 1) * nbk;
    for (unsigned b = t; b < F.nblocks; b += 256)
      if (!(((b % F.tiles_per_vec) >> (p - TB)) & 1u)) gd += rs[b];
    for (int g = 0; g < F.n_groups; ++g)
      if (p >= F.lo[g] && p < F.lo[g] + F.nb[g]) {
        const int j = p - F.lo[g];
        const double* r0 = F.part_g[g] + (size_t)(j * 2) * nbk;
        for (unsigned b = t; b < F.nblocks; b += 256) {
          ga += r0[b]; gb += r0[nbk + b];
        }
      }
  }
  __shared__ double sh[8][3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gd += __shfl_xor_sync(0xffffffffu, gd, o);
    ga += __shfl_xor_sync(0xffffffffu, ga, o);
    gb += __shfl_xor_sync(0xffffffffu, gb, o);
  }
  if ((t & 31) == 0) { sh[t >> 5][0] = gd; sh[t >> 5][1] = ga; sh[t >> 5][2] = gb; }
  __syncthreads();
  if (t == 0) {
    gd = ga = gb = 0.0;
    for (int w = 0; w < 8; ++w) { gd += sh[w][0]; ga += sh[w][1]; gb += sh[w][2]; }
    const int q = F.nq - 1 - p;
    F.d_corr[q * 4 + 0] = cplx{0.0, gd};
    F.d_corr[q * 4 + 1] = cplx{0.0, 0.0};
    F.d_corr[q * 4 + 2] = cplx{gb, ga};
    F.d_corr[q * 4 + 3] = cplx{0.0, 0.0};
  }
}
}  // namespace

// Site correlations of kbar with the stage input y = sum_j w_j in_j (ymat: buffer for y, required unless the
// input is plain).  d_corr (nullable): [nq][4] complex; d_wacc (nullable): [2^nq] += wscale * Im(conj(kbar) y).
int launch_stream_corr(const Geometry& g, cplx* d_corr, double* d_wacc, double wscale, const amp_t* kbar_, int n_in,
                       const amp_t* const* ins_, const double* w, amp_t* ymat_, cudaStream_t s) {
  const tcplx* kbar = tc(kbar_);
  tcplx* ymat = tc(ymat_);
  const tcplx* const* ins = reinterpret_cast<const tcplx* const*>(ins_);
  if (n_in > kMaxIn) throw Error(PD_ERR_INVALID, "stream corr takes at most 8 inputs");
  static bool attr_set[64] = {};
  const int dev = current_device();
  if (!attr_set[dev]) {
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_stream_corr_a, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * kAmpBytes));
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_stream_corr_g, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * kAmpBytes));
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_stream_corr_ag, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE * kAmpBytes));
    attr_set[dev] = true;
  }
  const unsigned grid = (unsigned)((g.dim >> TB) * (size_t)g.batch);
  const Groups gr = make_groups(g.nq);
  const int G = d_corr ? gr.G : 0;
  // per-device scratch for the per-CTA partial sums (a process may drive several devices)
  static double* parts[64] = {};
  static size_t caps[64] = {};
  const size_t need = (size_t)grid * (kCA + (size_t)G * kCG);
  if (caps[dev] < need) {
    if (parts[dev]) cudaFree(parts[dev]);
    PD_CUDA_CHECK(cudaMalloc(&parts[dev], sizeof(double) * need));
    caps[dev] = need;
  }
  double* d_part = parts[dev];
  const bool plain = n_in == 1 && w[0] == 1.0;
  const tcplx* ysrc = plain ? ins[0] : ymat;
  if (!plain && G > 0 && ymat == nullptr) throw Error(PD_ERR_STATE, "stream corr needs a buffer for the stage input");
  CorrParams A{};
  A.nq = g.nq; A.dim = g.dim; A.n_in = n_in; A.kbar = kbar; A.wacc = d_wacc; A.wscale = wscale; A.partial = d_part;
  for (int j = 0; j < n_in; ++j) { A.v[j] = ins[j]; A.w[j] = (treal)w[j]; }
  A.ymat = (plain || G == 0) ? nullptr : ymat;
  CorrParams B[8];
  CorrFinal F{};
  F.nq = g.nq; F.n_groups = G; F.part_a = d_part; F.nblocks = grid; F.tiles_per_vec = (unsigned)(g.dim >> TB);
  F.d_corr = d_corr;
  for (int gi = 0; gi < G; ++gi) {
    B[gi] = CorrParams{};
    B[gi].nq = g.nq; B[gi].dim = g.dim; B[gi].ysrc = ysrc; B[gi].kbar = kbar;
    B[gi].lo = gr.lo[gi]; B[gi].nb = gr.nb[gi]; B[gi].C = TB - gr.nb[gi];
    B[gi].partial = d_part + (size_t)grid * (kCA + (size_t)gi * kCG);
    F.lo[gi] = gr.lo[gi]; F.nb[gi] = gr.nb[gi]; F.part_g[gi] = B[gi].partial;
  }
  // a plain stage input (the adjoint sweep hands over the stage inputs the forward launches wrote): contiguous
  // tiles + first group in one L2-blocked launch; PD_CORR_FUSE=0 keeps them separate (A/B measurements)
  static const unsigned fuse = env_unsigned("PD_CORR_FUSE", 1), chunk_env = env_unsigned("PD_CORR_CHUNK", 6),
                        lag = std::max(1u, env_unsigned("PD_CORR_LAG", 4));
  const unsigned tiles_per_vec = (unsigned)(g.dim >> TB);
  unsigned chunk_log2 = 0;
  while ((1u << (chunk_log2 + 1)) <= tiles_per_vec && chunk_log2 < std::max<unsigned>(chunk_env, G ? gr.nb[0] : 0)) ++chunk_log2;
  const size_t n_chunks = (size_t)grid >> chunk_log2;
  int n = 0, first = 0;
  if (fuse && plain && G >= 1 && (unsigned)gr.nb[0] <= chunk_log2 && n_chunks >= 16) {
    const unsigned ag_grid = (unsigned)(2 * ((size_t)grid + ((size_t)lag << chunk_log2)));
    k_stream_corr_ag<<<ag_grid, NT, TILE * kAmpBytes, s>>>(A, B[0], chunk_log2, (unsigned)n_chunks, grid, lag);
    n = 1;
    first = 1;
  } else {
    k_stream_corr_a<<<grid, NT, TILE * kAmpBytes, s>>>(A);
    n = 1;
  }
  if (d_corr) {
    for (int gi = first; gi < G; ++gi) {
      k_stream_corr_g<<<grid, NT, TILE * kAmpBytes, s>>>(B[gi]);
      ++n;
    }
    k_stream_corr_final<<<g.nq, 256, 0, s>>>(F);
    ++n;
  }
  PD_CUDA_CHECK(cudaGetLastError());
  return n;
}

size_t stream_err_partial_count(const Geometry& g) { return (g.dim >> TB) * (size_t)g.batch; }

// One Dormand-Prince step with the stream kernels: stages 2..7 (k[0] = f(t, y) on entry, FSAL),
// ynew = y_{n+1}, and the embedded error estimate folded into the stage-7 launches (the A launch
// forms the partial error vector from the slopes it reads anyway, the last group launch finishes
// it): err_out[b] = sum |err/scale|^2 per batch column.  aux: scratch vector.
int launch_stream_dp5_step(const Geometry& g, const amp_t* y_, amp_t* const* k_, amp_t* ynew_, amp_t* ymat_, amp_t* aux_,
                           const SiteOps* stage_ops /* [7], index i = stage i+1 */, const double* beta,
                           const double* ew, double dt, double atol, double rtol, double* err_partial,
                           double* err_out, cudaStream_t s) {
  const tcplx* y = tc(y_);
  tcplx* const* k = reinterpret_cast<tcplx* const*>(k_);
  tcplx* ynew = tc(ynew_);
  tcplx* ymat = tc(ymat_);
  tcplx* aux = tc(aux_);
  set_attrs();
  int n = 0;
  for (int i = 1; i < 7; ++i) {
    StreamCoef cf;
    fill_coef(stage_ops[i], g.nq, cf);
    const bool uni = real_drive(cf, g.nq);
    const bool last = i == 6;
    tcplx* ym = last ? ynew : ymat;
    StreamParams A{};
    A.nq = g.nq; A.dim = g.dim; A.diag = g.diag_parts; A.ymat = ym; A.out = k[i];
    int m = 0;
    A.v[m] = y; A.w[m] = (treal)1; A.w2[m] = (treal)0; ++m;
    for (int j = 0; j < i; ++j) {
      const double b = beta[(i - 1) * 6 + j];
      if (b != 0.0 || (last && ew[j] != 0.0)) { A.v[m] = k[j]; A.w[m] = (treal)(dt * b); A.w2[m] = (treal)(last ? ew[j] : 0.0); ++m; }
    }
    A.n_in = m;
    A.aux = last ? aux : nullptr;
    const ErrTail tail{aux, y, ew[6], atol, rtol, err_partial};
    n += launch_stage(g, A, cf, uni, ym, last ? &tail : nullptr, s);
  }
  k_stream_sum_partials<<<g.batch, 256, 0, s>>>(err_partial, (int)(g.dim >> TB), err_out);
  PD_CUDA_CHECK(cudaGetLastError());
  return n + 1;
}

}  // namespace pd
