// "Gather" kernel family: one thread per amplitude, bit-flip partners through L1/L2.
// See cuda_backend.cuh for where this sits.  Everything here is the matrix-free form of
//   H(t) psi = 2*int_mat psi + sum(det terms) + sum(amp terms)      (reference hamiltonian.py:536-544)
// and of the Lindblad right-hand side, written as
//   out[idx] = kappa*Dstat(idx)*v[idx] + sum_q sum_p' T_q[p(idx)][p'] v[idx with site q := p'].
#include <cstdlib>

#include "cuda_backend.cuh"

namespace pd {

namespace {

struct PtrW {
  const amp_t* p[8];
  double w[8];
  int n;
};
struct CW {
  cplx w[128];
  int m;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction of R per-thread accumulators -> partial[blockIdx][R]
template <int R>
__device__ __forceinline__ void block_reduce_write(double (&acc)[R], double* partial) {
  __shared__ double sh[kThreads / 32][R];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    double v = warp_sum(acc[r]);
    if (lane == 0) sh[warp][r] = v;
  }
  __syncthreads();
  if (threadIdx.x < R) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += sh[w][threadIdx.x];
    partial[(size_t)blockIdx.x * R + threadIdx.x] = s;
  }
}

// out[r] (= or +=) scale * sum_b partial[b][r]; one warp per r, fixed order
__global__ void k_reduce_final(const double* __restrict__ partial, int nblocks, int R, int pstride,
                               double* __restrict__ out, int out_stride, double scale, int accumulate) {
  int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (r >= R) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partial[(size_t)b * pstride + r];
  s = warp_sum(s);
  if (lane == 0) {
    if (accumulate) out[(size_t)r * out_stride] += scale * s;
    else out[(size_t)r * out_stride] = scale * s;
  }
}

int finalize(const double* partial, int nblocks, int R, double* out, int out_stride, double scale,
             int accumulate, cudaStream_t s, int pstride = 0) {
  int wpb = 4;
  k_reduce_final<<<(R + wpb - 1) / wpb, wpb * 32, 0, s>>>(partial, nblocks, R, pstride ? pstride : R,
                                                         out, out_stride, scale, accumulate);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1;
}

int grid_for(size_t n, int per_thread = 1) {
  size_t b = (n + (size_t)kThreads * per_thread - 1) / ((size_t)kThreads * per_thread);
  return (int)std::min<size_t>(std::max<size_t>(b, 1), (size_t)148 * 16);
}
int rgrid_for(size_t n, int R, int ny) {
  size_t b = (n + kThreads - 1) / kThreads;
  size_t cap = (size_t)kMaxReduceBlocks * kMaxR * 2 / ((size_t)R * ny);
  cap = std::min<size_t>(cap, kMaxReduceBlocks);
  return (int)std::max<size_t>(1, std::min(b, cap));
}

__global__ void __launch_bounds__(kThreads) k_lincomb(amp_t* out, PtrW a, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double re = 0.0, im = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < a.n) {
        cplx v = a.p[j][i];
        re = fma(a.w[j], v.re, re);
        im = fma(a.w[j], v.im, im);
      }
    out[i] = cplx{re, im};
  }
}

__global__ void __launch_bounds__(kThreads) k_lincomb_c(amp_t* out, const amp_t* basis, size_t stride_v,
                                                        const __grid_constant__ CW cw, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    cplx acc{0.0, 0.0};
    for (int j = 0; j < cw.m; ++j) fma_acc(acc, cw.w[j], basis[(size_t)j * stride_v + i]);
    out[i] = acc;
  }
}

__global__ void __launch_bounds__(kThreads)
k_apply_ket(amp_t* __restrict__ out, const amp_t* __restrict__ in, const double* __restrict__ diag,
            const __grid_constant__ SiteOps so, int nq, size_t dim, size_t total) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    size_t s = idx & (dim - 1);
    cplx v = in[idx];
    double dg = diag[s];
    cplx dsum{so.kappa.re * dg, so.kappa.im * dg};
    cplx acc{0.0, 0.0};
    for (int q = 0; q < nq; ++q) {
      size_t m = (size_t)1 << (nq - 1 - q);
      bool a = (s & m) != 0;
      cplx td = a ? so.T[q * 4 + 3] : so.T[q * 4 + 0];
      cplx to = a ? so.T[q * 4 + 2] : so.T[q * 4 + 1];
      dsum = dsum + td;
      fma_acc(acc, to, in[idx ^ m]);
    }
    fma_acc(acc, dsum, v);
    out[idx] = acc;
  }
}

__global__ void __launch_bounds__(kThreads)
k_apply_density(amp_t* __restrict__ out, const amp_t* __restrict__ in, const double* __restrict__ diag,
                const __grid_constant__ SiteOpsDensity so, int nq, size_t total, int need_both) {
  __shared__ cplx T[kMaxSitesDensity * 16];
  for (int i = threadIdx.x; i < nq * 16; i += blockDim.x) T[i] = so.T[i];
  __syncthreads();
  size_t S = (size_t)1 << nq, dim = S * S;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    size_t e = idx & (dim - 1);
    size_t r = e >> nq, c = e & (S - 1);
    cplx v = in[idx];
    double dg = diag[r] - diag[c];
    cplx dsum{so.kappa.re * dg, so.kappa.im * dg};
    cplx acc{0.0, 0.0};
    for (int q = 0; q < nq; ++q) {
      size_t mc = (size_t)1 << (nq - 1 - q), mr = mc << nq;
      int p = ((e & mr) ? 2 : 0) | ((e & mc) ? 1 : 0);
      const cplx* Tp = &T[q * 16 + p * 4];
      dsum = dsum + Tp[p];
      fma_acc(acc, Tp[p ^ 2], in[idx ^ mr]);
      fma_acc(acc, Tp[p ^ 1], in[idx ^ mc]);
      if (need_both) fma_acc(acc, Tp[p ^ 3], in[idx ^ mr ^ mc]);
    }
    fma_acc(acc, dsum, v);
    out[idx] = acc;
  }
}

// Generator-specific Lindblad application (the default for density plans): same result as k_apply_density from the
// pieces of the super-operator (pd_common.hpp LindbladForm) instead of a general 4x4 matrix per site.
// k_apply_density issues ~760 thread instructions per entry (per site: four coefficient LDS.128 with a data-dependent
// address, three partner loads -- the double flip whether or not its coefficient is zero -- and twelve DFMA) and is
// bound by that, not by bytes (profiles/r02_lindblad.md).  Here per site: two partner loads and 2 DADD + 2 DFMA for a
// phase-free drive (coefficients from the constant bank, uniform index), the dissipator's diagonal from four popcounts,
// its off-diagonal entries only at the sites that feed them (bit scan over the matching sites).
// UD: every site has the same detuning coefficient (a global channel): its diagonal comes from two popcounts.
// IdxT: 32-bit entry indices while 4^N * batch < 2^31 (one IMAD.WIDE per partner address).
template <bool REAL, bool UD, class IdxT>
__global__ void __launch_bounds__(kThreads)
k_apply_lindblad(amp_t* __restrict__ out, const amp_t* __restrict__ in, const double* __restrict__ diag,
                 const __grid_constant__ LindbladForm lf, size_t total, int chunk_log2) {
  const int nq = lf.nq;
  const IdxT S = (IdxT)1 << nq, dim = S * S;
  const unsigned smask = (unsigned)(S - 1);
  // a CTA walks chunks of 2^chunk_log2 consecutive entries (kThreads at a time), so the column-bit partners inside
  // the chunk are lines the same CTA touches a moment earlier or later (L1) instead of L2 round trips
  const size_t n_items = total >> 8;                       // groups of kThreads = 256 entries (total is 4^N * batch)
  const size_t per_chunk = (size_t)1 << (chunk_log2 - 8);
  for (size_t it = (size_t)blockIdx.x * per_chunk; it < n_items; it += (size_t)gridDim.x * per_chunk)
   for (size_t sub = 0; sub < per_chunk && it + sub < n_items; ++sub) {
    const IdxT idx = (IdxT)(((it + sub) << 8) + threadIdx.x);
    const IdxT e = idx & (dim - 1);
    const unsigned r = (unsigned)(e >> nq), c = (unsigned)(e & (S - 1));
    const cplx v = in[idx];
    double dg = diag[r] - diag[c];
    if (UD) dg += lf.d[0] * (double)(__popc(~r & smask) - __popc(~c & smask));
    double sre = 0.0, sim = 0.0;
#pragma unroll 4
    for (int b = 0; b < nq; ++b) {
      const IdxT mc = (IdxT)1 << b, mr = mc << nq;
      const bool a = (r >> b) & 1u, cb = (c >> b) & 1u;
      const cplx pr = in[idx ^ mr], pc = in[idx ^ mc];
      const double gr = lf.gre[b];
      if (!UD) {
        const double db = lf.d[b];
        dg += (a ? 0.0 : db) - (cb ? 0.0 : db);
      }
      sre = fma(gr, pr.re - pc.re, sre);
      sim = fma(gr, pr.im - pc.im, sim);
      if (!REAL) {
        const double gi = lf.gim[b];
        const double tr = (a ? pr.re : -pr.re) + (cb ? pc.re : -pc.re);
        const double ti = (a ? pr.im : -pr.im) + (cb ? pc.im : -pc.im);
        sre = fma(-gi, ti, sre);
        sim = fma(gi, tr, sim);
      }
    }
    const cplx h{fma(dg, v.re, sre), fma(dg, v.im, sim)};
    cplx acc = lf.kappa * h;
    const int n3 = __popc(r & c), n2 = __popc(r & ~c & smask), n1 = __popc(~r & c & smask), n0 = nq - n1 - n2 - n3;
    const cplx dsum{n0 * lf.dd[0].re + n1 * lf.dd[1].re + n2 * lf.dd[2].re + n3 * lf.dd[3].re,
                    n0 * lf.dd[0].im + n1 * lf.dd[1].im + n2 * lf.dd[2].im + n3 * lf.dd[3].im};
    fma_acc(acc, dsum, v);
    for (int k = 0; k < lf.n_off; ++k) {
      const int p = lf.off_p[k], x = p ^ lf.off_pp[k];
      unsigned sel = ((p & 2) ? r : ~r) & ((p & 1) ? c : ~c) & smask;     // sites whose (a, c) pair is p
      const cplx coef = lf.off_c[k];
      while (sel) {
        const int b = __ffs((int)sel) - 1;
        sel &= sel - 1;
        const IdxT m = ((x & 2) ? ((IdxT)1 << (b + nq)) : 0) | ((x & 1) ? ((IdxT)1 << b) : 0);
        fma_acc(acc, coef, in[idx ^ m]);
      }
    }
    out[idx] = acc;
   }
}

// Column-tile variant of k_apply_density, OPT-IN (PD_DENSITY_CT=1; measured slower, profiles/r02_lindblad.md).
// A CTA stages 2^10 CONTIGUOUS entries in shared memory: every flip whose mask lies inside the tile -- the column
// bits of the last 10 sites (and, for small registers, low row bits) -- is served from shared memory, only the
// remaining partners travel through L2, and the (row, column) double flip is fetched only where its coefficient
// is non-zero: N = 12 with dephasing + relaxation fetches 1 + 17 entries per entry through L2 instead of 37.
// On B200 it runs 0.86 ms per application against 0.67 ms for the plain kernel: the plain kernel's low column
// partners already hit L1 (38 % L1 hit rate), both kernels wait on long-scoreboard stalls with ~50 % issue
// utilisation at half occupancy (54 registers), and the tile adds two barriers per 1024 entries.
constexpr int kDTB = 10;
constexpr int kDTile = 1 << kDTB;
constexpr int kDEpt = kDTile / kThreads;
__global__ void __launch_bounds__(kThreads)
k_apply_density_ct(amp_t* __restrict__ out, const amp_t* __restrict__ in, const double* __restrict__ diag,
                   const __grid_constant__ SiteOpsDensity so, int nq, size_t n_tiles) {
  __shared__ cplx T[kMaxSitesDensity * 16];
  __shared__ amp_t tile[kDTile];
  for (int i = threadIdx.x; i < nq * 16; i += blockDim.x) T[i] = so.T[i];
  const size_t S = (size_t)1 << nq, dim = S * S;
  const int t = threadIdx.x;
  for (size_t ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {
    const size_t base = ti << kDTB;
    __syncthreads();                              // the previous tile's readers are done (and T is in place)
#pragma unroll
    for (int i = 0; i < kDEpt; ++i) tile[t + kThreads * i] = in[base + t + kThreads * i];
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < kDEpt; ++i) {
      const int el = t + kThreads * i;
      const size_t idx = base + el;
      const size_t e = idx & (dim - 1);
      const size_t r = e >> nq, c = e & (S - 1);
      const cplx v = tile[el];
      const double dg = diag[r] - diag[c];
      cplx dsum{so.kappa.re * dg, so.kappa.im * dg};
      cplx acc{0.0, 0.0};
#pragma unroll 4
      for (int q = 0; q < nq; ++q) {
        const size_t mc = (size_t)1 << (nq - 1 - q), mr = mc << nq;
        const int p = ((e & mr) ? 2 : 0) | ((e & mc) ? 1 : 0);
        const cplx* Tp = &T[q * 16 + p * 4];
        dsum = dsum + Tp[p];
        const cplx prow = mr < (size_t)kDTile ? (cplx)tile[el ^ (int)mr] : (cplx)in[idx ^ mr];
        const cplx pcol = mc < (size_t)kDTile ? (cplx)tile[el ^ (int)mc] : (cplx)in[idx ^ mc];
        fma_acc(acc, Tp[p ^ 2], prow);
        fma_acc(acc, Tp[p ^ 1], pcol);
        if ((so.nzmask[q] >> (p * 4 + (p ^ 3))) & 1u) {
          const size_t mb = mr | mc;
          const cplx pboth = mb < (size_t)kDTile ? (cplx)tile[el ^ (int)mb] : (cplx)in[idx ^ mb];
          fma_acc(acc, Tp[p ^ 3], pboth);
        }
      }
      fma_acc(acc, dsum, v);
      out[idx] = acc;
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_build_diag(double* diag, int nq, const double* __restrict__ u) {
  __shared__ double su[kMaxQubits * kMaxQubits];
  for (int i = threadIdx.x; i < nq * nq; i += blockDim.x) su[i] = u[i];
  __syncthreads();
  size_t dim = (size_t)1 << nq;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < dim; s += stride) {
    double acc = 0.0;
    for (int i = 0; i < nq; ++i) {
      if (s >> (nq - 1 - i) & 1) continue;  // bit 1 = ground: r_i = 0
      for (int j = i + 1; j < nq; ++j)
        if (!(s >> (nq - 1 - j) & 1)) acc += su[i * nq + j];
    }
    diag[s] = acc;
  }
}

__global__ void __launch_bounds__(kThreads)
k_scaled_sumsq(const amp_t* __restrict__ x, const amp_t* __restrict__ xsub, const amp_t* __restrict__ ref,
               double atol, double rtol, size_t dim, double* partial) {
  size_t base = (size_t)blockIdx.y * dim;
  double acc[1] = {0.0};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
    cplx v = x[base + i];
    if (xsub) v = v - xsub[base + i];
    cplx rf = ref[base + i];
    double sc = atol + rtol * hypot(rf.re, rf.im);
    double a = v.re / sc, b = v.im / sc;
    acc[0] += a * a + b * b;
  }
  block_reduce_write<1>(acc, partial + (size_t)blockIdx.y * gridDim.x);
}

struct KPtr7 {
  const amp_t* k[7];
  double ew[7];
};
__global__ void __launch_bounds__(kThreads)
k_err_sumsq(KPtr7 kp, const amp_t* __restrict__ y0, const amp_t* __restrict__ y1, double atol,
            double rtol, size_t dim, double* partial) {
  size_t base = (size_t)blockIdx.y * dim;
  double acc[1] = {0.0};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += stride) {
    double er = 0.0, ei = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (kp.ew[j] != 0.0) {
        cplx v = kp.k[j][base + i];
        er = fma(kp.ew[j], v.re, er);
        ei = fma(kp.ew[j], v.im, ei);
      }
    cplx a = y0[base + i], b = y1[base + i];
    double sc = atol + rtol * fmax(hypot(a.re, a.im), hypot(b.re, b.im));
    er /= sc; ei /= sc;
    acc[0] += er * er + ei * ei;
  }
  block_reduce_write<1>(acc, partial + (size_t)blockIdx.y * gridDim.x);
}

__global__ void __launch_bounds__(kThreads)
k_re_dot(const amp_t* __restrict__ a, const amp_t* __restrict__ b, size_t n, double* partial) {
  double acc[1] = {0.0};
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    cplx u = a[i], v = b[i];
    acc[0] += u.re * v.re + u.im * v.im;
  }
  block_reduce_write<1>(acc, partial);
}

// Site correlations for QC ket sites per blockIdx.y:
//   C_q[a][a'] = sum_{b, s: bit_q(s)=a} conj(kbar[b,s]) * y[b, s with bit_q := a']
constexpr int kQC = 4;
__global__ void __launch_bounds__(kThreads)
k_corr_ket(const amp_t* __restrict__ kbar, const amp_t* __restrict__ y, int nq, size_t dim, int batch,
           double* partial, double* wacc, double wscale) {
  int q0 = blockIdx.y * kQC;
  double acc[kQC * 8];
#pragma unroll
  for (int i = 0; i < kQC * 8; ++i) acc[i] = 0.0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < dim; s += stride) {
    double wsum = 0.0;
    for (int b = 0; b < batch; ++b) {
      size_t idx = (size_t)b * dim + s;
      cplx kb = conj(kbar[idx]);
      cplx ys = y[idx];
      cplx self = kb * ys;
      wsum += self.im;
#pragma unroll
      for (int j = 0; j < kQC; ++j) {
        int q = q0 + j;
        if (q < nq) {
          size_t m = (size_t)1 << (nq - 1 - q);
          bool a = (s & m) != 0;
          cplx fl = kb * y[idx ^ m];
          // T layout [a][a']: slot 0 = C[0][0], 1 = C[0][1], 2 = C[1][0], 3 = C[1][1]
          acc[j * 8 + 0] += a ? 0.0 : self.re;
          acc[j * 8 + 1] += a ? 0.0 : self.im;
          acc[j * 8 + 2] += a ? 0.0 : fl.re;
          acc[j * 8 + 3] += a ? 0.0 : fl.im;
          acc[j * 8 + 4] += a ? fl.re : 0.0;
          acc[j * 8 + 5] += a ? fl.im : 0.0;
          acc[j * 8 + 6] += a ? self.re : 0.0;
          acc[j * 8 + 7] += a ? self.im : 0.0;
        }
      }
    }
    if (wacc && blockIdx.y == 0) wacc[s] += wscale * wsum;
  }
  block_reduce_write<kQC * 8>(acc, partial + (size_t)blockIdx.y * gridDim.x * (kQC * 8));
}

// All density sites in ONE pass: per site only the three real sums the gradient distribution reads
// (engine.hpp::distribute, density branch) are accumulated,
//   gd_q = sum_idx ((a==0) - (b==0)) Im(self),  ga_q = sum_idx Im(frow) - Im(fcol),
//   gb_q = sum_idx (a ? 1 : -1) Re(frow) + (b ? 1 : -1) Re(fcol)
// with a / b the row / column bit of site q, self = conj(kbar) y, frow / fcol = conj(kbar) times y
// with the row / column bit flipped.  (A per-site 4x4 correlation kernel, the first version, cost one
// pass over the 4^N vector per site.)
constexpr int kDF = 3 * kMaxSitesDensity;
// MAXQ: sites the instantiation keeps accumulators for (3 per site, in registers); two CTAs per SM (128 registers:
// three CTAs spill, one more register than 128 halves the occupancy).  Partial layout [block][3 * MAXQ].
template <int MAXQ>
__global__ void __launch_bounds__(kThreads, 2)
k_corr_density_fused(const amp_t* __restrict__ kbar, const amp_t* __restrict__ y, int nq, size_t total,
                     double* partial, double* wacc, double wscale, int chunk_log2) {
  constexpr int kAcc = 3 * MAXQ;
  size_t S = (size_t)1 << nq, dim = S * S;
  double acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
  // chunk_log2 >= 8 (total is a multiple of 256): a CTA walks 2^chunk_log2 consecutive entries, kThreads at a time
  // (column partners inside the chunk are lines it touches anyway); chunk_log2 == 0: plain grid-stride walk
  const size_t n_items = chunk_log2 ? total >> 8 : 0;
  const size_t per_chunk = chunk_log2 ? (size_t)1 << (chunk_log2 - 8) : 1;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t it = chunk_log2 ? (size_t)blockIdx.x * per_chunk : (size_t)blockIdx.x * blockDim.x + threadIdx.x;
       it < (chunk_log2 ? n_items : total); it += chunk_log2 ? (size_t)gridDim.x * per_chunk : stride)
   for (size_t sub = 0; sub < per_chunk && (!chunk_log2 || it + sub < n_items); ++sub) {
    const size_t idx = chunk_log2 ? ((it + sub) << 8) + threadIdx.x : it;
    size_t e = idx & (dim - 1);
    cplx kb = conj(kbar[idx]);
    cplx self = kb * y[idx];
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      if (q < nq) {
        size_t mc = (size_t)1 << (nq - 1 - q), mr = mc << nq;
        bool a = (e & mr) != 0, b = (e & mc) != 0;
        // Im(kb (yr - yc)) and Re(kb (sa yr + sb yc)) from the combined partners: 4 additions + 4 FMAs per site
        // instead of two complex products
        const cplx yr = y[idx ^ mr], yc = y[idx ^ mc];
        const double dre = yr.re - yc.re, dim_ = yr.im - yc.im;
        const double tre = (a ? yr.re : -yr.re) + (b ? yc.re : -yc.re);
        const double tim = (a ? yr.im : -yr.im) + (b ? yc.im : -yc.im);
        acc[q * 3 + 0] += (a == b) ? 0.0 : (a ? -self.im : self.im);
        acc[q * 3 + 1] = fma(kb.re, dim_, fma(kb.im, dre, acc[q * 3 + 1]));
        acc[q * 3 + 2] = fma(kb.re, tre, fma(-kb.im, tim, acc[q * 3 + 2]));
      }
    }
    if (wacc) {
      double w = wscale * self.im;
      atomicAdd(&wacc[e >> nq], w);
      atomicAdd(&wacc[e & (S - 1)], -w);
    }
   }
  block_reduce_write<kAcc>(acc, partial);
}
// Column-tile variant (same sums, partial layout [block][3 * 16]): the y tile of 2^10 contiguous entries sits in shared
// memory, so the column-bit partners of the last 10 sites never go through L2 (N = 12: 2 + 14 fetches per entry
// instead of 2 + 24).
__global__ void __launch_bounds__(kThreads)
k_corr_density_ct(const amp_t* __restrict__ kbar, const amp_t* __restrict__ y, int nq, size_t n_tiles,
                  double* partial, double* wacc, double wscale) {
  __shared__ amp_t tile[kDTile];
  const size_t S = (size_t)1 << nq, dim = S * S;
  const int t = threadIdx.x;
  double acc[kDF];
#pragma unroll
  for (int i = 0; i < kDF; ++i) acc[i] = 0.0;
  for (size_t ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {
    const size_t base = ti << kDTB;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kDEpt; ++i) tile[t + kThreads * i] = y[base + t + kThreads * i];
    __syncthreads();
    for (int i = 0; i < kDEpt; ++i) {
      const int el = t + kThreads * i;
      const size_t idx = base + el;
      const size_t e = idx & (dim - 1);
      const cplx kb = conj(kbar[idx]);
      const cplx self = kb * (cplx)tile[el];
#pragma unroll
      for (int q = 0; q < kMaxSitesDensity; ++q) {
        if (q < nq) {
          const size_t mc = (size_t)1 << (nq - 1 - q), mr = mc << nq;
          const bool a = (e & mr) != 0, b = (e & mc) != 0;
          const cplx yr = mr < (size_t)kDTile ? (cplx)tile[el ^ (int)mr] : (cplx)y[idx ^ mr];
          const cplx yc = mc < (size_t)kDTile ? (cplx)tile[el ^ (int)mc] : (cplx)y[idx ^ mc];
          const cplx frow = kb * yr, fcol = kb * yc;
          acc[q * 3 + 0] += ((a ? 0.0 : 1.0) - (b ? 0.0 : 1.0)) * self.im;
          acc[q * 3 + 1] += frow.im - fcol.im;
          acc[q * 3 + 2] += (a ? frow.re : -frow.re) + (b ? fcol.re : -fcol.re);
        }
      }
      if (wacc) {
        const double w = wscale * self.im;
        atomicAdd(&wacc[e >> nq], w);
        atomicAdd(&wacc[e & (S - 1)], -w);
      }
    }
  }
  block_reduce_write<kDF>(acc, partial);
}
// d_corr[q][16]: zero except the three entries engine.hpp::distribute turns back into (gd, ga, gb)
__global__ void k_corr_density_scatter(const double* __restrict__ partial, int nblocks, int nq, cplx* d_corr,
                                       int pstride) {
  int q = blockIdx.x, lane = threadIdx.x;   // one warp per site
  double v[3] = {0.0, 0.0, 0.0};
  for (int b = lane; b < nblocks; b += 32)
    for (int k = 0; k < 3; ++k) v[k] += partial[(size_t)b * pstride + q * 3 + k];
  for (int k = 0; k < 3; ++k) v[k] = warp_sum(v[k]);
  if (lane < 16) d_corr[q * 16 + lane] = cplx{0.0, 0.0};
  __syncwarp();
  if (lane == 0) {
    d_corr[q * 16 + 5] = cplx{0.0, v[0]};      // p = (a=0,b=1): weight +1 on Im C[p][p]
    d_corr[q * 16 + 2] = cplx{-v[2], v[1]};    // p = 0, prow = 2: ga = +Im, gb = -Re
  }
}

__global__ void __launch_bounds__(kThreads)
k_pair_reduce(const double* __restrict__ wacc, int nq, double* out) {
  int i = blockIdx.x / nq, j = blockIdx.x % nq;
  __shared__ double sh[kThreads / 32];
  double acc = 0.0;
  if (i < j) {
    size_t dim = (size_t)1 << nq;
    size_t mi = (size_t)1 << (nq - 1 - i), mj = (size_t)1 << (nq - 1 - j);
    for (size_t s = threadIdx.x; s < dim; s += blockDim.x)
      if (!(s & mi) && !(s & mj)) acc += wacc[s];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += sh[w];
    out[blockIdx.x] = s;
  }
}

// blockIdx.y = time index.  ket: sum_{b,s} obs[s] |psi|^2.  density: sum_r obs[r] rho[r][r].
__global__ void __launch_bounds__(kThreads)
k_expect_diag(const amp_t* __restrict__ states, const double* __restrict__ obs, int kind, int nq,
              size_t dim, int batch, double* partial) {
  double acc[2] = {0.0, 0.0};
  size_t base = (size_t)blockIdx.y * dim * batch;
  size_t S = (size_t)1 << nq;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  if (kind == PD_KET) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim * batch; i += stride) {
      cplx v = states[base + i];
      acc[0] += obs[i & (dim - 1)] * (v.re * v.re + v.im * v.im);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < S * batch; i += stride) {
      size_t b = i >> nq, r = i & (S - 1);
      cplx v = states[base + b * dim + r * S + r];
      acc[0] += obs[r] * v.re;
      acc[1] += obs[r] * v.im;
    }
  }
  block_reduce_write<2>(acc, partial + (size_t)blockIdx.y * gridDim.x * 2);
}

}  // namespace

int launch_lincomb(const Geometry& g, amp_t* out, int n_in, const amp_t* const* ins, const double* w,
                   cudaStream_t s) {
  if (n_in < 1 || n_in > 8) throw Error(PD_ERR_INVALID, "lincomb takes 1..8 inputs");
  PtrW a{};
  a.n = n_in;
  for (int i = 0; i < n_in; ++i) { a.p[i] = ins[i]; a.w[i] = w[i]; }
  size_t n = g.dim * g.batch;
  k_lincomb<<<grid_for(n, 2), kThreads, 0, s>>>(out, a, n);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_lincomb_c(size_t n, amp_t* out, int m, const amp_t* basis, size_t stride, const cplx* ws,
                     cudaStream_t s) {
  if (m < 1 || m > 128) throw Error(PD_ERR_INVALID, "lincomb_c takes 1..128 vectors");
  CW cw{};
  cw.m = m;
  for (int i = 0; i < m; ++i) cw.w[i] = ws[i];
  k_lincomb_c<<<grid_for(n, 2), kThreads, 0, s>>>(out, basis, stride, cw, n);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_apply_ket(const Geometry& g, amp_t* out, const amp_t* in, const SiteOps& so, cudaStream_t s) {
  size_t total = g.dim * g.batch;
  k_apply_ket<<<grid_for(total), kThreads, 0, s>>>(out, in, g.diag, so, g.nq, g.dim, total);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_apply_density(const Geometry& g, amp_t* out, const amp_t* in, const SiteOpsDensity& so,
                         cudaStream_t s) {
  size_t total = g.dim * g.batch;
  int need_both = 0;
  for (int q = 0; q < so.nsites; ++q)
    for (int p = 0; p < 4; ++p)
      if (so.nzmask[q] >> (p * 4 + (p ^ 3)) & 1) need_both = 1;
  // PD_LINDBLAD_FORM=0 keeps the general 4x4 kernels (A/B measurements)
  static const bool use_form = [] { const char* e = std::getenv("PD_LINDBLAD_FORM"); return !e || e[0] != '0'; }();
  if (use_form && so.form.ok && so.form.nq == g.nq && total >= 256) {
    static const int chunk_env = [] { const char* e = std::getenv("PD_LINDBLAD_CHUNK"); return e ? std::atoi(e) : 12; }();
    int chunk_log2 = std::max(8, std::min(chunk_env, 2 * g.nq));
    while (chunk_log2 > 8 && (total >> chunk_log2) < (size_t)148 * 6) --chunk_log2;   // keep every SM busy
    const int grid = (int)std::min<size_t>(total >> chunk_log2, (size_t)148 * 16);
    bool ud = true;
    for (int b = 1; b < g.nq; ++b) ud = ud && so.form.d[b] == so.form.d[0];
    static const bool narrow_ok = [] { const char* e = std::getenv("PD_LINDBLAD_IDX32"); return !e || e[0] != '0'; }();
    const bool narrow = narrow_ok && total < ((size_t)1 << 31);
    auto go = [&](auto* f) { f<<<grid, kThreads, 0, s>>>(out, in, g.diag, so.form, total, chunk_log2); };
    if (narrow) {
      if (so.form.real_drive) { if (ud) go(k_apply_lindblad<true, true, unsigned>); else go(k_apply_lindblad<true, false, unsigned>); }
      else { if (ud) go(k_apply_lindblad<false, true, unsigned>); else go(k_apply_lindblad<false, false, unsigned>); }
    } else {
      if (so.form.real_drive) { if (ud) go(k_apply_lindblad<true, true, size_t>); else go(k_apply_lindblad<true, false, size_t>); }
      else { if (ud) go(k_apply_lindblad<false, true, size_t>); else go(k_apply_lindblad<false, false, size_t>); }
    }
    PD_CUDA_CHECK(cudaGetLastError());
    return 1;
  }
  // PD_DENSITY_CT=1 selects the column-tile kernel (A/B measurements; slower on B200)
  static const bool use_ct = [] { const char* e = std::getenv("PD_DENSITY_CT"); return e && e[0] == '1'; }();
  if (use_ct && total >= (size_t)kDTile && total % kDTile == 0) {
    const size_t n_tiles = total >> kDTB;
    const int grid = (int)std::min<size_t>(n_tiles, (size_t)148 * 8);
    k_apply_density_ct<<<grid, kThreads, 0, s>>>(out, in, g.diag, so, g.nq, n_tiles);
    PD_CUDA_CHECK(cudaGetLastError());
    return 1;
  }
  k_apply_density<<<grid_for(total), kThreads, 0, s>>>(out, in, g.diag, so, g.nq, total, need_both);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1;
}

// Dint split for 4096-amplitude tiles (layout: pd_common.hpp Geometry::diag_parts), read off the full diagonal:
// bit value 1 = ground state = no interaction, so setting a group of bits switches its qubits off.
__global__ void __launch_bounds__(kThreads) k_build_diag_parts(double* parts, int nq, const double* __restrict__ diag) {
  const size_t T = (size_t)1 << (nq - 12), dim = (size_t)1 << nq;
  const size_t total = 4096 + 13 * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    if (i < 4096) parts[i] = diag[(dim - 4096) | i];
    else if (i < 4096 + T) parts[i] = diag[((i - 4096) << 12) | 0xFFF];
    else {
      const size_t j = i - 4096 - T, h = j / 12;
      const unsigned p = (unsigned)(j % 12);
      parts[i] = diag[(h << 12) | (0xFFFu & ~(1u << p))] - diag[(h << 12) | 0xFFF];
    }
  }
}
int launch_build_diag_parts(double* parts, int nq, const double* diag, cudaStream_t s) {
  const size_t total = 4096 + 13 * ((size_t)1 << (nq - 12));
  k_build_diag_parts<<<grid_for(total), kThreads, 0, s>>>(parts, nq, diag);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_build_diag(double* diag, int nq, const double* d_pair_u, cudaStream_t s) {
  k_build_diag<<<grid_for((size_t)1 << nq), kThreads, 0, s>>>(diag, nq, d_pair_u);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_scaled_sumsq(const Geometry& g, double* out, const amp_t* x, const amp_t* xsub,
                        const amp_t* ref, double atol, double rtol, double* scratch, cudaStream_t s) {
  int gx = rgrid_for(g.dim, 1, g.batch);
  k_scaled_sumsq<<<dim3(gx, g.batch), kThreads, 0, s>>>(x, xsub, ref, atol, rtol, g.dim, scratch);
  PD_CUDA_CHECK(cudaGetLastError());
  int n = 1;
  for (int b = 0; b < g.batch; ++b) n += finalize(scratch + (size_t)b * gx, gx, 1, out + b, 1, 1.0, 0, s);
  return n;
}

int launch_err_sumsq(const Geometry& g, double* out, const amp_t* const* k, const double* ew,
                     const amp_t* y0, const amp_t* y1, double atol, double rtol, double* scratch,
                     cudaStream_t s) {
  KPtr7 kp{};
  for (int j = 0; j < 7; ++j) { kp.k[j] = k[j]; kp.ew[j] = ew[j]; }
  int gx = rgrid_for(g.dim, 1, g.batch);
  k_err_sumsq<<<dim3(gx, g.batch), kThreads, 0, s>>>(kp, y0, y1, atol, rtol, g.dim, scratch);
  PD_CUDA_CHECK(cudaGetLastError());
  int n = 1;
  for (int b = 0; b < g.batch; ++b) n += finalize(scratch + (size_t)b * gx, gx, 1, out + b, 1, 1.0, 0, s);
  return n;
}

int launch_re_dot(const Geometry& g, double* out, const amp_t* a, const amp_t* b, double* scratch,
                  cudaStream_t s) {
  size_t n = g.dim * g.batch;
  int gx = rgrid_for(n, 1, 1);
  k_re_dot<<<gx, kThreads, 0, s>>>(a, b, n, scratch);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1 + finalize(scratch, gx, 1, out, 1, 1.0, 0, s);
}

int launch_corr(const Geometry& g, cplx* d_corr, double* d_wacc, double wscale, const amp_t* kbar,
                const amp_t* y, double* scratch, cudaStream_t s) {
  int n = 0;
  if (g.kind == PD_KET) {
    int ny = (g.nq + kQC - 1) / kQC;
    if (!d_corr) ny = 1;
    int gx = rgrid_for(g.dim, kQC * 8, ny);
    k_corr_ket<<<dim3(gx, ny), kThreads, 0, s>>>(kbar, y, g.nq, g.dim, g.batch, scratch, d_wacc, wscale);
  PD_CUDA_CHECK(cudaGetLastError());
    ++n;
    if (d_corr)
      for (int c = 0; c < ny; ++c) {
        int sites = std::min(kQC, g.nq - c * kQC);
        n += finalize(scratch + (size_t)c * gx * (kQC * 8), gx, sites * 8,
                      (double*)(d_corr + (size_t)c * kQC * 4), 1, 1.0, 0, s, kQC * 8);
      }
  } else {
    size_t total = g.dim * g.batch;
    int gx = rgrid_for(total, kDF, 1);
    int pstride = kDF;
    static const bool use_ct = [] { const char* e = std::getenv("PD_DENSITY_CT"); return e && e[0] == '1'; }();
    if (use_ct && total >= (size_t)kDTile && total % kDTile == 0) {
      const size_t n_tiles = total >> kDTB;
      gx = (int)std::min<size_t>(n_tiles, (size_t)gx);
      k_corr_density_ct<<<gx, kThreads, 0, s>>>(kbar, y, g.nq, n_tiles, scratch, d_wacc, wscale);
    } else {
      static const int chunk_env = [] { const char* e = std::getenv("PD_LINDBLAD_CHUNK"); return e ? std::atoi(e) : 12; }();
      int chunk_log2 = (chunk_env > 0 && total % 256 == 0 && total >= 256) ? std::max(8, std::min(chunk_env, 2 * g.nq)) : 0;
      while (chunk_log2 > 8 && (total >> chunk_log2) < (size_t)gx) --chunk_log2;
      if (g.nq <= 12) {
        pstride = 36;
        k_corr_density_fused<12><<<gx, kThreads, 0, s>>>(kbar, y, g.nq, total, scratch, d_wacc, wscale, chunk_log2);
      } else {
        k_corr_density_fused<16><<<gx, kThreads, 0, s>>>(kbar, y, g.nq, total, scratch, d_wacc, wscale, chunk_log2);
      }
    }
    PD_CUDA_CHECK(cudaGetLastError());
    ++n;
    if (d_corr) {
      k_corr_density_scatter<<<g.nq, 32, 0, s>>>(scratch, gx, g.nq, d_corr, pstride);
      PD_CUDA_CHECK(cudaGetLastError());
      ++n;
    }
  }
  return n;
}

int launch_pair_reduce(const Geometry& g, double* d_pair, const double* d_wacc, cudaStream_t s) {
  k_pair_reduce<<<g.nq * g.nq, kThreads, 0, s>>>(d_wacc, g.nq, d_pair);
  PD_CUDA_CHECK(cudaGetLastError());
  return 1;
}

int launch_expect_diag(const Geometry& g, const amp_t* states, int n_t, const double* obs, cplx* out,
                       double* scratch, cudaStream_t s) {
  size_t work = g.kind == PD_KET ? g.dim * g.batch : ((size_t)1 << g.nq) * g.batch;
  int n = 0;
  // time indices in slabs so the partial buffer stays within the reduce scratch
  int slab = std::max(1, std::min(n_t, kMaxReduceBlocks * kMaxR / (2 * 64)));
  for (int t0 = 0; t0 < n_t; t0 += slab) {
    int nt = std::min(slab, n_t - t0);
    int gx = rgrid_for(work, 2, nt);
    gx = std::min(gx, 64);
    k_expect_diag<<<dim3(gx, nt), kThreads, 0, s>>>(states + (size_t)t0 * g.dim * g.batch, obs, g.kind,
                                                    g.nq, g.dim, g.batch, scratch);
  PD_CUDA_CHECK(cudaGetLastError());
    ++n;
    for (int t = 0; t < nt; ++t)
      n += finalize(scratch + (size_t)t * gx * 2, gx, 2, (double*)(out + t0 + t), 1, 1.0, 0, s);
  }
  return n;
}

}  // namespace pd
