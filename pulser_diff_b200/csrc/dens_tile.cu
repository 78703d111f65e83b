// Tile kernels for the Lindblad generator on vec(rho) ("density tiles", N = 8..13 sites).
//
// vec(rho) index e = r << N | c (row bits high).  The generator is a sum of 4x4 site super-operators
// F_q acting on the (column bit, row bit) pair of site q (pd_common.hpp site_ops_density: -i[H, .] plus the
// dissipator of the collapse operators, reference hamiltonian.py:98-143 + backend.py:502-509) and a static
// diagonal kappa (Dint[r] - Dint[c]).  The gather kernel (k_apply_density) fetches 3 partners per site and
// entry through L1/L2: 36 x 16 B per entry at N = 12, which makes it L2-bandwidth bound.
//
// Here one launch closes BOTH bits of up to six sites inside a 4096-entry tile:
//   * the tile is a set of 12 index bits: C contiguous low bits (>= 4, so every global access is a piece of
//     >= 256 B) plus strided bits; the first tile type takes column bits 0-5 and row bits 0-5 (six complete
//     sites), the later types 4-8 passive low column bits plus the column and row bits of two to four sites;
//   * a thread owns 16 entries whose index differs in the four bits of TWO complete sites: those sites are
//     applied in registers (partners and the uniform 4x4 coefficients cost no shared-memory traffic); the
//     other sites of the tile read their partners, and the three off-diagonal coefficients of the thread's own
//     (c, r) value, from shared memory;
//   * the FIRST launch of a stage forms the stage combination Y = sum_j w_j v_j on the fly (written once as
//     Ymat for the later launches) and carries every diagonal term; later launches do out += F_sites Ymat.
// A stage is 2 launches at N <= 10 and 3 at N = 11..13, each a pure stream.
#include <cstdlib>
#include <cstring>

#include "cuda_backend.cuh"

namespace pd {

namespace {

constexpr int DT_BITS = 12;
constexpr int DT_TILE = 1 << DT_BITS;
constexpr int DT_NT = 256;
constexpr int DT_EPT = DT_TILE / DT_NT;   // 16
constexpr int DT_MAXIN = 8;

struct DensGeom {
  int n_free;
  int free_gbit[20];     // global bits outside the tile, ascending (deposit of the tile number)
  int tile_gbit[DT_BITS];   // global bit of tile bit j
  int reg_tb[4];         // tile bits that make the register index: (c_a, r_a, c_b, r_b)
  int thr_tb[8];         // tile bits that make the thread index, ascending
  int n_reg_sites;       // 1 or 2 complete sites on the register bits
  int reg_q[2];          // their site indices (T lookup)
  int n_thr_sites;       // <= 4 complete sites on thread bits
  int thr_q[4], thr_c[4], thr_r[4];   // site index, thread-index bit of its column / row bit
  int n_out_sites;       // FIRST launch: sites with both bits outside the tile (diagonal terms only)
  int out_q[16], out_p[16];           // site index, bit position p (column bit p, row bit N + p)
};

struct DensParams {
  int nq, n_in, need_both;
  int gather_rest;       // FIRST launch also applies the sites outside the tile, partners fetched through L2 from
                         // ysrc (one launch per application; needs the stage input materialised beforehand)
  size_t dim;
  unsigned tiles_per_vec;
  const cplx* v[DT_MAXIN];
  double w[DT_MAXIN];
  cplx* ymat;            // FIRST: combined input written here (nullable)
  const cplx* ysrc;      // later launches: the stage input
  cplx* out;
  const double* diag;
};

__device__ __forceinline__ cplx ld_stream(const cplx* p) {
  double2 v = __ldcs(reinterpret_cast<const double2*>(p));
  return {v.x, v.y};
}

template <bool FIRST>
__global__ void __launch_bounds__(DT_NT, 2)
k_dens_tile(const __grid_constant__ DensParams P, const __grid_constant__ DensGeom G,
            const __grid_constant__ SiteOpsDensity so) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Y = reinterpret_cast<cplx*>(smem_raw);
  cplx* Ts = Y + DT_TILE;
  const int t = threadIdx.x;
  for (int i = t; i < P.nq * 16; i += DT_NT) Ts[i] = so.T[i];
  const unsigned tile = blockIdx.x % P.tiles_per_vec, col = blockIdx.x / P.tiles_per_vec;
  size_t base = 0;
  for (int j = 0; j < G.n_free; ++j) base |= (size_t)((tile >> j) & 1u) << G.free_gbit[j];
  size_t gthr = 0;
  int sthr = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int b = (t >> j) & 1;
    gthr |= (size_t)b << G.tile_gbit[G.thr_tb[j]];
    sthr |= b << G.thr_tb[j];
  }
  size_t greg[4];
  int sreg[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    greg[b] = (size_t)1 << G.tile_gbit[G.reg_tb[b]];
    sreg[b] = 1 << G.reg_tb[b];
  }
  const size_t e0 = base | gthr;                     // index inside the vector of the thread's element 0
  const size_t g0 = (size_t)col * P.dim + e0;
  auto goff = [&](int i) -> size_t {
    return ((i & 1) ? greg[0] : 0) + ((i & 2) ? greg[1] : 0) + ((i & 4) ? greg[2] : 0) + ((i & 8) ? greg[3] : 0);
  };
  auto soff = [&](int i) -> int {
    return ((i & 1) ? sreg[0] : 0) | ((i & 2) ? sreg[1] : 0) | ((i & 4) ? sreg[2] : 0) | ((i & 8) ? sreg[3] : 0);
  };

  // ---- the tile: stage combination (FIRST) or the materialised stage input
  cplx y[DT_EPT];
  if (FIRST) {
#pragma unroll
    for (int q0 = 0; q0 < DT_EPT; q0 += 4) {
      cplx a4[4];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) a4[ii] = {0.0, 0.0};
      for (int j = 0; j < P.n_in; ++j) {
        const cplx* vj = P.v[j] + g0;
        const double wj = P.w[j];
        cplx x[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) x[ii] = ld_stream(vj + goff(q0 + ii));
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          a4[ii].re = fma(wj, x[ii].re, a4[ii].re);
          a4[ii].im = fma(wj, x[ii].im, a4[ii].im);
        }
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        y[q0 + ii] = a4[ii];
        Y[sthr | soff(q0 + ii)] = a4[ii];
        if (P.ymat) P.ymat[g0 + goff(q0 + ii)] = a4[ii];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < DT_EPT; ++i) y[i] = ld_stream(P.ysrc + g0 + goff(i));
#pragma unroll
    for (int i = 0; i < DT_EPT; ++i) Y[sthr | soff(i)] = y[i];
  }
  __syncthreads();

  // ---- per-thread constants: the off-diagonal coefficients of the sites on thread bits
  // (the coefficients themselves stay in shared memory: 24 more doubles per thread would spill)
  int crow[4];                // offset of row p of the site's 4x4 block in Ts
  int pk = 0;                 // the thread's (c, r) value per site, two bits each
  int xc[4], xr[4];
  cplx dthr{0.0, 0.0};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (s < G.n_thr_sites) {
      const int p = ((t >> G.thr_c[s]) & 1) | (((t >> G.thr_r[s]) & 1) << 1);
      crow[s] = G.thr_q[s] * 16 + p * 4;
      pk |= p << (2 * s);
      if (FIRST) dthr = dthr + Ts[crow[s] + p];
      xc[s] = 1 << G.thr_tb[G.thr_c[s]];
      xr[s] = 1 << G.thr_tb[G.thr_r[s]];
    }
  }
  if (FIRST) {
    for (int s = 0; s < G.n_out_sites; ++s) {
      const int pb = G.out_p[s];
      const int p = (int)((base >> pb) & 1) | ((int)((base >> (P.nq + pb)) & 1) << 1);
      dthr = dthr + Ts[G.out_q[s] * 16 + p * 5];
    }
  }
  const cplx* Ta = Ts + G.reg_q[0] * 16;
  const cplx* Tb = Ts + G.reg_q[1] * 16;
  const bool two = G.n_reg_sites > 1;
  const bool both = P.need_both != 0;
  const size_t cmask = ((size_t)1 << P.nq) - 1;

#pragma unroll
  for (int i = 0; i < DT_EPT; ++i) {
    cplx acc{0.0, 0.0};
    if (!FIRST) acc = ld_stream(P.out + g0 + goff(i));
    const int pa = i & 3, pb = i >> 2;
    // sites on register bits: partners are this thread's own elements
    fma_acc(acc, Ta[pa * 4 + (pa ^ 1)], y[i ^ 1]);
    fma_acc(acc, Ta[pa * 4 + (pa ^ 2)], y[i ^ 2]);
    if (both) fma_acc(acc, Ta[pa * 4 + (pa ^ 3)], y[i ^ 3]);
    if (two) {
      fma_acc(acc, Tb[pb * 4 + (pb ^ 1)], y[i ^ 4]);
      fma_acc(acc, Tb[pb * 4 + (pb ^ 2)], y[i ^ 8]);
      if (both) fma_acc(acc, Tb[pb * 4 + (pb ^ 3)], y[i ^ 12]);
    }
    // sites on thread bits: partners from the shared tile
    const int sb = sthr | soff(i);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < G.n_thr_sites) {
        const int p = (pk >> (2 * s)) & 3;
        fma_acc(acc, Ts[crow[s] + (p ^ 1)], Y[sb ^ xc[s]]);
        fma_acc(acc, Ts[crow[s] + (p ^ 2)], Y[sb ^ xr[s]]);
        if (both) fma_acc(acc, Ts[crow[s] + (p ^ 3)], Y[sb ^ xc[s] ^ xr[s]]);
      }
    }
    if (FIRST && P.gather_rest) {
      // sites outside the tile: their (c, r) value is the same for the whole CTA, so the coefficients are
      // uniform; the partners are the same element of three other tiles
      const cplx* src = P.ysrc + g0 + goff(i);
      for (int s = 0; s < G.n_out_sites; ++s) {
        const int pbit = G.out_p[s];
        const int p = (int)((base >> pbit) & 1) | ((int)((base >> (P.nq + pbit)) & 1) << 1);
        const cplx* Tp = Ts + G.out_q[s] * 16 + p * 4;
        const long long dc = (p & 1) ? -((long long)1 << pbit) : ((long long)1 << pbit);
        const long long dr = (p & 2) ? -((long long)1 << (P.nq + pbit)) : ((long long)1 << (P.nq + pbit));
        fma_acc(acc, Tp[p ^ 1], ld_stream(src + dc));
        fma_acc(acc, Tp[p ^ 2], ld_stream(src + dr));
        if (both) fma_acc(acc, Tp[p ^ 3], ld_stream(src + dc + dr));
      }
    }
    if (FIRST) {
      const size_t e = e0 + goff(i);
      const double dg = __ldg(P.diag + (e >> P.nq)) - __ldg(P.diag + (e & cmask));
      cplx dd = dthr + Ta[pa * 5];
      if (two) dd = dd + Tb[pb * 5];
      dd.re = fma(so.kappa.re, dg, dd.re);
      dd.im = fma(so.kappa.im, dg, dd.im);
      fma_acc(acc, dd, y[i]);
    }
    P.out[g0 + goff(i)] = acc;
  }
}

struct TileTypes {
  int n;
  DensGeom g[4];
};

// bit position p = N - 1 - q: column bit p, row bit N + p
void finish_geom(DensGeom& G, int nq, bool first) {
  bool in_tile[40] = {};
  for (int j = 0; j < DT_BITS; ++j) in_tile[G.tile_gbit[j]] = true;
  G.n_free = 0;
  for (int b = 0; b < 2 * nq; ++b)
    if (!in_tile[b]) G.free_gbit[G.n_free++] = b;
  G.n_out_sites = 0;
  if (first)
    for (int p = 0; p < nq; ++p)
      if (!in_tile[p] && !in_tile[nq + p]) {
        G.out_q[G.n_out_sites] = nq - 1 - p;
        G.out_p[G.n_out_sites] = p;
        ++G.n_out_sites;
      }
}

TileTypes make_tile_types(int nq) {
  TileTypes tt{};
  // type 0: sites p = 0..5; tile bits 0-5 = column bits 0-5, tile bits 6-11 = row bits 0-5; the register index
  // takes (c4, r4, c5, r5), so a thread's lanes keep column bits 0-3 contiguous (256 B pieces per 16 lanes)
  {
    DensGeom& G = tt.g[tt.n++];
    for (int j = 0; j < 6; ++j) { G.tile_gbit[j] = j; G.tile_gbit[6 + j] = nq + j; }
    const int reg[4] = {4, 10, 5, 11};
    const int thr[8] = {0, 1, 2, 3, 6, 7, 8, 9};
    std::memcpy(G.reg_tb, reg, sizeof(reg));
    std::memcpy(G.thr_tb, thr, sizeof(thr));
    G.n_reg_sites = 2;
    G.reg_q[0] = nq - 1 - 4;
    G.reg_q[1] = nq - 1 - 5;
    G.n_thr_sites = 4;
    for (int j = 0; j < 4; ++j) { G.thr_q[j] = nq - 1 - j; G.thr_c[j] = j; G.thr_r[j] = 4 + j; }
    finish_geom(G, nq, true);
  }
  // later types: the remaining sites p = 6..N-1 in groups of 2..4 behind C = 12 - 2k passive column bits
  const int rest = nq - 6;
  int sizes[2] = {0, 0};
  if (rest <= 4) sizes[0] = rest;
  else if (rest == 5) { sizes[0] = 3; sizes[1] = 2; }
  else if (rest == 6) { sizes[0] = 3; sizes[1] = 3; }
  else { sizes[0] = 4; sizes[1] = 3; }
  int p0 = 6;
  for (int gi = 0; gi < 2 && sizes[gi] > 0; ++gi) {
    const int k = sizes[gi], C = DT_BITS - 2 * k;
    DensGeom& G = tt.g[tt.n++];
    // passive bits: the lowest C index bits that do not belong to the group's sites
    for (int j = 0, b = 0; j < C; ++b) {
      const bool own = (b >= p0 && b < p0 + k) || (b >= nq + p0 && b < nq + p0 + k);
      if (!own) G.tile_gbit[j++] = b;
    }
    for (int j = 0; j < k; ++j) { G.tile_gbit[C + j] = p0 + j; G.tile_gbit[C + k + j] = nq + p0 + j; }
    // the last two sites of the group sit on the register bits
    G.reg_tb[0] = C + k - 2; G.reg_tb[1] = C + 2 * k - 2; G.reg_tb[2] = C + k - 1; G.reg_tb[3] = C + 2 * k - 1;
    G.n_reg_sites = 2;
    G.reg_q[0] = nq - 1 - (p0 + k - 2);
    G.reg_q[1] = nq - 1 - (p0 + k - 1);
    int n = 0;
    for (int j = 0; j < C; ++j) G.thr_tb[n++] = j;
    for (int j = 0; j < k - 2; ++j) G.thr_tb[n++] = C + j;
    for (int j = 0; j < k - 2; ++j) G.thr_tb[n++] = C + k + j;
    G.n_thr_sites = k - 2;
    for (int j = 0; j < k - 2; ++j) { G.thr_q[j] = nq - 1 - (p0 + j); G.thr_c[j] = C + j; G.thr_r[j] = C + (k - 2) + j; }
    finish_geom(G, nq, false);
    p0 += k;
  }
  return tt;
}

const TileTypes& tile_types(int nq) {
  static TileTypes cache[32];
  static bool have[32] = {};
  if (!have[nq]) { cache[nq] = make_tile_types(nq); have[nq] = true; }
  return cache[nq];
}

bool any_double_flip(const SiteOpsDensity& so, int nq) {
  for (int q = 0; q < nq; ++q)
    for (int p = 0; p < 4; ++p) {
      const cplx z = so.T[q * 16 + p * 4 + (p ^ 3)];
      if (z.re != 0.0 || z.im != 0.0) return true;
    }
  return false;
}

}  // namespace

bool dens_tile_supported(const Geometry& g) {
  return g.kind == PD_DENSITY && g.nq >= 8 && g.nq <= 13 &&
         (g.dim >> DT_BITS) * (size_t)g.batch < ((size_t)1 << 31);
}

// out = F (sum_j w_j in_j); ymat receives the combination unless the input is plain (one input, weight 1).
int launch_dens_stage(const Geometry& g, cplx* out, cplx* ymat, int n_in, const cplx* const* ins, const double* w,
                      const SiteOpsDensity& so, cudaStream_t s) {
  if (n_in > DT_MAXIN) throw Error(PD_ERR_INVALID, "density stage takes at most 8 inputs");
  static bool attr_set[64] = {};
  int dev = 0;
  PD_CUDA_CHECK(cudaGetDevice(&dev));
  const int smem = DT_TILE * 16 + kMaxSitesDensity * 16 * 16;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_dens_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PD_CUDA_CHECK(cudaFuncSetAttribute(k_dens_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dev] = true;
  }
  const TileTypes& tt = tile_types(g.nq);
  const bool plain = n_in == 1 && w[0] == 1.0;
  // PD_DENS_MODE: 1 (default) = one launch per application: six sites inside the tile, the others gathered
  // through L2 (the combination is formed by a separate pass when there is one); 2 = tiles only, 2-3 launches
  static const int mode = [] { const char* e = std::getenv("PD_DENS_MODE"); return e ? std::atoi(e) : 1; }();
  if (!plain && ymat == nullptr) throw Error(PD_ERR_STATE, "density stage needs a buffer for the combined input");
  DensParams P{};
  P.nq = g.nq; P.n_in = n_in; P.need_both = any_double_flip(so, g.nq) ? 1 : 0; P.dim = g.dim;
  P.tiles_per_vec = (unsigned)(g.dim >> DT_BITS);
  for (int j = 0; j < n_in; ++j) { P.v[j] = ins[j]; P.w[j] = w[j]; }
  P.ymat = plain ? nullptr : ymat;
  P.ysrc = plain ? ins[0] : ymat;
  P.out = out;
  P.diag = g.diag;
  const unsigned grid = P.tiles_per_vec * (unsigned)g.batch;
  if (mode == 1) {
    int n = 0;
    if (!plain) {
      n += launch_lincomb(g, ymat, n_in, ins, w, s);
      P.n_in = 1; P.v[0] = ymat; P.w[0] = 1.0; P.ymat = nullptr; P.ysrc = ymat;
    }
    P.gather_rest = 1;
    k_dens_tile<true><<<grid, DT_NT, smem, s>>>(P, tt.g[0], so);
    PD_CUDA_CHECK(cudaGetLastError());
    return n + 1;
  }
  k_dens_tile<true><<<grid, DT_NT, smem, s>>>(P, tt.g[0], so);
  for (int ti = 1; ti < tt.n; ++ti) k_dens_tile<false><<<grid, DT_NT, smem, s>>>(P, tt.g[ti], so);
  PD_CUDA_CHECK(cudaGetLastError());
  return tt.n;
}

}  // namespace pd
