// Micro-benchmarks that decide how the N >= 24 ket kernels may reuse data:
//   (a) streaming read / read-modify-write bandwidth versus working-set size: is an L2 hit cheaper than
//       an HBM access on B200, i.e. can a second launch over an L2-resident super-tile be faster?
//   (b) cluster tiles: remote shared-memory (DSMEM) reads of a partner CTA's tile while the same CTAs
//       stream from HBM -- do the two paths overlap, and how many SMs does a cluster launch keep busy?
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2bench l2bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e = (x);                                                         \
    if (e != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

// ---- (a) ---------------------------------------------------------------------------------------------
// Every CTA walks the whole working set `reps` times in 64 KiB tiles (tile index strided by the grid), so
// the working set is re-read while it is (or is not) resident in L2.  MODE 0: read only; 1: out += in.
template <int MODE>
__global__ void __launch_bounds__(256, 3)
k_ws(const double2* __restrict__ in, double2* __restrict__ out, size_t n_tiles, int reps, double* sink) {
  const int t = threadIdx.x;
  double acc = 0.0;
  for (int r = 0; r < reps; ++r)
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const double2* p = in + (tile << 12);
#pragma unroll
      for (int q0 = 0; q0 < 16; q0 += 8) {
        double2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldcg(p + t + 256 * (q0 + i));
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y;
        } else {
          double2* o = out + (tile << 12);
          double2 w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = __ldcg(o + t + 256 * (q0 + i));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            w[i].x += v[i].x; w[i].y += v[i].y;
            __stcg(o + t + 256 * (q0 + i), w[i]);
          }
        }
      }
    }
  if (acc == 1.2345e300) *sink = acc;
}

// ---- (a2) -------------------------------------------------------------------------------------------
// Producer/consumer through L2 as in k_stream_ag: phase 0 writes a region (every CTA its own tiles), phase 1
// (a second launch) reads tiles written by ANOTHER CTA (shift) with load policy POL (0 ld.cg, 1 ld.cs evict-first,
// 2 default ld) and rewrites them: out[i] = in[i] + out[i] where `in` was just written and `out` is dirty.
template <int POL>
__device__ __forceinline__ double2 ldp(const double2* p) {
  if (POL == 0) return __ldcg(p);
  if (POL == 1) return __ldcs(p);
  return *p;
}
__global__ void __launch_bounds__(256, 3) k_fill(double2* a, double2* b, size_t n_tiles) {
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      a[(tile << 12) + threadIdx.x + 256 * i] = make_double2(1.0, tile);
      b[(tile << 12) + threadIdx.x + 256 * i] = make_double2(2.0, i);
    }
}
template <int POL>
__global__ void __launch_bounds__(256, 3) k_consume(const double2* __restrict__ in, double2* __restrict__ out, size_t n_tiles,
                                                    size_t shift) {
  const int t = threadIdx.x;
  for (size_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
    const size_t tile = (tl + shift) % n_tiles;
    const double2* p = in + (tile << 12);
    double2* o = out + (tile << 12);
#pragma unroll
    for (int q0 = 0; q0 < 16; q0 += 8) {
      double2 v[8], w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ldp<POL>(p + t + 256 * (q0 + i));
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = ldp<POL>(o + t + 256 * (q0 + i));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        w[i].x += v[i].x; w[i].y += v[i].y;
        o[t + 256 * (q0 + i)] = w[i];
      }
    }
  }
}

// ---- (b) ---------------------------------------------------------------------------------------------
// Cluster of CS CTAs, 64 KiB tile per CTA.  Per iteration a CTA (1) streams one 64 KiB tile from HBM into its
// shared memory (WHAT & 1) and (2) reads `nbits` partner tiles through DSMEM (WHAT & 2).  Work of different
// iterations is not separated by barriers beyond one cluster barrier per iteration, as a real kernel would.
template <int CS>
__global__ void __launch_bounds__(256, 3)
k_cluster(const double2* __restrict__ in, size_t n_tiles, int what, int nbits, double* sink, unsigned* smids) {
  extern __shared__ __align__(16) unsigned char smem[];
  double2* T = reinterpret_cast<double2*>(smem);
  cg::cluster_group cl = cg::this_cluster();
  const int t = threadIdx.x;
  const unsigned rank = cl.block_rank();
  if (t == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smids[blockIdx.x] = smid;
  }
  double acc = 0.0;
  for (int i = t; i < 4096; i += 256) T[i] = make_double2(i, rank);
  cl.sync();
  for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (what & 1) {
      const double2* p = in + (tile << 12);
#pragma unroll
      for (int q0 = 0; q0 < 16; q0 += 8) {
        double2 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldcs(p + t + 256 * (q0 + i));
#pragma unroll
        for (int i = 0; i < 8; ++i) T[t + 256 * (q0 + i)] = v[i];
      }
    }
    cl.sync();
    if (what & 2) {
      for (int b = 0; b < nbits; ++b) {
        const double2* P = cl.map_shared_rank(T, rank ^ (1u << b));
#pragma unroll
        for (int q0 = 0; q0 < 16; q0 += 8) {
          double2 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = P[t + 256 * (q0 + i)];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y;
        }
      }
    }
    cl.sync();
  }
  if (acc == 1.2345e300) *sink = acc;
}

template <int CS>
void run_cluster(const double2* in, size_t n_tiles, double* sink, unsigned* d_smid, cudaEvent_t e0, cudaEvent_t e1) {
  CK(cudaFuncSetAttribute(k_cluster<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  if (CS > 8) CK(cudaFuncSetAttribute(k_cluster<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int nb = 0;
  while ((1 << nb) < CS) ++nb;
  for (int ctas_per_sm = 1; ctas_per_sm <= 3; ctas_per_sm += 2) {
    const unsigned grid = 148u * ctas_per_sm / CS * CS;
    for (int what = 1; what <= 3; ++what) {
      if (CS == 1 && what != 1) continue;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = 65536;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      float best = 1e30f;
      cudaError_t err = cudaSuccess;
      for (int rep = 0; rep < 3 && err == cudaSuccess; ++rep) {
        CK(cudaEventRecord(e0));
        err = cudaLaunchKernelEx(&cfg, k_cluster<CS>, in, n_tiles, what, nb, sink, d_smid);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      if (err != cudaSuccess) {
        printf("{\"bench\":\"cluster\",\"cs\":%d,\"ctas_per_sm\":%d,\"error\":\"%s\"}\n", CS, ctas_per_sm, cudaGetErrorString(err));
        cudaGetLastError();
        continue;
      }
      std::vector<unsigned> h(grid);
      CK(cudaMemcpy(h.data(), d_smid, 4 * grid, cudaMemcpyDeviceToHost));
      std::vector<int> cnt(256, 0);
      for (unsigned s : h) cnt[s & 255]++;
      int used = 0;
      for (int c : cnt) used += c > 0;
      const double hbm = (what & 1) ? (double)n_tiles * 65536 : 0.0;
      const double ds = (what & 2) ? (double)n_tiles * 65536 * nb : 0.0;
      printf("{\"bench\":\"cluster\",\"cs\":%d,\"ctas_per_sm\":%d,\"what\":\"%s\",\"ms\":%.4f,\"hbm_GBs\":%.1f,\"dsmem_GBs\":%.1f,\"sms_used\":%d,\"grid\":%u}\n",
             CS, ctas_per_sm, what == 1 ? "hbm" : what == 2 ? "dsmem" : "both", best, hbm / best / 1e6, ds / best / 1e6, used, grid);
      fflush(stdout);
    }
  }
}

int main(int argc, char**) {
  const size_t dim = (size_t)1 << 26;
  double2 *in, *out;
  double* sink;
  unsigned* d_smid;
  CK(cudaMalloc(&in, dim * 16));
  CK(cudaMalloc(&out, dim * 16));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMalloc(&d_smid, 4 * 4096));
  CK(cudaMemset(in, 0, dim * 16));
  CK(cudaMemset(out, 0, dim * 16));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  // (a)
  const int mibs[] = {8, 16, 32, 48, 64, 96, 128, 256, 1024};
  for (int mode = 0; mode < 2; ++mode)
    for (int mib : mibs) {
      const size_t n_tiles = (size_t)mib * 16;                       // 64 KiB tiles
      const int reps = (int)(4096 / mib) < 1 ? 1 : (int)(4096 / mib);  // ~4 GiB of loads per launch
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) k_ws<0><<<148 * 3, 256>>>(in, out, n_tiles, reps, sink);
        else k_ws<1><<<148 * 3, 256>>>(in, out, n_tiles, reps, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      CK(cudaGetLastError());
      const double bytes = (double)n_tiles * 65536 * reps * (mode == 0 ? 1.0 : 3.0);
      printf("{\"bench\":\"working_set_%s\",\"MiB_per_vector\":%d,\"reps\":%d,\"ms\":%.4f,\"GBs\":%.1f}\n",
             mode == 0 ? "read" : "rmw", mib, reps, best, bytes / best / 1e6);
      fflush(stdout);
    }
  // (a2)
  for (int mib : {16, 32}) {
    const size_t nt = (size_t)mib * 16;
    for (int pol = 0; pol < 3; ++pol)
      for (size_t shift : {(size_t)0, (size_t)37}) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          k_fill<<<148 * 3, 256>>>(in, out, nt);
          CK(cudaEventRecord(e0));
          if (pol == 0) k_consume<0><<<148 * 3, 256>>>(in, out, nt, shift);
          if (pol == 1) k_consume<1><<<148 * 3, 256>>>(in, out, nt, shift);
          if (pol == 2) k_consume<2><<<148 * 3, 256>>>(in, out, nt, shift);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (rep > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("{\"bench\":\"produce_consume\",\"MiB_per_vector\":%d,\"policy\":\"%s\",\"shift\":%zu,\"ms\":%.4f,\"GBs\":%.1f}\n", mib,
               pol == 0 ? "cg" : pol == 1 ? "cs" : "default", shift, best, 3.0 * nt * 65536 / best / 1e6);
        fflush(stdout);
      }
  }
  // (b)
  const size_t n_tiles = dim >> 12;
  run_cluster<1>(in, n_tiles, sink, d_smid, e0, e1);
  run_cluster<2>(in, n_tiles, sink, d_smid, e0, e1);
  if (argc > 1) {
    run_cluster<4>(in, n_tiles, sink, d_smid, e0, e1);
    run_cluster<8>(in, n_tiles, sink, d_smid, e0, e1);
    run_cluster<16>(in, n_tiles, sink, d_smid, e0, e1);
  }
  return 0;
}
