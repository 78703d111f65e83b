// Micro-benchmarks that size the tiled ket kernels (DESIGN.md "tile shapes"):
//   (a) strided-tile copy bandwidth versus contiguous piece size (type-B tile access pattern)
//   (b) distributed-shared-memory read bandwidth inside a CTA pair (cross-CTA bit flip)
//   (c) cp.async.bulk (1-D TMA) global->shared throughput versus piece size
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o membench membench.cu
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

constexpr int TB = 12;
constexpr int NT = 256;
constexpr int EPT = (1 << TB) / NT;

// tile `tile` of 2^(TB-C) rows x 2^C columns; rows are the top TB-C bits of the index
__device__ __forceinline__ size_t gidx(int n, int C, size_t tile, int e) {
  if (C >= TB) return (tile << TB) + (size_t)e;
  size_t row = (size_t)(e >> C), col = (size_t)(e & ((1 << C) - 1));
  return (row << (n - (TB - C))) + (tile << C) + col;
}

template <int MODE>  // 0 copy, 1 read-only, 2 read 4 write 2
__global__ void __launch_bounds__(NT, 4)
k_tile(const double2* __restrict__ in, double2* __restrict__ out, int n, int C, double* sink) {
  const size_t tile = blockIdx.x;
  const int t = threadIdx.x;
  double2 v[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) v[i] = __ldg(in + gidx(n, C, tile, t + NT * i));
  if (MODE == 2) {
    const size_t dim = (size_t)1 << n;
#pragma unroll
    for (int j = 1; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        double2 w = __ldg(in + j * dim + gidx(n, C, tile, t + NT * i));
        v[i].x += w.x; v[i].y += w.y;
      }
  }
  if (MODE == 1) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) s += v[i].x + v[i].y;
    if (s == 1.2345e300) *sink = s;
  } else {
#pragma unroll
    for (int i = 0; i < EPT; ++i) out[gidx(n, C, tile, t + NT * i)] = v[i];
    if (MODE == 2) {
      const size_t dim = (size_t)1 << n;
#pragma unroll
      for (int i = 0; i < EPT; ++i) out[dim + gidx(n, C, tile, t + NT * i)] = v[i];
    }
  }
}

// ---- (b) DSMEM ---------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1)
k_dsmem(double* sink, int iters, long long* cycles, int remote) {
  extern __shared__ __align__(16) unsigned char smem[];
  double2* T = reinterpret_cast<double2*>(smem);
  cg::cluster_group cl = cg::this_cluster();
  const int t = threadIdx.x;
  for (int i = t; i < 4096; i += blockDim.x) T[i] = make_double2(i, blockIdx.x);
  cl.sync();
  const double2* P = remote ? cl.map_shared_rank(T, cl.block_rank() ^ 1) : T;
  double2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_double2(0, 0);
  long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double2 v = P[(t + 512 * i + it * 32) & 4095];
      acc[i].x += v.x; acc[i].y += v.y;
    }
  }
  long long c1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  if (s == 1.2345e300) *sink = s;
  cl.sync();
  if (t == 0) cycles[blockIdx.x] = c1 - c0;
}

// ---- (c) 1-D bulk copies ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

// Each CTA streams `tiles_per_cta` tiles of 64 KiB (rows x piece) through a 2-slot ring with one
// producer warp issuing one bulk copy per lane and piece.  Consumers only sum a few words.
__global__ void __launch_bounds__(288, 1)
k_bulk(const double2* __restrict__ in, int n, int C, int tiles_per_cta, double* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[2], empty[2];
  double2* ring = reinterpret_cast<double2*>(smem);
  const int t = threadIdx.x;
  if (t == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int piece = 1 << C;              // amplitudes per piece
  const int rows = (1 << TB) >> C;
  if (t >= 256) {                        // producer warp
    const int lane = t - 256;
    for (int it = 0; it < tiles_per_cta; ++it) {
      const int s = it & 1;
      const size_t tile = (size_t)blockIdx.x * tiles_per_cta + it;
      if (it >= 2) mbar_wait(&empty[s], ((it >> 1) - 1) & 1);
      if (lane == 0) mbar_expect(&full[s], 65536);
      __syncwarp();
      for (int r = lane; r < rows; r += 32)
        bulk_g2s(ring + s * 4096 + r * piece, in + gidx(n, C, tile, r << C), piece * 16, &full[s]);
    }
  } else {
    double acc = 0;
    for (int it = 0; it < tiles_per_cta; ++it) {
      const int s = it & 1;
      mbar_wait(&full[s], (it >> 1) & 1);
      double2 v = ring[s * 4096 + t];
      acc += v.x + v.y;
      __syncwarp();
      if ((t & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    if (acc == 1.2345e300) *sink = acc;
  }
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 26;
  size_t dim = (size_t)1 << n;
  double2 *in, *out;
  double* sink;
  CK(cudaMalloc(&in, dim * 16 * 4));
  CK(cudaMalloc(&out, dim * 16 * 2));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(in, 0, dim * 16 * 4));
  CK(cudaMemset(out, 0, dim * 16 * 2));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  unsigned grid = (unsigned)(dim >> TB);
  int Cs[] = {0, 1, 2, 3, 4, 5, 6, 12};
  for (int mode = 0; mode < 3; ++mode)
    for (int C : Cs) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        if (mode == 0) k_tile<0><<<grid, NT>>>(in, out, n, C, sink);
        if (mode == 1) k_tile<1><<<grid, NT>>>(in, out, n, C, sink);
        if (mode == 2) k_tile<2><<<grid, NT>>>(in, out, n, C, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      double bytes = (mode == 0 ? 2.0 : mode == 1 ? 1.0 : 6.0) * dim * 16;
      printf("{\"bench\":\"tile_%s\",\"n\":%d,\"piece_bytes\":%d,\"ms\":%.4f,\"GBs\":%.1f}\n",
             mode == 0 ? "copy" : mode == 1 ? "read" : "r4w2", n, C >= TB ? 65536 : 16 << C, best, bytes / best / 1e6);
      fflush(stdout);
    }
  // (b)
  {
    long long* cyc;
    CK(cudaMalloc(&cyc, 8 * 296));
    CK(cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int remote = 0; remote < 2; ++remote) {
      int iters = 2000;
      k_dsmem<<<296, 512, 65536>>>(sink, iters, cyc, remote);
      CK(cudaDeviceSynchronize());
      k_dsmem<<<296, 512, 65536>>>(sink, iters, cyc, remote);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(296);
      CK(cudaMemcpy(h.data(), cyc, 8 * 296, cudaMemcpyDeviceToHost));
      double avg = 0;
      for (auto c : h) avg += c;
      avg /= 296;
      double bytes = (double)iters * 8 * 512 * 16;
      printf("{\"bench\":\"dsmem_read\",\"remote\":%d,\"bytes_per_clk_per_sm\":%.2f,\"cycles\":%.0f}\n", remote, bytes / avg, avg);
    }
  }
  // (c)
  {
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    int Cb[] = {0, 1, 2, 3, 5, 8, 12};
    for (int C : Cb) {
      int tiles_per_cta = (int)((dim >> TB) / 148);
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k_bulk<<<148, 288, 131072>>>(in, n, C, tiles_per_cta, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      CK(cudaGetLastError());
      double bytes = (double)tiles_per_cta * 148 * 65536;
      printf("{\"bench\":\"bulk_1d_read\",\"n\":%d,\"piece_bytes\":%d,\"ms\":%.4f,\"GBs\":%.1f}\n", n,
             C >= TB ? 65536 : 16 << C, best, bytes / best / 1e6);
      fflush(stdout);
    }
  }
  return 0;
}
