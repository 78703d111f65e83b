// Sharded register (axis 2): the flips of the qubits that index the GPU.
//
// With the state sharded by its top g qubits, sigma^x on global qubit q pairs this rank's slice
// with the SAME local index on rank ^ (1 << (g-1-q)).  Each rank keeps its slice in a buffer the
// peers have mapped (NVLink peer memory); this kernel reads the partner slices in place,
//
//   out[i] += shift * psi[i] + sum_k coef_k * peer_k[i],
//
// so the transfer IS the accumulation: no receive buffer, no separate axpy pass, `out` is read
// and written once whatever the number of global qubits.  The loads of one thread (4 local
// indices x up to 4 peers) are all issued before the first use, which is what keeps enough bytes
// in flight for NVLink latency.
#include "cuda_backend.cuh"

namespace pd {

namespace {

constexpr int kPeersPerLaunch = 4;
constexpr int kAmpsPerThread = 4;
constexpr int kAccThreads = 256;

// amplitudes travel in the storage precision of the build (pd_common.hpp amp_t); the sums run in double
#if defined(PD_C64)
using avec = float2;
__device__ __forceinline__ avec make_avec(double re, double im) { return make_float2((float)re, (float)im); }
#else
using avec = double2;
__device__ __forceinline__ avec make_avec(double re, double im) { return make_double2(re, im); }
#endif

struct PeerSet {
  const avec* src[kPeersPerLaunch];
  double2 coef[kPeersPerLaunch];
};

template <int NP>
__global__ void __launch_bounds__(kAccThreads)
k_sharded_accumulate(avec* __restrict__ out, const avec* __restrict__ psi, double shift,
                     PeerSet ps, size_t n_amp) {
  const size_t chunk = (size_t)kAccThreads * kAmpsPerThread;
  for (size_t base = (size_t)blockIdx.x * chunk; base < n_amp; base += (size_t)gridDim.x * chunk) {
    avec r[NP > 0 ? NP : 1][kAmpsPerThread], o[kAmpsPerThread], y[kAmpsPerThread];
#pragma unroll
    for (int k = 0; k < NP; ++k)
#pragma unroll
      for (int j = 0; j < kAmpsPerThread; ++j) {
        size_t i = base + (size_t)j * kAccThreads + threadIdx.x;
        r[k][j] = i < n_amp ? __ldcs(ps.src[k] + i) : make_avec(0.0, 0.0);
      }
#pragma unroll
    for (int j = 0; j < kAmpsPerThread; ++j) {
      size_t i = base + (size_t)j * kAccThreads + threadIdx.x;
      o[j] = i < n_amp ? out[i] : make_avec(0.0, 0.0);
      y[j] = (i < n_amp && psi) ? __ldg(psi + i) : make_avec(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < kAmpsPerThread; ++j) {
      size_t i = base + (size_t)j * kAccThreads + threadIdx.x;
      double re = o[j].x + shift * y[j].x, im = o[j].y + shift * y[j].y;
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        re += ps.coef[k].x * r[k][j].x - ps.coef[k].y * r[k][j].y;
        im += ps.coef[k].x * r[k][j].y + ps.coef[k].y * r[k][j].x;
      }
      if (i < n_amp) out[i] = make_avec(re, im);
    }
  }
}

}  // namespace

int launch_sharded_accumulate(size_t n_amp, amp_t* out, const amp_t* psi, double shift, int n_peers,
                              const amp_t* const* peers, const cplx* coef, cudaStream_t s) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t chunk = (size_t)kAccThreads * kAmpsPerThread;
  size_t want = (n_amp + chunk - 1) / chunk;
  int grid = (int)std::min<size_t>(want, (size_t)sms * 8);
  if (grid < 1) grid = 1;
  int launches = 0, done = 0;
  do {
    int np = std::min(kPeersPerLaunch, n_peers - done);
    PeerSet ps{};
    for (int k = 0; k < np; ++k) {
      ps.src[k] = (const avec*)peers[done + k];
      ps.coef[k] = make_double2(coef[done + k].re, coef[done + k].im);
    }
    avec* o = (avec*)out;
    const avec* y = done == 0 ? (const avec*)psi : nullptr;
    double sh = done == 0 ? shift : 0.0;
    switch (np) {
      case 0: k_sharded_accumulate<0><<<grid, kAccThreads, 0, s>>>(o, y, sh, ps, n_amp); break;
      case 1: k_sharded_accumulate<1><<<grid, kAccThreads, 0, s>>>(o, y, sh, ps, n_amp); break;
      case 2: k_sharded_accumulate<2><<<grid, kAccThreads, 0, s>>>(o, y, sh, ps, n_amp); break;
      case 3: k_sharded_accumulate<3><<<grid, kAccThreads, 0, s>>>(o, y, sh, ps, n_amp); break;
      default: k_sharded_accumulate<4><<<grid, kAccThreads, 0, s>>>(o, y, sh, ps, n_amp); break;
    }
    PD_CUDA_CHECK(cudaGetLastError());
    ++launches;
    done += np;
  } while (done < n_peers);
  return launches;
}

}  // namespace pd
