// Adjoint-sweep kernel variants of the small-register family (see small_ket.cuh).
#include "small_ket.cuh"
namespace pd {
namespace sk {
void launch_backward(int nq, const SkBwd& P, int nC, cudaStream_t st) {
  if (nq <= 8) launch_coop(k_small_backward<8>, P, nC, st);
  else if (nq <= 12) launch_coop(k_small_backward<12>, P, nC, st);
  else launch_coop(k_small_backward<16>, P, nC, st);
}
}  // namespace sk
}  // namespace pd
