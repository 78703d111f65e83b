// Adjoint-sweep kernel variants of the small-register family (see small_ket.cuh).
#include "small_ket.cuh"
namespace pd {
namespace sk {
void launch_backward(int nq, const SkBwd& P, int nC, cudaStream_t st) {
  if (nq <= 4) launch_units(k_small_backward<4>, P, nC, P.n_units, 2 * P.L, st);
  else if (nq <= 8) launch_units(k_small_backward<8>, P, nC, P.n_units, 2 * P.L, st);
  else if (nq <= 12) launch_units(k_small_backward<12>, P, nC, P.n_units, 2 * P.L, st);
  else launch_units(k_small_backward<16>, P, nC, P.n_units, 2 * P.L, st);
}
}  // namespace sk
}  // namespace pd
