// Lanczos kernel variants of the small-register family (see small_ket.cuh).
#include "small_ket.cuh"
namespace pd {
namespace sk {
void launch_lanczos(int nq, const SkLanczos& P, int nC, cudaStream_t st) {
  if (nq <= 8) launch_units(k_small_lanczos<8>, P, nC, 1, 2 * P.dim, st);
  else if (nq <= 12) launch_units(k_small_lanczos<12>, P, nC, 1, 2 * P.dim, st);
  else launch_units(k_small_lanczos<16>, P, nC, 1, 2 * P.dim, st);
}
}  // namespace sk
}  // namespace pd
