// Forward-evolution kernel variants of the small-register family (see small_ket.cuh):
// partner arrays sized for registers of up to 8, 12 and 16 qubits.
#include "small_ket.cuh"
namespace pd {
namespace sk {
void launch_forward(int nq, const SkFwd& P, int nC, cudaStream_t st) {
  if (nq <= 4) launch_units(k_small_forward<4>, P, nC, P.n_units, 2 * P.L, st);
  else if (nq <= 8) launch_units(k_small_forward<8>, P, nC, P.n_units, 2 * P.L, st);
  else if (nq <= 12) launch_units(k_small_forward<12>, P, nC, P.n_units, 2 * P.L, st);
  else launch_units(k_small_forward<16>, P, nC, P.n_units, 2 * P.L, st);
}
}  // namespace sk
}  // namespace pd
