// Tiled (shared-memory staged) ket kernels -- placeholder until the tiled family lands.
#include "cuda_backend.cuh"
namespace pd {
bool tiled_ket_supported(const Geometry&) { return false; }
int launch_tiled_stage_ket(const Geometry&, cplx*, cplx*, int, const cplx* const*, const double*,
                           const SiteOps&, cplx*, cudaStream_t) {
  throw Error(PD_ERR_STATE, "tiled ket kernels not built");
}
}  // namespace pd
