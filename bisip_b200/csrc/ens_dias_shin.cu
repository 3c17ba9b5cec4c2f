// ensemble_kernel<VecEvaluator<DiasRow | ShinRow>>.
#include "ens_vec.cuh"

namespace bisip {

int launch_ens_dias_shin(const EnsembleParams& P, dim3 grid, size_t smem, cudaStream_t st) {
  if (P.d.model == BISIP_MODEL_DIAS) return launch_vec_ensemble<DiasRow, 8>(P, grid, smem, st, "ensemble_dias");
  return launch_vec_ensemble<ShinRow, 6>(P, grid, smem, st, "ensemble_shin");
}

}  // namespace bisip
