// ensemble_kernel<VecEvaluator<DiasRow | ShinRow>>.
#include "ens_vec.cuh"

// developer knob: resident CTAs per SM of the 128-thread Shin kernel (6 -> 80 registers, 8 -> 64)
#ifndef BISIP_SHIN_MB128
#define BISIP_SHIN_MB128 6
#endif

namespace bisip {

int launch_ens_dias_shin(const EnsembleParams& P, dim3 grid, size_t smem, cudaStream_t st) {
  if (P.d.model == BISIP_MODEL_DIAS) return launch_vec_ensemble<DiasRow, 8>(P, grid, smem, st, "ensemble_dias");
  return launch_vec_ensemble<ShinRow, BISIP_SHIN_MB128>(P, grid, smem, st, "ensemble_shin");
}

}  // namespace bisip
