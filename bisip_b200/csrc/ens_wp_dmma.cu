// ensemble_wp_kernel<DecompMmaWarpEvaluator<KC>>: the default FP64 two-stage DMMA contraction under the warp-private
// sampler, 33..64 taus (KC = 3, 4), W <= 256.
#include "ens_wp.cuh"

namespace bisip {

int launch_ens_wp_dmma(const EnsembleParams& P, dim3 grid, cudaStream_t st) {
  const size_t other = wp_smem_bytes(P.W, P.d.ndim);
  if (ceil_div(P.d.n_tau, 16) == 3)
    return launch_wp<DecompMmaWarpEvaluator<3>, 128>(P, grid, other + DecompMmaWarpEvaluator<3>::smem_doubles(P.d) * 8, st, "ensemble_wp_decomp");
  return launch_wp<DecompMmaWarpEvaluator<4>, 128>(P, grid, other + DecompMmaWarpEvaluator<4>::smem_doubles(P.d) * 8, st, "ensemble_wp_decomp");
}

}  // namespace bisip
