// tcgen05 ensemble kernels (TF32 / 3xTF32, operands and accumulators in tensor memory).
#include "launch.cuh"

namespace bisip {

int launch_ens_umma(const EnsembleParams& P, dim3 grid, const UmmaPlan& up, cudaStream_t st) {
  const bool x3 = prec_planes(P.d.precision) == 3;
  if (up.cluster)
    return x3 ? launch_cluster(ensemble_kernel<DecompUmmaEvaluator<3, true>, 1>, grid, 2, up.smem, st, "ensemble_decomp_umma_3xtf32_cluster", &P)
              : launch_cluster(ensemble_kernel<DecompUmmaEvaluator<1, true>, 1>, grid, 2, up.smem, st, "ensemble_decomp_umma_tf32_cluster", &P);
  if (up.two_per_sm)
    return x3 ? launch(ensemble_kernel<DecompUmmaEvaluator<3>, 2>, grid, up.smem, st, "ensemble_decomp_umma_3xtf32", &P)
              : launch(ensemble_kernel<DecompUmmaEvaluator<1>, 2>, grid, up.smem, st, "ensemble_decomp_umma_tf32", &P);
  return x3 ? launch(ensemble_kernel<DecompUmmaEvaluator<3>, 1>, grid, up.smem, st, "ensemble_decomp_umma_3xtf32", &P)
            : launch(ensemble_kernel<DecompUmmaEvaluator<1>, 1>, grid, up.smem, st, "ensemble_decomp_umma_tf32", &P);
}

}  // namespace bisip
