// ensemble_kernel<DecompEvaluator<KC>>: FP64 two-stage contraction on DMMA tiles, n_tau <= 64 (the default path).
#include "launch.cuh"

namespace bisip {

int launch_ens_dmma(const EnsembleParams& P, dim3 grid, size_t smem, cudaStream_t st) {
  const int KC = ceil_div(P.d.n_tau, 16);
  switch (KC) {
    case 1: return launch(ensemble_kernel<DecompEvaluator<1>, 2>, grid, smem, st, "ensemble_decomp", &P);
    case 2: return launch(ensemble_kernel<DecompEvaluator<2>, 2>, grid, smem, st, "ensemble_decomp", &P);
    case 3: return launch(ensemble_kernel<DecompEvaluator<3>, 2>, grid, smem, st, "ensemble_decomp", &P);
    default: return launch(ensemble_kernel<DecompEvaluator<4>, 2>, grid, smem, st, "ensemble_decomp", &P);
  }
}

}  // namespace bisip
