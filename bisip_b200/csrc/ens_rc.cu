// Clustered ensemble kernels: FP64 DMMA with stage-1 recompute (n_tau > 64) and the mma.sync TF32 / 3xTF32 tiles.
#include "launch.cuh"

namespace bisip {

int launch_ens_rc(const EnsembleParams& P, dim3 g, const RcPlan& plan, cudaStream_t st) {
#define BISIP_RC_LAUNCH(EVAL, NAME)                                                                       \
  return plan.two_per_sm ? launch_cluster(ensemble_kernel<EVAL, 2>, g, plan.cs, plan.smem, st, NAME, &P)  \
                         : launch_cluster(ensemble_kernel<EVAL, 1>, g, plan.cs, plan.smem, st, NAME, &P)
  switch (prec_planes(P.d.precision)) {
    case 1: BISIP_RC_LAUNCH(DecompTF32Evaluator<1>, "ensemble_decomp_tf32");
    case 3: BISIP_RC_LAUNCH(DecompTF32Evaluator<3>, "ensemble_decomp_3xtf32");
    default: BISIP_RC_LAUNCH(DecompRCEvaluator, "ensemble_decomp_rc");
  }
#undef BISIP_RC_LAUNCH
}

}  // namespace bisip
