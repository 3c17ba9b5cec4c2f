// Launch helpers of the warp-private sampler (sampler_wp.cuh): CTA size from the walker count, resident CTAs per SM
// from the register budget (RB = registers per thread the kernel is built for: 64 or 80).
#pragma once
#include "launch.cuh"
#include "sampler_wp.cuh"

namespace bisip {

inline int wp_threads(int W) { return W <= 32 ? 32 : W <= 64 ? 64 : W <= 128 ? 128 : 256; }

// 65536 registers per SM: RB = 48 -> 1280 threads resident, 64 -> 1024, 80 -> 768, 128 -> 512 (at most 32 CTAs)
template <class Eval, int RB>
int launch_wp(const EnsembleParams& P, dim3 grid, size_t smem, cudaStream_t st, const char* name) {
  constexpr int T = RB <= 48 ? 1280 : RB <= 64 ? 1024 : RB <= 80 ? 768 : 512;
  switch (wp_threads(P.W)) {
    case 32: return launch(ensemble_wp_kernel<Eval, (T / 32 > 32 ? 32 : T / 32), 32>, grid, smem, st, name, &P, 32);
    case 64: return launch(ensemble_wp_kernel<Eval, T / 64, 64>, grid, smem, st, name, &P, 64);
    case 128: return launch(ensemble_wp_kernel<Eval, T / 128, 128>, grid, smem, st, name, &P, 128);
    default: return launch(ensemble_wp_kernel<Eval, T / 256, 256>, grid, smem, st, name, &P, 256);
  }
}

}  // namespace bisip
