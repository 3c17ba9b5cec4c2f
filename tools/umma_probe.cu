// Developer probe: pins down the tcgen05 operand conventions used by csrc/decomp_umma.cuh on real hardware
// before the sampler depends on them.  One CTA computes D[128][128] = A[128][64] * B[128][64]^T with
// tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8 per instruction, 8 k-steps), accumulators in TMEM:
//   variant 0: A from shared memory (K-major, no swizzle), LBO = K-direction core-matrix stride, SBO = M/N-direction
//   variant 1: same with LBO/SBO swapped (must FAIL if variant 0 is the right reading of the descriptor)
//   variant 2: A from tensor memory (tcgen05.st 32x32b: lane = row, column = k), B as in variant 0
// and prints the max abs error against a host GEMM for each.  Inputs are small integers (exact in TF32).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu && tools/umma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 128, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, no-swizzle operand tile: core matrix = 8 rows x 16 bytes (4 tf32), stored as 128 contiguous
// bytes; core matrices adjacent in K are `kstr` bytes apart, adjacent in M/N `mnstr` bytes apart
__host__ __device__ inline int tile_off(int r, int k, int kstr, int mnstr) {
  return (r & 7) * 16 + (r >> 3) * mnstr + (k >> 2) * kstr + (k & 3) * 4;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo, int sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                    // 128 x 64 x 4 = 32 KB
  uint8_t* sB = smem + 32768;            // 32 KB
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kstr = 128, mnstr = 128 * (K / 4);   // K-adjacent core matrices contiguous
  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<float*>(sA + tile_off(r, k, kstr, mnstr)) = A[i];
    *reinterpret_cast<float*>(sB + tile_off(r, k, kstr, mnstr)) = B[i];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;");      // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tmem_base_s;
  const uint32_t tD = tb, tA = tb + 128;
  if (variant == 2) {   // A -> TMEM: thread = row (lane 32*warp + lane), 64 columns
    const uint32_t taddr = tA + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < K; c += 8) {
      uint32_t v[8];
      for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(A[(size_t)tid * K + c + e]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + c), "r"(v[0]),
                   "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  if (tid == 0) {
    const int lbo = variant == 1 ? mnstr : kstr, sbo = variant == 1 ? kstr : mnstr;
    for (int j = 0; j < K / 8; ++j) {
      const uint64_t db = make_desc(smem_u32(sB) + 2 * kstr * j, lbo, sbo);
      const uint32_t acc = j > 0 ? 1u : 0u;
      if (variant == 2) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tD), "r"(tA + 8 * j), "l"(db),
                     "r"(kIdesc), "r"(acc));
      } else {
        const uint64_t da = make_desc(smem_u32(sA) + 2 * kstr * j, lbo, sbo);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tD), "l"(da), "l"(db),
                     "r"(kIdesc), "r"(acc));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  {
    uint32_t done = 0;
    const uint32_t baddr = smem_u32(&bar);
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(baddr), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tD + ((uint32_t)(32 * warp) << 16);
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int e = 0; e < 8; ++e) D[(size_t)tid * N + c + e] = __uint_as_float(v[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256));
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N);
  srand(1);
  for (auto& v : A) v = (float)(rand() % 17 - 8);
  for (auto& v : B) v = (float)(rand() % 13 - 6);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      R[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  int ok_mask = 0;
  for (int v = 0; v < 3; ++v) {
    cudaMemset(dD, 0xff, D.size() * 4);
    probe<<<1, 128, 65536>>>(dA, dB, dD, v);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", v, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    int bad = 0;
    for (int i = 0; i < M * N; ++i) {
      const double d = fabs((double)D[i] - R[i]);
      if (!(d <= 1e-3)) ++bad;
      if (d > err || d != d) err = d;
    }
    printf("variant %d: max |err| %.3g, mismatches %d / %d  (D[0][0..3] = %g %g %g %g ; ref %g %g %g %g)\n", v, err, bad,
           M * N, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
    if (bad == 0) ok_mask |= 1 << v;
  }
  printf("ok_mask %d\n", ok_mask);
  return 0;
}
