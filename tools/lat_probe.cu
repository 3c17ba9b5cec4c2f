// Developer probe: dependent-issue latency (cycles per op in a serial chain, one warp) and per-SM-sub-partition
// throughput (8 warps per CTA, one CTA per SM) of the instructions the sampler's serial phases are made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lat_probe tools/lat_probe.cu && tools/lat_probe
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 2048;
template <int OP> __device__ __forceinline__ void body(double& d, float& f, unsigned& u, const double* sm, int lane) {
  if (OP == 0) d = fma(d, 1.0000001, 1e-9);
  if (OP == 1) d = d + 1e-9;
  if (OP == 2) { f = (float)d; d = (double)f + 1e-9; }           // F2F.F32.F64 + F2F.F64.F32 + DADD
  if (OP == 3) f = fmaf(f, 1.0000001f, 1e-9f);
  if (OP == 4) u = u * 0xD2511F53u + 12345u;
  if (OP == 5) u = __umulhi(u, 0xD2511F53u) ^ 0x9E3779B9u;
  if (OP == 6) { d = sm[(u & 31)] + d; u = (unsigned)__double2loint(d) & 31; }   // LDS -> DADD -> address
  if (OP == 7) f = __logf(f) + 3.f;
  if (OP == 8) f = logf(f) + 3.f;
  if (OP == 9) d = log(d) + 3.0;
}
template <int OP> __global__ void k(long long* out, double seed) {
  __shared__ double sm[32];
  if (threadIdx.x < 32) sm[threadIdx.x] = 1e-12 * threadIdx.x;
  __syncthreads();
  double d = seed + threadIdx.x; float f = (float)seed + 2.f; unsigned u = threadIdx.x + 7;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < kIters; ++i) body<OP>(d, f, u, sm, threadIdx.x & 31);
  long long t1 = clock64();
  if (d == 12345.678 || f == 3.14f || u == 99) out[1] = 1;   // keep the chain alive
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
template <int OP> void run(const char* name, long long* dout) {
  long long h;
  float r[2];
  for (int v = 0; v < 2; ++v) {
    k<OP><<<1, v ? 256 : 32>>>(dout, 1.5);
    cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
    r[v] = (float)h / kIters;
  }
  printf("%-34s latency %6.1f cyc/op (1 warp)   8 warps/CTA: %6.1f cyc/iter = %5.2f cyc per warp-op per SMSP\n", name, r[0], r[1], r[1] / 2);
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<0>("DFMA", d); run<1>("DADD", d); run<2>("F2F f64->f32->f64 + DADD", d); run<3>("FFMA", d);
  run<4>("IMAD", d); run<5>("IMAD.HI + LOP", d); run<6>("LDS -> DADD -> addr", d); run<7>("__logf + FADD", d);
  run<8>("logf + FADD", d); run<9>("log (f64) + DADD", d);
  return 0;
}
