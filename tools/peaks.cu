// Microbenchmarks for the roofline denominators that MEASURED_PEAKS.json lacks:
// FP64 vector (DFMA), FP32 vector (FFMA), FP64 tensor (DMMA via mma.sync .f64, all
// four shapes), plus a fragment-layout self-check for the shapes the decomposition
// kernel uses.  Prints one JSON object.  Build: see Makefile target `peaks`.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, double s) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], s, 1e-9);
  }
  double r = 0; for (int i = 0; i < ILP; ++i) r += acc[i];
  if (r == 123.456) out[0] = r;
}
template <int ILP>
__global__ void k_ffma(float* out, int iters, float s) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], s, 1e-9f);
  }
  float r = 0; for (int i = 0; i < ILP; ++i) r += acc[i];
  if (r == 123.456f) out[0] = r;
}
// SHAPE: 0 m8n8k4, 1 m16n8k4, 2 m16n8k8, 3 m16n8k16
template <int SHAPE, int ILP>
__global__ void k_dmma(double* out, int iters, double s) {
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = s * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = s * (threadIdx.x - i);
  double d[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (SHAPE == 0) { double (&dd)[2] = *reinterpret_cast<double(*)[2]>(&d[i][0]); dmma884(dd, a[0], b[0]); }
      if (SHAPE == 1) { const double (&aa)[2] = *reinterpret_cast<const double(*)[2]>(&a[0]); dmma1684(d[i], aa, b[0]); }
      if (SHAPE == 2) { const double (&aa)[4] = *reinterpret_cast<const double(*)[4]>(&a[0]);
                        const double (&bb)[2] = *reinterpret_cast<const double(*)[2]>(&b[0]); dmma1688(d[i], aa, bb); }
      if (SHAPE == 3) dmma16816(d[i], a, b);
    }
  }
  double r = 0; for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) r += d[i][j];
  if (r == 123.456) out[0] = r;
}

template <class F> double time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

// ---- layout self-check: D = A(16xK) * B(Kx8) with documented fragment maps ----
template <int SHAPE>
__global__ void k_layout(const double* A, const double* B, double* D, int K) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double d[4] = {0, 0, 0, 0};
  if (SHAPE == 0) {  // m8n8k4: A 8xK
    double dd[2] = {0, 0};
    for (int k0 = 0; k0 < K; k0 += 4) dmma884(dd, A[g * K + k0 + t], B[(k0 + t) * 8 + g]);
    D[g * 8 + 2 * t] = dd[0]; D[g * 8 + 2 * t + 1] = dd[1];
    return;
  }
  if (SHAPE == 1) for (int k0 = 0; k0 < K; k0 += 4) {
    double a[2] = {A[g * K + k0 + t], A[(g + 8) * K + k0 + t]};
    dmma1684(d, a, B[(k0 + t) * 8 + g]);
  }
  if (SHAPE == 2) for (int k0 = 0; k0 < K; k0 += 8) {
    double a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = A[(g + 8 * (i & 1)) * K + k0 + t + 4 * (i >> 1)];
    for (int i = 0; i < 2; ++i) b[i] = B[(k0 + t + 4 * i) * 8 + g];
    dmma1688(d, a, b);
  }
  if (SHAPE == 3) for (int k0 = 0; k0 < K; k0 += 16) {
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = A[(g + 8 * (i & 1)) * K + k0 + t + 4 * (i >> 1)];
    for (int i = 0; i < 4; ++i) b[i] = B[(k0 + t + 4 * i) * 8 + g];
    dmma16816(d, a, b);
  }
  D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1];
  D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}

template <int SHAPE> double layout_err() {
  const int K = 16, M = (SHAPE == 0 ? 8 : 16);
  std::vector<double> A(M * K), B(K * 8), D(M * 8, 0.0), R(M * 8, 0.0);
  for (int i = 0; i < M * K; ++i) A[i] = sin(0.37 * i + 0.1);
  for (int i = 0; i < K * 8; ++i) B[i] = cos(0.11 * i - 0.3);
  for (int m = 0; m < M; ++m) for (int n = 0; n < 8; ++n) { double s = 0; for (int k = 0; k < K; ++k) s = fma(A[m * K + k], B[k * 8 + n], s); R[m * 8 + n] = s; }
  double *dA, *dB, *dD; CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&dD, D.size() * 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  k_layout<SHAPE><<<1, 32>>>(dA, dB, dD, K); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost));
  double e = 0; for (int i = 0; i < M * 8; ++i) e = fmax(e, fabs(D[i] - R[i]));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return e;
}


// `peaks --mix`: does an FP64 instruction cost ONE issue slot (and two cycles of its half-rate pipe) or does it hold the
// warp scheduler's dispatch port for both cycles?  8 independent DFMA chains per thread, interleaved with NI independent
// integer instructions per DFMA (KIND 0: IMAD, 1: SHF rotate-by-self, 2: shared-memory load).  If the companion
// instructions hide in the DFMA's second cycle the DFMA rate stays at peak for NI = 1; if the port is held it drops to 2/3.
template <int NI, int KIND>
__global__ void k_mix(double* out, int iters, double s, unsigned m, unsigned c) {
  __shared__ unsigned sh[256];
  sh[threadIdx.x & 255] = c;
  __syncthreads();
  double acc[8];
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = threadIdx.x + i; x[i] = threadIdx.x * 8 + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] = fma(acc[i], s, 1e-9);
#pragma unroll
      for (int k = 0; k < NI; ++k) {
        if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %0, %1;" : "+r"(x[i]) : "r"(c));
        if (KIND == 1) asm volatile("shf.l.wrap.b32 %0, %0, %0, %0;" : "+r"(x[i]));
        if (KIND == 2) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sh[(x[i] + k) & 255]))); x[i] ^= v; }
      }
    }
  }
  double r = 0; unsigned xs = 0;
  for (int i = 0; i < 8; ++i) { r += acc[i]; xs += x[i]; }
  if (r == 123.456 || xs == 0x12345u) out[0] = r + xs;
}


// `peaks --mix` part 2: FP64 rate by operand pattern.  The peak loop above feeds DFMA one register pair (acc) + a uniform
// register + an immediate; real code reads three register pairs per DFMA.  MODE 1: acc = fma(acc, b_i, c_i) with b_i, c_i
// in distinct registers; 2: acc = fma(b_i, c_i, acc); 3: DMUL acc * b_i; 4: DADD acc + b_i; 5: two-chain mix
// d = fma(a, b, c) where a, b, c are all results of other chains (the shape of the model loops).
__constant__ double kDopConst[8] = {1e-9, 2e-9, 3e-9, 4e-9, 5e-9, 6e-9, 7e-9, 8e-9};
template <int MODE>
__global__ void k_dop(double* out, const double* __restrict__ in, int iters) {
  double acc[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = threadIdx.x + i; b[i] = in[threadIdx.x * 16 + i]; c[i] = in[threadIdx.x * 16 + 8 + i]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 1) acc[i] = fma(acc[i], b[i], c[i]);
      if (MODE == 2) acc[i] = fma(b[i], c[i], acc[i]);
      if (MODE == 3) acc[i] = acc[i] * b[i];
      if (MODE == 4) acc[i] = acc[i] + c[i];
      if (MODE == 5) acc[i] = fma(acc[(i + 3) & 7], b[i], acc[(i + 5) & 7]);
      if (MODE == 6) acc[i] = fma(c[i], c[i], acc[i]);          // the same register twice + one more
      if (MODE == 7) acc[i] = fma(acc[i], b[i], acc[i]);
      if (MODE == 8) acc[i] = fma(acc[i], b[i], kDopConst[i]);  // two registers + a constant-bank operand
    }
  }
  double r = 0;
  for (int i = 0; i < 8; ++i) r += acc[i];
  if (r == 123.456) out[0] = r;
}
template <int MODE> static double dop_rate(double* dout, int sms, int iters) {
  const int grid = sms * 8, blk = 256;
  static double* din = nullptr;
  if (!din) {
    std::vector<double> h(256 * 16);
    for (int t = 0; t < 256; ++t) for (int i = 0; i < 16; ++i) h[t * 16 + i] = i < 8 ? 1.0000001 + 1e-12 * (t + i) : 1e-9 * (1 + i + t);
    CK(cudaMalloc(&din, h.size() * 8)); CK(cudaMemcpy(din, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  }
  double ms = time_ms([&] { k_dop<MODE><<<grid, blk>>>(dout, din, iters); });
  return 8.0 * iters * (double)grid * blk / ms / 1e9;      // 1e12 FP64 thread-instructions per second
}

// dependent-issue latency (one warp, one chain, clock64 around 4096 dependent instructions)
template <int KIND>
__global__ void k_lat(double* out, long long* cyc, double s, float sf, unsigned m) {
  double a = threadIdx.x + 1.0; float f = threadIdx.x + 1.0f; unsigned x = threadIdx.x + 1u;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < 4096; ++i) {
    if (KIND == 0) a = fma(a, s, 1e-9);
    if (KIND == 1) f = fmaf(f, sf, 1e-9f);
    if (KIND == 2) asm volatile("mad.lo.u32 %0, %0, %0, %1;" : "+r"(x) : "r"(m));
    if (KIND == 3) a = a * s;
    if (KIND == 4) a = a + s;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  if (a == 123.456 || f == 123.456f || x == 0x12345u) out[0] = a + f + x;
}
template <int KIND> static double lat_cycles(double* dout) {
  static long long* dc = nullptr;
  if (!dc) CK(cudaMalloc(&dc, 8));
  k_lat<KIND><<<1, 32>>>(dout, dc, 1.0000001, 1.0000001f, 3u); CK(cudaDeviceSynchronize());
  k_lat<KIND><<<1, 32>>>(dout, dc, 1.0000001, 1.0000001f, 3u); CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost));
  return h / 4096.0;
}
template <int NI, int KIND> static double mix_rate(double* dout, int sms, int iters) {
  const int grid = sms * 8, blk = 256;
  double ms = time_ms([&] { k_mix<NI, KIND><<<grid, blk>>>(dout, iters, 1.0000001, 3u, 7u); });
  return 2.0 * 8 * iters * (double)grid * blk / ms / 1e9;      // DFMA TFLOP/s
}
static int mix() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* dout; CK(cudaMalloc(&dout, 64));
  const int iters = 10000;
  printf("{\"gpu\": \"%s\", \"what\": \"DFMA TFLOP/s with NI companion instructions per DFMA\"", p.name);
  printf(", \"imad\": [%.2f, %.2f, %.2f, %.2f]", mix_rate<0, 0>(dout, sms, iters), mix_rate<1, 0>(dout, sms, iters), mix_rate<2, 0>(dout, sms, iters), mix_rate<3, 0>(dout, sms, iters));
  printf(", \"shf\": [%.2f, %.2f, %.2f, %.2f]", mix_rate<0, 1>(dout, sms, iters), mix_rate<1, 1>(dout, sms, iters), mix_rate<2, 1>(dout, sms, iters), mix_rate<3, 1>(dout, sms, iters));
  printf(", \"lds\": [%.2f, %.2f, %.2f]", mix_rate<0, 2>(dout, sms, iters), mix_rate<1, 2>(dout, sms, iters), mix_rate<2, 2>(dout, sms, iters));
  printf(", \"fp64_tinst_per_s_e12\": {\"peak_loop\": %.2f, \"fma_acc_r_r\": %.2f, \"fma_r_r_acc\": %.2f, \"mul_acc_r\": %.2f, \"add_acc_r\": %.2f, \"fma_cross_chain\": %.2f, \"fma_r_r_acc_same_r\": %.2f, \"fma_acc_r_acc\": %.2f, \"fma_acc_r_constbank\": %.2f}",
         mix_rate<0, 0>(dout, sms, iters) / 2, dop_rate<1>(dout, sms, iters), dop_rate<2>(dout, sms, iters), dop_rate<3>(dout, sms, iters),
         dop_rate<4>(dout, sms, iters), dop_rate<5>(dout, sms, iters), dop_rate<6>(dout, sms, iters), dop_rate<7>(dout, sms, iters),
         dop_rate<8>(dout, sms, iters));
  printf(", \"dependent_latency_cycles\": {\"dfma\": %.1f, \"dmul\": %.1f, \"dadd\": %.1f, \"ffma\": %.1f, \"imad\": %.1f}",
         lat_cycles<0>(dout), lat_cycles<3>(dout), lat_cycles<4>(dout), lat_cycles<1>(dout), lat_cycles<2>(dout));
  printf("}\n");
  return 0;
}

#include <chrono>
#include <cstring>
// `peaks --sustain SEC`: run the m16n8k16 DMMA kernel back to back for SEC seconds (the regime of a
// kernel timed inside a long step, under the power cap) and report the rate of the last half.
static int sustain(double seconds) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* dout; CK(cudaMalloc(&dout, 64));
  const int iters = 40000, ILP = 4, blk = 128, grid = sms * 4;
  const double flop_per_launch = 2.0 * 16 * 8 * 16 * ILP * (double)iters * grid * (blk / 32);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_dmma<3, ILP><<<grid, blk>>>(dout, iters, 1e-3); CK(cudaDeviceSynchronize());
  auto t0 = std::chrono::steady_clock::now();
  double burst = 0, last = 0; int n = 0;
  std::vector<double> rates;
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 4; ++i) k_dmma<3, ILP><<<grid, blk>>>(dout, iters, 1e-3);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    last = 4 * flop_per_launch / ms / 1e9; rates.push_back(last);
    if (last > burst) burst = last; ++n;
  }
  double s = 0; int m = 0;
  for (size_t i = rates.size() / 2; i < rates.size(); ++i) { s += rates[i]; ++m; }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dmma_tflops_burst\": %.3f, \"dmma_tflops_sustained\": %.3f, \"seconds\": %.1f, \"samples\": %d}\n",
         p.name, sms, burst, m ? s / m : last, seconds, n);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && !strcmp(argv[1], "--sustain")) return sustain(atof(argv[2]));
  if (argc >= 2 && !strcmp(argv[1], "--mix")) return mix();
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* dout; CK(cudaMalloc(&dout, 64)); float* fout = (float*)dout;
  const int iters = 20000;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  printf(", \"layout_err\": {\"m8n8k4\": %.3e, \"m16n8k4\": %.3e, \"m16n8k8\": %.3e, \"m16n8k16\": %.3e}",
         layout_err<0>(), layout_err<1>(), layout_err<2>(), layout_err<3>());
  // vector pipes: 8 CTAs/SM x 256 threads
  {
    const int ILP = 8; int grid = sms * 8, blk = 256;
    double ms = time_ms([&] { k_dfma<ILP><<<grid, blk>>>(dout, iters, 1.0000001); });
    printf(", \"dfma_tflops\": %.3f", 2.0 * ILP * iters * (double)grid * blk / ms / 1e9);
    ms = time_ms([&] { k_ffma<ILP><<<grid, blk>>>(fout, iters, 1.0000001f); });
    printf(", \"ffma_tflops\": %.3f", 2.0 * ILP * iters * (double)grid * blk / ms / 1e9);
  }
  // DMMA: sweep warps/SM
  const char* names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
  const double flop[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
  printf(", \"dmma_tflops\": {");
  for (int sh = 0; sh < 4; ++sh) {
    printf("%s\"%s\": {", sh ? ", " : "", names[sh]);
    int wcfg[4] = {4, 8, 16, 32};
    for (int wi = 0; wi < 4; ++wi) {
      int warps = wcfg[wi]; int blk = 128, grid = sms * (warps * 32 / blk);
      const int ILP = 4;
      double ms = 0;
      if (sh == 0) ms = time_ms([&] { k_dmma<0, ILP><<<grid, blk>>>(dout, iters / 4, 1e-3); });
      if (sh == 1) ms = time_ms([&] { k_dmma<1, ILP><<<grid, blk>>>(dout, iters / 4, 1e-3); });
      if (sh == 2) ms = time_ms([&] { k_dmma<2, ILP><<<grid, blk>>>(dout, iters / 4, 1e-3); });
      if (sh == 3) ms = time_ms([&] { k_dmma<3, ILP><<<grid, blk>>>(dout, iters / 4, 1e-3); });
      double tf = flop[sh] * ILP * (iters / 4) * (double)grid * (blk / 32) / ms / 1e9;
      printf("%s\"w%d\": %.3f", wi ? ", " : "", warps, tf);
    }
    printf("}");
  }
  printf("}}\n");
  return 0;
}
