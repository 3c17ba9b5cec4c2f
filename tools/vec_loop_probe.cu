// Developer probe: what the frequency loop of the Cole-Cole / Dias / Shin evaluators (csrc/models.cuh, vec_row_chi)
// sustains ALONE — no stretch move, no barriers, full occupancy — so that the distance between the ensemble kernels and
// the FP64 pipe can be split into "the loop" and "the sampler around it".  Same row constants, same shared-memory
// records, two lanes per proposal as in the warp-private sampler; every thread re-evaluates rows for `reps` rounds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/vec_loop_probe tools/vec_loop_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../bisip_b200/csrc/models.cuh"

using namespace bisip;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

constexpr int kProbeRows = 64;   // proposals per CTA (128 threads, two lanes per row)

template <class Row, int ILP, int MINB>
__global__ void __launch_bounds__(128, MINB) probe(const double* w, const double* y, const double* yerr, const double* theta,
                                                 int ndim, int N, int n_modes, int reps, double* out) {
  extern __shared__ __align__(16) double smem[];
  VecSmem s;
  double* red = vec_carve(s, smem, N, kProbeRows, Row::kRC);
  vec_init(s, N, w, y, yerr, red);
  if (threadIdx.x < kProbeRows)
    Row::prepare(theta + ((size_t)blockIdx.x * kProbeRows + threadIdx.x) * ndim, n_modes, s.rowc + (size_t)threadIdx.x * Row::kRC);
  __syncthreads();
  const int row = threadIdx.x >> 1, sub = threadIdx.x & 1;
  double tot = 0.0;
  bool ok = true;
  for (int r = 0; r < reps; ++r) {
    Row rr;
    rr.load(s.rowc + (size_t)((row + r) & (kProbeRows - 1)) * Row::kRC, n_modes);
    tot += vec_row_chi<Row, ILP, true>(rr, s.fq + sub * kFq, sub, N, 2, 2 * kFq, ok);
  }
  tot += __shfl_xor_sync(0xffffffffu, tot, 1);
  if (sub == 0) out[(size_t)blockIdx.x * kProbeRows + row] = ok ? tot : -tot;
}

template <class Row, int ILP, int MINB>
static void run(const char* name, int ndim, int n_modes, const double* lo, const double* hi, double fp64_inst_per_freq) {
  const int N = 64, sms = 148, grid = sms * MINB, reps = 400;
  std::vector<double> w(N), y(2 * N), ye(2 * N), th((size_t)grid * kProbeRows * ndim);
  for (int j = 0; j < N; ++j) w[j] = 2 * M_PI * pow(10.0, -2.0 + 6.0 * j / (N - 1));
  for (int c = 0; c < 2 * N; ++c) { y[c] = c < N ? 0.9 : -0.05; ye[c] = 0.01; }
  unsigned long long st = 88172645463325252ull;
  for (auto& v : th) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = (st >> 11) * (1.0 / 9007199254740992.0); }
  for (size_t i = 0; i < th.size(); ++i) th[i] = lo[i % ndim] + (hi[i % ndim] - lo[i % ndim]) * (0.05 + 0.9 * th[i]);
  double *dw, *dy, *de, *dt, *dout;
  CK(cudaMalloc(&dw, N * 8)); CK(cudaMalloc(&dy, 2 * N * 8)); CK(cudaMalloc(&de, 2 * N * 8));
  CK(cudaMalloc(&dt, th.size() * 8)); CK(cudaMalloc(&dout, (size_t)grid * kProbeRows * 8));
  CK(cudaMemcpy(dw, w.data(), N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, y.data(), 2 * N * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(de, ye.data(), 2 * N * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, th.data(), th.size() * 8, cudaMemcpyHostToDevice));
  const size_t smem = (vec_smem_doubles(N, kProbeRows, Row::kRC) + 16) * 8;
  auto k = probe<Row, ILP, MINB>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 1e30;
  for (int it = 0; it < 4; ++it) {
    CK(cudaEventRecord(e0));
    k<<<grid, 128, smem>>>(dw, dy, de, dt, ndim, N, n_modes, reps, dout);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  std::vector<double> o((size_t)grid * kProbeRows);
  CK(cudaMemcpy(o.data(), dout, o.size() * 8, cudaMemcpyDeviceToHost));
  int slow = 0; for (double v : o) slow += v < 0;
  const double evals = (double)grid * kProbeRows * reps;
  const double rate = evals / (best * 1e-3);
  printf("{\"probe\": \"%s\", \"ilp\": %d, \"ctas_per_sm\": %d, \"evals_per_s\": %.4e, \"fp64_inst_per_freq\": %.0f, \"fp64_tinst_per_s_e12\": %.2f, "
         "\"rows_on_slow_path\": %d}\n", name, ILP, MINB, rate, fp64_inst_per_freq, rate * N * fp64_inst_per_freq / 1e12, slow);
  cudaFree(dw); cudaFree(dy); cudaFree(de); cudaFree(dt); cudaFree(dout);
}

int main() {
  // prior boxes of bisip_b200/batch.py (reference models.py:212-216, 266-269, 289-293, 309-314)
  const double dlo[5] = {0.9, 0.0, -15, 0.0, 0.0}, dhi[5] = {1.1, 1.0, 0.0, 25.0, 1.0};
  const double slo[6] = {0.0, 0.0, -15, -7, 0, 0}, shi[6] = {1.0, 1.0, -13, -5, 1, 1};
  const double c1lo[4] = {0.9, 0.0, -15, 0.0}, c1hi[4] = {1.1, 1.0, 5, 1.0};
  const double c2lo[7] = {0.9, 0.0, 0.0, -15, -15, 0.0, 0.0}, c2hi[7] = {1.1, 1.0, 1.0, 5, 5, 1.0, 1.0};
  run<DiasRow, 1, 8>("dias", 5, 1, dlo, dhi, 24);
  run<DiasRow, 2, 8>("dias", 5, 1, dlo, dhi, 24);
  run<DiasRow, 4, 6>("dias", 5, 1, dlo, dhi, 24);
  run<ShinRow, 2, 8>("shin", 6, 1, slo, shi, 48);
  run<ShinRow, 2, 6>("shin", 6, 1, slo, shi, 48);
  run<ColeColeRowT<1>, 2, 8>("colecole1", 4, 1, c1lo, c1hi, 33);
  run<ColeColeRowT<2>, 2, 6>("colecole2", 7, 2, c2lo, c2hi, 59);
  return 0;
}
