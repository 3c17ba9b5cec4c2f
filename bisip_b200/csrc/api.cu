// extern "C" boundary of libbisip_b200.so — see include/bisip_b200.h for the contract.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>

#include "common.cuh"
#include "decomp_eval.cuh"
#include "models.cuh"
#include "sampler.cuh"
#include "stats.cuh"

using namespace bisip;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define BISIP_CUDA(expr)                                                                      \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(BISIP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));        \
  } while (0)

// The kernels must run on the device that owns the caller's buffers, whatever device is current in the calling
// thread (one host thread may drive several GPUs): the entry points switch to the device of their primary pointer
// and restore the previous one on return.  A host pointer / unknown pointer leaves the current device alone.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* p) {
    cudaPointerAttributes at;
    if (p == nullptr || cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return; }
    if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
    if (at.device != prev) switched = (cudaSetDevice(at.device) == cudaSuccess);
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int check_desc(const bisip_model_desc* d) {
  if (!d) return fail(BISIP_ERR_BAD_ARG, "desc is null");
  if (d->n_freq <= 0) return fail(BISIP_ERR_BAD_ARG, "n_freq must be positive");
  switch (d->model) {
    case BISIP_MODEL_COLECOLE:
      if (d->n_modes < 1 || d->n_modes > kMaxModes)
        return fail(BISIP_ERR_UNSUPPORTED, "ColeCole n_modes must be in [1,8]");
      if (d->ndim != 1 + 3 * d->n_modes) return fail(BISIP_ERR_BAD_ARG, "ColeCole ndim != 1+3*n_modes");
      break;
    case BISIP_MODEL_DIAS:
      if (d->ndim != 5) return fail(BISIP_ERR_BAD_ARG, "Dias2000 ndim != 5");
      break;
    case BISIP_MODEL_SHIN:
      if (d->ndim != 6) return fail(BISIP_ERR_BAD_ARG, "Shin2015 ndim != 6");
      break;
    case BISIP_MODEL_DECOMP:
      if (d->n_tau <= 0 || d->n_coef <= 0) return fail(BISIP_ERR_BAD_ARG, "Decomp n_tau/n_coef must be positive");
      if (d->ndim != 1 + d->n_coef) return fail(BISIP_ERR_BAD_ARG, "Decomp ndim != 1+n_coef");
      if (d->n_coef > 8) return fail(BISIP_ERR_UNSUPPORTED, "Decomp poly_deg > 7 not supported");
      if (d->precision < BISIP_PREC_FP64 || d->precision > BISIP_PREC_FP64_COLLAPSED)
        return fail(BISIP_ERR_BAD_ARG, "unknown precision");
      break;
    default:
      return fail(BISIP_ERR_BAD_ARG, "unknown model id");
  }
  return BISIP_OK;
}

int device_smem_optin() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return v;
}

// ------------------------------------------------------------------------------------------
// Batched forward / log-probability kernels (same evaluators as the sampler).
// grid (chunks, B): CTA (c, b) handles theta rows [c*kRows, ...) of spectrum b, striding by gridDim.x.
constexpr int kRows = 128;

// Vector-model ensemble kernels: CTA size and resident CTAs per SM.  Their evaluation phase is short, so
// the serial phases of the stretch move (proposals, accept, split) are about half of a step: many co-resident
// spectra hide them, and for <= 128 walkers 128-thread CTAs (8 or 6 per SM instead of 4 x 256) keep fewer
// lanes idle in those phases.  Values from the sweeps in profiles/r01c_vec_occupancy.md and
// profiles/r01d_vec_kernels.md (W=128, N=64).
#ifdef BISIP_VEC_MINB
constexpr int kMinBVec = BISIP_VEC_MINB;
#else
constexpr int kMinBVec = 4;      // 256-thread CTAs, 64 registers
#endif
constexpr int kVecSmallW = 128;  // walkers up to which the 128-thread variant is used
#ifdef BISIP_COLLAPSED_MINB
constexpr int kMinBCollapsed = BISIP_COLLAPSED_MINB;
#else
constexpr int kMinBCollapsed = 3;   // collapsed decomposition, 256-thread CTAs: 80 registers and no spills; at 4 per SM
                                    // (64 registers) the spill reloads cost more than the fourth CTA hides (-4 %)
#endif

// launch ensemble_kernel<VecEvaluator<Row>> in the shape picked for W walkers; MB128 = CTAs/SM of the
// 128-thread variant (8 -> 64 registers, 6 -> 80 registers)
template <class Row, int MB128>
int launch_vec_ensemble(const EnsembleParams& P, dim3 grid, size_t smem, cudaStream_t st, const char* name);

struct BatchParams {
  bisip_model_desc d;
  int B, n_theta;
  const double* theta;
  const double* w; long long w_stride;
  const double* taus; const double* log_taus; long long tau_stride;
  const double* y; const double* yerr; const double* bounds;
  double* Z; double* lp;
};

template <int KC, bool WANT_Z>
__global__ void __launch_bounds__(kThreads) decomp_batch_kernel(const BatchParams P) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y, ndim = P.d.ndim, N = P.d.n_freq;
  DecompShape sh(P.d.n_freq, P.d.n_tau, P.d.n_coef);
  DecompSmem s;
  double* p = decomp_carve(s, smem, sh, kRows);
  double* prop = p; p += kRows * ndim;
  double* chi = p; p += kRows;
  double* bnd = p; p += 2 * ndim;
  double* red = p;
  if (!WANT_Z) for (int i = threadIdx.x; i < 2 * ndim; i += kThreads) bnd[i] = P.bounds[i];
  decomp_init(s, sh, P.d.c_exp, P.w + (size_t)b * P.w_stride, P.taus + (size_t)b * P.tau_stride,
              P.log_taus + (size_t)b * P.tau_stride * P.d.n_coef,
              WANT_Z ? nullptr : P.y + (size_t)b * 2 * N, WANT_Z ? nullptr : P.yerr + (size_t)b * 2 * N, red);
  for (int r0 = blockIdx.x * kRows; r0 < P.n_theta; r0 += gridDim.x * kRows) {
    const int n = min(kRows, P.n_theta - r0);
    const double* th = P.theta + ((size_t)b * P.n_theta + r0) * ndim;
    for (int i = threadIdx.x; i < kRows * ndim; i += kThreads) prop[i] = (i < n * ndim) ? th[i] : 0.0;
    __syncthreads();
    if (WANT_Z) {
      decomp_eval_Z<KC>(s, sh, prop, ndim, n, P.Z + ((size_t)b * P.n_theta + r0) * 2 * N);
    } else {
      NoSide ns;
      decomp_eval_chi<KC>(s, sh, prop, ndim, n, kRows, chi, ns);
      __syncthreads();
      for (int q = threadIdx.x; q < n; q += kThreads)
        P.lp[(size_t)b * P.n_theta + r0 + q] =
            in_bounds(prop + q * ndim, bnd, ndim) ? -0.5 * (chi[q] + s.llconst) : neg_inf();
    }
    __syncthreads();
  }
}

// Collapsed decomposition (decomp_collapsed.cuh): same grid as decomp_batch_kernel.
template <bool WANT_Z>
__global__ void __launch_bounds__(kThreads) decomp_c_batch_kernel(const BatchParams P) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y, ndim = P.d.ndim, N = P.d.n_freq;
  DecompCShape sh(P.d.n_freq, P.d.n_tau, P.d.n_coef);
  DecompCSmem s;
  double* p = decomp_c_carve(s, smem, sh, kRows);
  double* prop = p; p += kRows * ndim;
  double* chi = p; p += kRows;
  double* bnd = p; p += 2 * ndim;
  double* red = p;
  if (!WANT_Z) for (int i = threadIdx.x; i < 2 * ndim; i += kThreads) bnd[i] = P.bounds[i];
  decomp_c_init(s, sh, P.d.c_exp, P.w + (size_t)b * P.w_stride, P.taus + (size_t)b * P.tau_stride,
                P.log_taus + (size_t)b * P.tau_stride * P.d.n_coef,
                WANT_Z ? nullptr : P.y + (size_t)b * 2 * N, WANT_Z ? nullptr : P.yerr + (size_t)b * 2 * N, red);
  for (int r0 = blockIdx.x * kRows; r0 < P.n_theta; r0 += gridDim.x * kRows) {
    const int n = min(kRows, P.n_theta - r0);
    const double* th = P.theta + ((size_t)b * P.n_theta + r0) * ndim;
    for (int i = threadIdx.x; i < kRows * ndim; i += kThreads) prop[i] = (i < n * ndim) ? th[i] : 0.0;
    __syncthreads();
    if (WANT_Z) {
      decomp_c_eval_Z(s, sh, prop, ndim, n, P.Z + ((size_t)b * P.n_theta + r0) * 2 * N);
    } else {
      decomp_c_eval_chi(s, sh, prop, ndim, n, kRows, chi);
      __syncthreads();
      for (int q = threadIdx.x; q < n; q += kThreads)
        P.lp[(size_t)b * P.n_theta + r0 + q] =
            in_bounds(prop + q * ndim, bnd, ndim) ? -0.5 * (chi[q] + s.llconst) : neg_inf();
    }
    __syncthreads();
  }
}

template <class Row, bool WANT_Z>
__global__ void __launch_bounds__(kThreads) vec_batch_kernel(const BatchParams P) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y, ndim = P.d.ndim, N = P.d.n_freq;
  VecSmem s;
  double* p = vec_carve(s, smem, N, kRows, Row::kRC);
  double* prop = p; p += kRows * ndim;
  double* chi = p; p += kRows;
  double* bnd = p; p += 2 * ndim;
  double* red = p;
  if (!WANT_Z) for (int i = threadIdx.x; i < 2 * ndim; i += kThreads) bnd[i] = P.bounds[i];
  vec_init(s, N, P.w + (size_t)b * P.w_stride, WANT_Z ? nullptr : P.y + (size_t)b * 2 * N,
           WANT_Z ? nullptr : P.yerr + (size_t)b * 2 * N, red);
  for (int r0 = blockIdx.x * kRows; r0 < P.n_theta; r0 += gridDim.x * kRows) {
    const int n = min(kRows, P.n_theta - r0);
    const double* th = P.theta + ((size_t)b * P.n_theta + r0) * ndim;
    for (int i = threadIdx.x; i < kRows * ndim; i += kThreads) prop[i] = (i < n * ndim) ? th[i] : 0.0;
    __syncthreads();
    vec_prepare_rows<Row>(s, P.d.n_modes, prop, ndim, n);
    __syncthreads();
    if (WANT_Z) {
      vec_eval_Z<Row>(s, N, P.d.n_modes, n, P.Z + ((size_t)b * P.n_theta + r0) * 2 * N);
    } else {
      vec_eval_chi<Row>(s, N, P.d.n_modes, n, chi);
      __syncthreads();
      for (int q = threadIdx.x; q < n; q += kThreads)
        P.lp[(size_t)b * P.n_theta + r0 + q] =
            in_bounds(prop + q * ndim, bnd, ndim) ? -0.5 * (chi[q] + s.llconst) : neg_inf();
    }
    __syncthreads();
  }
}

// per-proposal constants of the vector-model kernel that check_desc()/run_batch() will pick
int vec_row_consts(const bisip_model_desc& d) {
  switch (d.model) {
    case BISIP_MODEL_DIAS: return DiasRow::kRC;
    case BISIP_MODEL_SHIN: return ShinRow::kRC;
    default:
      switch (d.n_modes) {
        case 1: return ColeColeRowT<1>::kRC;
        case 2: return ColeColeRowT<2>::kRC;
        case 3: return ColeColeRowT<3>::kRC;
        case 4: return ColeColeRowT<4>::kRC;
        default: return ColeColeRow::kRC;
      }
  }
}

size_t batch_smem_bytes(const bisip_model_desc& d) {
  size_t dbl = (size_t)kRows * d.ndim + kRows + 2 * d.ndim + kWarps;
  if (d.model == BISIP_MODEL_DECOMP)
    dbl += decomp_smem_doubles(DecompShape(d.n_freq, d.n_tau, d.n_coef), kRows);
  else
    dbl += vec_smem_doubles(d.n_freq, kRows, vec_row_consts(d));
  return dbl * 8;
}

template <typename K>
int launch(K kernel, dim3 grid, size_t smem, cudaStream_t st, const char* name, const void* params_ptr,
           int threads = kThreads) {
  if ((int)smem > device_smem_optin())
    return fail(BISIP_ERR_UNSUPPORTED, std::string(name) + ": problem does not fit in shared memory");
  BISIP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {const_cast<void*>(params_ptr)};
  BISIP_CUDA(cudaLaunchKernel((const void*)kernel, grid, dim3(threads), args, smem, st));
  g_launches.fetch_add(1);
  return BISIP_OK;
}

template <class Row, int MB128>
int launch_vec_ensemble(const EnsembleParams& P, dim3 grid, size_t smem, cudaStream_t st, const char* name) {
  if (P.W <= kVecSmallW)
    return launch(ensemble_kernel<VecEvaluator<Row>, MB128, 128>, grid, smem, st, name, &P, 128);
  return launch(ensemble_kernel<VecEvaluator<Row>, kMinBVec, kThreads>, grid, smem, st, name, &P, kThreads);
}

template <typename K>
int launch_cluster(K kernel, dim3 grid, int cluster, size_t smem, cudaStream_t st, const char* name,
                   const void* params_ptr) {
  if ((int)smem > device_smem_optin())
    return fail(BISIP_ERR_UNSUPPORTED, std::string(name) + ": problem does not fit in shared memory");
  BISIP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* args[] = {const_cast<void*>(params_ptr)};
  BISIP_CUDA(cudaLaunchKernelExC(&cfg, (const void*)kernel, args));
  g_launches.fetch_add(1);
  return BISIP_OK;
}

// Large tau grids: pick the cluster size (column split) so that K fits; prefer two CTAs per SM.
struct RcPlan { int cs; bool two_per_sm; size_t smem; };
// 0 = FP64, 1 = TF32, 3 = 3xTF32 (operand planes per product)
int prec_planes(int precision) {
  switch (precision) {
    case BISIP_PREC_TF32: case BISIP_PREC_TF32_MMA: return 1;
    case BISIP_PREC_3XTF32: case BISIP_PREC_3XTF32_MMA: return 3;
    default: return 0;
  }
}
size_t rc_eval_doubles(const bisip_model_desc& d, int rows_pad, int cs) {
  switch (prec_planes(d.precision)) {
    case 1: return DecompTF32Evaluator<1>::smem_doubles(d, rows_pad, cs);
    case 3: return DecompTF32Evaluator<3>::smem_doubles(d, rows_pad, cs);
    default: return DecompRCEvaluator::smem_doubles(d, rows_pad, cs);
  }
}
int plan_rc(const bisip_model_desc& d, size_t other_bytes, int rows_pad, RcPlan* out) {
  // Smallest cluster first; for a given cluster size two CTAs per SM if they fit, else one.  A wider cluster is
  // worse than a lower occupancy: every CTA of a cluster repeats the sampler's serial phases, and measured on
  // B200 (W=256, N=64) a 2-CTA cluster at one CTA/SM runs n_tau=256 at 24.7 TFLOP/s while 4-CTA clusters at two
  // CTAs/SM reach 20.6-22.0 TFLOP/s for n_tau=160..240 (profiles/r01e_cluster_plan.md).
  const size_t two = 113 * 1024, one = (size_t)device_smem_optin();
  const int cands[3] = {1, 2, 4};
  for (int i = 0; i < 3; ++i) {
    const size_t smem = other_bytes + rc_eval_doubles(d, rows_pad, cands[i]) * 8;
    if (smem <= two) { *out = {cands[i], true, smem}; return BISIP_OK; }
    if (smem <= one) { *out = {cands[i], false, smem}; return BISIP_OK; }
  }
  return fail(BISIP_ERR_UNSUPPORTED, "Decomp: n_tau x n_freq too large for a 4-CTA cluster's shared memory");
}
// the clustered ("rc") layout serves every mma.sync reduced-precision run and every FP64 run with n_tau > 64
bool use_rc(const bisip_model_desc& d) {
  return d.model == BISIP_MODEL_DECOMP && d.precision != BISIP_PREC_FP64_COLLAPSED &&
         (d.n_tau > 64 || d.precision != BISIP_PREC_FP64);
}

// tcgen05 path (decomp_umma.cuh): BISIP_PREC_TF32 / _3XTF32 whenever one M = 128 tile holds a half-step
// (rows <= 128), 2N <= 128 columns, and the B planes fit in shared memory next to `other_bytes`.
// n_tau <= 64: 256 tensor-memory columns per CTA, two CTAs per SM — the request is padded so that a third CTA
// (which would spin in tcgen05.alloc) never becomes resident.  n_tau > 64: 512 columns, so the request is padded
// past half an SM's shared memory and exactly one CTA is resident.
struct UmmaPlan { bool ok; bool two_per_sm; size_t smem; bool cluster; };
UmmaPlan plan_umma(const bisip_model_desc& d, size_t other_bytes, int rows, bool allow_cluster = false) {
  UmmaPlan pl{false, false, 0, false};
  const int planes = prec_planes(d.precision);
  if (d.model != BISIP_MODEL_DECOMP || (d.precision != BISIP_PREC_TF32 && d.precision != BISIP_PREC_3XTF32)) return pl;
  if (rows > kUmmaRows || !DecompUmmaShape::fits(d.n_freq, d.n_tau)) return pl;
  const DecompUmmaShape sh(d.n_freq, d.n_tau, d.n_coef);
  size_t smem = other_bytes + decomp_umma_smem_doubles(sh, planes) * 8;
  if (smem > (size_t)device_smem_optin()) {
    // K planes too large for one CTA: a 2-CTA cluster splits the real | imaginary columns (sampler kernel only)
    if (!allow_cluster) return pl;
    const DecompUmmaShape shc(d.n_freq, d.n_tau, d.n_coef, 1, 0);
    smem = other_bytes + decomp_umma_smem_doubles(shc, planes) * 8;
    if (smem > (size_t)device_smem_optin()) return pl;
    pl.ok = true;
    pl.cluster = true;
    pl.smem = smem < 116 * 1024 ? 116 * 1024 : smem;      // 512 tensor-memory columns: one CTA per SM
    return pl;
  }
  pl.ok = true;
  pl.two_per_sm = sh.nchunks == 1 && smem <= 113 * 1024;
  const size_t floor_bytes = pl.two_per_sm ? 77 * 1024 : 116 * 1024;
  pl.smem = smem < floor_bytes ? floor_bytes : smem;
  return pl;
}

// uniform access to the three clustered evaluators for the batch kernel
template <int PREC> struct RcOps;
template <> struct RcOps<0> {
  using Smem = DecompRCSmem;
  static __device__ double* carve(Smem& s, double* b, const DecompRCShape& sh, int rp) { return decomp_rc_carve(s, b, sh, rp); }
  template <class... A> static __device__ void init(A&&... a) { decomp_rc_init(a...); }
  template <class... A> static __device__ void chi(A&&... a) { decomp_rc_eval_chi(a...); }
  template <class... A> static __device__ void Z(A&&... a) { decomp_rc_eval_Z(a...); }
};
template <> struct RcOps<1> {
  using Smem = DecompTF32Smem;
  static __device__ double* carve(Smem& s, double* b, const DecompRCShape& sh, int rp) { return decomp_tf32_carve<1>(s, b, sh, rp); }
  template <class... A> static __device__ void init(A&&... a) { decomp_tf32_init<1>(a...); }
  template <class... A> static __device__ void chi(A&&... a) { decomp_tf32_eval_chi<1>(a...); }
  template <class... A> static __device__ void Z(A&&... a) { decomp_tf32_eval_Z<1>(a...); }
};
template <> struct RcOps<3> {
  using Smem = DecompTF32Smem;
  static __device__ double* carve(Smem& s, double* b, const DecompRCShape& sh, int rp) { return decomp_tf32_carve<3>(s, b, sh, rp); }
  template <class... A> static __device__ void init(A&&... a) { decomp_tf32_init<3>(a...); }
  template <class... A> static __device__ void chi(A&&... a) { decomp_tf32_eval_chi<3>(a...); }
  template <class... A> static __device__ void Z(A&&... a) { decomp_tf32_eval_Z<3>(a...); }
};

// Batched forward / log-probability for large tau grids: grid (chunks*CS, B), cluster (CS,1,1).
template <int PREC, bool WANT_Z>
__global__ void __launch_bounds__(kThreads) decomp_rc_batch_kernel(const BatchParams P) {
  using Ops = RcOps<PREC>;
  extern __shared__ __align__(16) double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int b = blockIdx.y, ndim = P.d.ndim, N = P.d.n_freq;
  const int chunk0 = blockIdx.x / cs, nchunks = gridDim.x / cs;
  DecompRCShape sh(P.d.n_freq, P.d.n_tau, P.d.n_coef, cs, crank);
  typename Ops::Smem s;
  double* p = Ops::carve(s, smem, sh, kRows);
  double* prop = p; p += kRows * ndim;
  double* chi = p; p += kRows;
  double* bnd = p; p += 2 * ndim;
  double* red = p;
  if (!WANT_Z) for (int i = threadIdx.x; i < 2 * ndim; i += kThreads) bnd[i] = P.bounds[i];
  Ops::init(s, sh, P.d.c_exp, P.w + (size_t)b * P.w_stride, P.taus + (size_t)b * P.tau_stride,
                 P.log_taus + (size_t)b * P.tau_stride * P.d.n_coef,
                 WANT_Z ? nullptr : P.y + (size_t)b * 2 * N, WANT_Z ? nullptr : P.yerr + (size_t)b * 2 * N, red);
  for (int r0 = chunk0 * kRows; r0 < P.n_theta; r0 += nchunks * kRows) {
    const int n = min(kRows, P.n_theta - r0);
    const double* th = P.theta + ((size_t)b * P.n_theta + r0) * ndim;
    for (int i = threadIdx.x; i < kRows * ndim; i += kThreads) prop[i] = (i < n * ndim) ? th[i] : 0.0;
    __syncthreads();
    if (WANT_Z) {
      Ops::Z(s, sh, prop, ndim, n, P.Z + ((size_t)b * P.n_theta + r0) * 2 * N);
    } else {
      const int rows_cap = kRows;
      Ops::chi(s, sh, prop, ndim, n, rows_cap, chi);
      __syncthreads();
      if (crank == 0)
        for (int q = threadIdx.x; q < n; q += kThreads)
          P.lp[(size_t)b * P.n_theta + r0 + q] =
              in_bounds(prop + q * ndim, bnd, ndim) ? -0.5 * (chi[q] + s.llconst) : neg_inf();
    }
    __syncthreads();
  }
  if (cs > 1) cluster.sync();
}

// Batched forward / log-probability on the tcgen05 path: grid (chunks, B), 128 theta rows per tile.
template <int PREC, bool WANT_Z>
__global__ void __launch_bounds__(kThreads) decomp_umma_batch_kernel(const BatchParams P) {
  extern __shared__ __align__(16) double smem[];
  const int b = blockIdx.y, ndim = P.d.ndim, N = P.d.n_freq;
  DecompUmmaShape sh(P.d.n_freq, P.d.n_tau, P.d.n_coef);
  DecompUmmaSmem s;
  double* p = decomp_umma_carve<PREC>(s, smem, sh);
  double* prop = p; p += kRows * ndim;
  double* chi = p; p += kRows;
  double* bnd = p; p += 2 * ndim;
  double* red = p;
  if (!WANT_Z) for (int i = threadIdx.x; i < 2 * ndim; i += kThreads) bnd[i] = P.bounds[i];
  decomp_umma_init<PREC>(s, sh, P.d.c_exp, P.w + (size_t)b * P.w_stride, P.taus + (size_t)b * P.tau_stride,
                         P.log_taus + (size_t)b * P.tau_stride * P.d.n_coef,
                         WANT_Z ? nullptr : P.y + (size_t)b * 2 * N, WANT_Z ? nullptr : P.yerr + (size_t)b * 2 * N, red);
  for (int r0 = blockIdx.x * kRows; r0 < P.n_theta; r0 += gridDim.x * kRows) {
    const int n = min(kRows, P.n_theta - r0);
    const double* th = P.theta + ((size_t)b * P.n_theta + r0) * ndim;
    for (int i = threadIdx.x; i < kRows * ndim; i += kThreads) prop[i] = (i < n * ndim) ? th[i] : 0.0;
    __syncthreads();
    decomp_umma_eval<PREC, WANT_Z>(s, sh, prop, ndim, n, chi,
                                   WANT_Z ? P.Z + ((size_t)b * P.n_theta + r0) * 2 * N : nullptr);
    __syncthreads();
    if (!WANT_Z)
      for (int q = threadIdx.x; q < n; q += kThreads)
        P.lp[(size_t)b * P.n_theta + r0 + q] =
            in_bounds(prop + q * ndim, bnd, ndim) ? -0.5 * (chi[q] + s.llconst) : neg_inf();
    __syncthreads();
  }
  decomp_umma_release(s, sh);
}

template <bool WANT_Z>
int run_batch(const BatchParams& P, cudaStream_t st) {
  {
    const size_t other = ((size_t)kRows * P.d.ndim + kRows + 2 * P.d.ndim + kWarps) * 8;
    const UmmaPlan up = plan_umma(P.d, other, kRows);
    if (up.ok) {
      int chunks = ceil_div(P.n_theta, kRows);
      const int cap = max(1, (148 * 2) / max(1, P.B));
      if (chunks > cap) chunks = cap;
      const dim3 g(chunks, P.B);
      if (prec_planes(P.d.precision) == 3)
        return launch(decomp_umma_batch_kernel<3, WANT_Z>, g, up.smem, st, "decomp_umma_3xtf32_batch", &P);
      return launch(decomp_umma_batch_kernel<1, WANT_Z>, g, up.smem, st, "decomp_umma_tf32_batch", &P);
    }
  }
  if (P.d.model == BISIP_MODEL_DECOMP && P.d.precision == BISIP_PREC_FP64_COLLAPSED) {
    const size_t smem = (((size_t)kRows * P.d.ndim + kRows + 2 * P.d.ndim + kWarps) +
                         decomp_c_smem_doubles(DecompCShape(P.d.n_freq, P.d.n_tau, P.d.n_coef), kRows)) * 8;
    int chunks = ceil_div(P.n_theta, kRows);
    const int cap = max(1, (148 * 4) / max(1, P.B));
    if (chunks > cap) chunks = cap;
    return launch(decomp_c_batch_kernel<WANT_Z>, dim3(chunks, P.B), smem, st, "decomp_collapsed_batch", &P);
  }
  if (use_rc(P.d)) {
    RcPlan plan;
    const size_t other = ((size_t)kRows * P.d.ndim + kRows + 2 * P.d.ndim + kWarps) * 8;
    if (int rc = plan_rc(P.d, other, kRows, &plan)) return rc;
    int chunks = ceil_div(P.n_theta, kRows);
    const int cap = max(1, (148 * 2) / max(1, P.B * plan.cs));
    if (chunks > cap) chunks = cap;
    const dim3 g(chunks * plan.cs, P.B);
    switch (prec_planes(P.d.precision)) {
      case 1: return launch_cluster(decomp_rc_batch_kernel<1, WANT_Z>, g, plan.cs, plan.smem, st, "decomp_tf32_batch", &P);
      case 3: return launch_cluster(decomp_rc_batch_kernel<3, WANT_Z>, g, plan.cs, plan.smem, st, "decomp_3xtf32_batch", &P);
      default: return launch_cluster(decomp_rc_batch_kernel<0, WANT_Z>, g, plan.cs, plan.smem, st, "decomp_rc_batch", &P);
    }
  }
  const size_t smem = batch_smem_bytes(P.d);
  int chunks = ceil_div(P.n_theta, kRows);
  const int cap = max(1, (148 * 8) / max(1, P.B));   // enough CTAs to fill the chip, no more
  if (chunks > cap) chunks = cap;
  dim3 grid(chunks, P.B);
  switch (P.d.model) {
    case BISIP_MODEL_COLECOLE:
      switch (P.d.n_modes) {
        case 1: return launch(vec_batch_kernel<ColeColeRowT<1>, WANT_Z>, grid, smem, st, "colecole_batch", &P);
        case 2: return launch(vec_batch_kernel<ColeColeRowT<2>, WANT_Z>, grid, smem, st, "colecole_batch", &P);
        case 3: return launch(vec_batch_kernel<ColeColeRowT<3>, WANT_Z>, grid, smem, st, "colecole_batch", &P);
        case 4: return launch(vec_batch_kernel<ColeColeRowT<4>, WANT_Z>, grid, smem, st, "colecole_batch", &P);
        default: return launch(vec_batch_kernel<ColeColeRow, WANT_Z>, grid, smem, st, "colecole_batch", &P);
      }
    case BISIP_MODEL_DIAS: return launch(vec_batch_kernel<DiasRow, WANT_Z>, grid, smem, st, "dias_batch", &P);
    case BISIP_MODEL_SHIN: return launch(vec_batch_kernel<ShinRow, WANT_Z>, grid, smem, st, "shin_batch", &P);
    default: {
      const int KC = ceil_div(P.d.n_tau, 16);
      switch (KC) {
        case 1: return launch(decomp_batch_kernel<1, WANT_Z>, grid, smem, st, "decomp_batch", &P);
        case 2: return launch(decomp_batch_kernel<2, WANT_Z>, grid, smem, st, "decomp_batch", &P);
        case 3: return launch(decomp_batch_kernel<3, WANT_Z>, grid, smem, st, "decomp_batch", &P);
        default: return launch(decomp_batch_kernel<4, WANT_Z>, grid, smem, st, "decomp_batch", &P);
      }
    }
  }
}

__global__ void build_kernel_matrix(const double* w, int N, const double* taus, int S, double c_exp, double* K) {
  double cs, sn;
  sincospi(0.5 * c_exp, &sn, &cs);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * N; i += gridDim.x * blockDim.x) {
    const int k = i / N, j = i - k * N;
    double kre, kim;
    debye_kernel_term(w[j], taus[k], c_exp, cs, sn, kre, kim);
    K[(size_t)k * 2 * N + j] = kre;
    K[(size_t)k * 2 * N + N + j] = kim;
  }
}

}  // namespace

extern "C" {

int bisip_abi_version(void) { return BISIP_ABI_VERSION; }
const char* bisip_last_error(void) { return g_err.c_str(); }
int64_t bisip_launch_count(void) { return g_launches.load(); }

int bisip_decomp_kernel_kind(const bisip_model_desc* desc, int n_walkers) {
  if (int rc = check_desc(desc)) return rc;
  if (desc->model != BISIP_MODEL_DECOMP || n_walkers < 2) return fail(BISIP_ERR_BAD_ARG, "bisip_decomp_kernel_kind: bad argument");
  const size_t other = sampler_smem_bytes(n_walkers, desc->ndim);
  if (desc->precision == BISIP_PREC_FP64_COLLAPSED) return BISIP_KERNEL_FP64_COLLAPSED;
  {
    const UmmaPlan up = plan_umma(*desc, other, (n_walkers + 1) / 2, true);
    if (up.ok) return up.cluster ? BISIP_KERNEL_TCGEN05_CLUSTER : BISIP_KERNEL_TCGEN05;
  }
  if (!use_rc(*desc)) return BISIP_KERNEL_DMMA;
  RcPlan plan;
  if (int rc = plan_rc(*desc, other, sampler_rows_pad(n_walkers), &plan)) return rc;
  return desc->precision == BISIP_PREC_FP64 ? BISIP_KERNEL_DMMA_CLUSTER : BISIP_KERNEL_MMA_TF32;
}

int bisip_n_keep(int nsteps, int discard, int thin) {
  if (thin < 1 || discard < 0) return 0;
  const int first = discard + thin - 1;
  return nsteps <= first ? 0 : (nsteps - first + thin - 1) / thin;
}

int bisip_forward(const bisip_model_desc* desc, int n_spectra, int n_theta, const double* theta, const double* w,
                  int64_t w_stride, const double* taus, const double* log_taus, int64_t tau_stride, double* Z,
                  void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (n_spectra <= 0 || n_theta <= 0 || !theta || !w || !Z) return fail(BISIP_ERR_BAD_ARG, "bisip_forward: bad argument");
  if (desc->model == BISIP_MODEL_DECOMP && (!taus || !log_taus)) return fail(BISIP_ERR_BAD_ARG, "bisip_forward: taus/log_taus null");
  if (n_spectra > 65535) return fail(BISIP_ERR_UNSUPPORTED, "bisip_forward: n_spectra > 65535 per call");
  DeviceGuard guard(theta);
  BatchParams P{*desc, n_spectra, n_theta, theta, w, w_stride, taus, log_taus, tau_stride, nullptr, nullptr, nullptr, Z, nullptr};
  return run_batch<true>(P, (cudaStream_t)stream);
}

int bisip_log_probability(const bisip_model_desc* desc, int n_spectra, int n_theta, const double* theta,
                          const double* w, int64_t w_stride, const double* taus, const double* log_taus,
                          int64_t tau_stride, const double* y, const double* yerr, const double* bounds,
                          double* lp_out, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (n_spectra <= 0 || n_theta <= 0 || !theta || !w || !y || !yerr || !bounds || !lp_out)
    return fail(BISIP_ERR_BAD_ARG, "bisip_log_probability: bad argument");
  if (desc->model == BISIP_MODEL_DECOMP && (!taus || !log_taus)) return fail(BISIP_ERR_BAD_ARG, "bisip_log_probability: taus/log_taus null");
  if (n_spectra > 65535) return fail(BISIP_ERR_UNSUPPORTED, "bisip_log_probability: n_spectra > 65535 per call");
  DeviceGuard guard(theta);
  BatchParams P{*desc, n_spectra, n_theta, theta, w, w_stride, taus, log_taus, tau_stride, y, yerr, bounds, nullptr, lp_out};
  return run_batch<false>(P, (cudaStream_t)stream);
}

int bisip_decomp_build_kernel(const double* w, int n_freq, const double* taus, int n_tau, double c_exp, double* K,
                              void* stream) {
  if (!w || !taus || !K || n_freq <= 0 || n_tau <= 0) return fail(BISIP_ERR_BAD_ARG, "bisip_decomp_build_kernel: bad argument");
  DeviceGuard guard(K);
  build_kernel_matrix<<<ceil_div(n_freq * n_tau, 256), 256, 0, (cudaStream_t)stream>>>(w, n_freq, taus, n_tau, c_exp, K);
  BISIP_CUDA(cudaGetLastError());
  g_launches.fetch_add(1);
  return BISIP_OK;
}

int bisip_ensemble_run(const bisip_model_desc* desc, int n_spectra, int n_walkers, int nsteps, int step0,
                       uint64_t seed, uint32_t spectrum0, double a, int discard, int thin, const double* w,
                       int64_t w_stride, const double* taus, const double* log_taus, int64_t tau_stride,
                       const double* y, const double* yerr, const double* bounds, double* coords, double* lp,
                       double* chain, double* logp, int32_t* accepted, int32_t* flags, void* stream) {
  if (int rc = check_desc(desc)) return rc;
  if (n_spectra <= 0 || n_walkers < 2 || nsteps < 0 || thin < 1 || discard < 0 || !w || !y || !yerr || !bounds || !coords)
    return fail(BISIP_ERR_BAD_ARG, "bisip_ensemble_run: bad argument");
  if (desc->model == BISIP_MODEL_DECOMP && (!taus || !log_taus)) return fail(BISIP_ERR_BAD_ARG, "bisip_ensemble_run: taus/log_taus null");
  if (!(a > 1.0)) return fail(BISIP_ERR_BAD_ARG, "bisip_ensemble_run: stretch scale a must be > 1");
  DeviceGuard guard(coords);
  cudaStream_t st = (cudaStream_t)stream;
  EnsembleParams P;
  P.d = *desc; P.B = n_spectra; P.W = n_walkers; P.nsteps = nsteps; P.step0 = step0; P.seed = seed;
  P.spectrum0 = spectrum0; P.a = a; P.discard = discard; P.thin = thin;
  P.nkeep = bisip_n_keep(nsteps, discard, thin);
  P.w = w; P.w_stride = w_stride; P.taus = taus; P.log_taus = log_taus; P.tau_stride = tau_stride;
  P.y = y; P.yerr = yerr; P.bounds = bounds; P.coords = coords; P.lp = lp; P.chain = chain; P.logp = logp;
  P.accepted = accepted; P.flags = flags;
  {
    uint32_t km = 1;
    while ((int)km < n_walkers) km <<= 1;
    P.kmask = km - 1;
    int ex;
    P.a_pow2 = (frexp(a, &ex) == 0.5) ? 1 : 0;
    P.inv_a = 1.0 / a;
  }
  {
    // developer knob; default = a quarter of one ensemble step, estimated from the tile count
    const char* e = getenv("BISIP_STAGGER_NS");
    P.stagger_ns = e ? strtoull(e, nullptr, 10) : 6000ull;
  }
  if (flags) BISIP_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * (size_t)n_spectra, st));
  const int rp = sampler_rows_pad(n_walkers);
  size_t smem = sampler_smem_bytes(n_walkers, desc->ndim);
  dim3 grid(n_spectra);
  switch (desc->model) {
    case BISIP_MODEL_COLECOLE:
      smem += vec_smem_doubles(desc->n_freq, rp, vec_row_consts(*desc)) * 8;
      switch (desc->n_modes) {
        case 1: return launch_vec_ensemble<ColeColeRowT<1>, 8>(P, grid, smem, st, "ensemble_colecole");
        case 2: return launch_vec_ensemble<ColeColeRowT<2>, 6>(P, grid, smem, st, "ensemble_colecole");
        case 3: return launch(ensemble_kernel<VecEvaluator<ColeColeRowT<3>>, 2>, grid, smem, st, "ensemble_colecole", &P);
        case 4: return launch(ensemble_kernel<VecEvaluator<ColeColeRowT<4>>, 1>, grid, smem, st, "ensemble_colecole", &P);
        default: return launch(ensemble_kernel<VecEvaluator<ColeColeRow>, 1>, grid, smem, st, "ensemble_colecole", &P);
      }
    case BISIP_MODEL_DIAS:
      smem += VecEvaluator<DiasRow>::smem_doubles(*desc, rp) * 8;
      return launch_vec_ensemble<DiasRow, 8>(P, grid, smem, st, "ensemble_dias");
    case BISIP_MODEL_SHIN:
      smem += VecEvaluator<ShinRow>::smem_doubles(*desc, rp) * 8;
      return launch_vec_ensemble<ShinRow, 6>(P, grid, smem, st, "ensemble_shin");
    default: {
      if (desc->precision == BISIP_PREC_FP64_COLLAPSED) {
        smem += DecompCollapsedEvaluator::smem_doubles(*desc, rp) * 8;
        // measured (profiles/r01h_collapsed_sweep.log): 32 walkers 4.6e9 evals/s at 8 CTAs/SM vs 4.1e9 at 6;
        // 128 walkers 8.1e9 at 6 (80 registers, no spills) vs 6.7e9 at 8
        if (n_walkers <= 64)
          return launch(ensemble_kernel<DecompCollapsedEvaluator, 8, 128>, grid, smem, st, "ensemble_decomp_collapsed", &P, 128);
        if (n_walkers <= kVecSmallW)
          return launch(ensemble_kernel<DecompCollapsedEvaluator, 6, 128>, grid, smem, st, "ensemble_decomp_collapsed", &P, 128);
        return launch(ensemble_kernel<DecompCollapsedEvaluator, kMinBCollapsed, kThreads>, grid, smem, st,
                      "ensemble_decomp_collapsed", &P, kThreads);
      }
      {
        const UmmaPlan up = plan_umma(*desc, smem, (n_walkers + 1) / 2, true);
        if (up.ok && up.cluster) {
          if ((long long)n_spectra * 2 > 2147483647LL) return fail(BISIP_ERR_UNSUPPORTED, "too many spectra per call");
          const dim3 g(n_spectra * 2);
          return prec_planes(desc->precision) == 3
                     ? launch_cluster(ensemble_kernel<DecompUmmaEvaluator<3, true>, 1>, g, 2, up.smem, st, "ensemble_decomp_umma_3xtf32_cluster", &P)
                     : launch_cluster(ensemble_kernel<DecompUmmaEvaluator<1, true>, 1>, g, 2, up.smem, st, "ensemble_decomp_umma_tf32_cluster", &P);
        }
        if (up.ok) {
          const bool x3 = prec_planes(desc->precision) == 3;
          if (up.two_per_sm)
            return x3 ? launch(ensemble_kernel<DecompUmmaEvaluator<3>, 2>, grid, up.smem, st, "ensemble_decomp_umma_3xtf32", &P)
                      : launch(ensemble_kernel<DecompUmmaEvaluator<1>, 2>, grid, up.smem, st, "ensemble_decomp_umma_tf32", &P);
          return x3 ? launch(ensemble_kernel<DecompUmmaEvaluator<3>, 1>, grid, up.smem, st, "ensemble_decomp_umma_3xtf32", &P)
                    : launch(ensemble_kernel<DecompUmmaEvaluator<1>, 1>, grid, up.smem, st, "ensemble_decomp_umma_tf32", &P);
        }
      }
      if (use_rc(*desc)) {
        RcPlan plan;
        if (int rc = plan_rc(*desc, smem, rp, &plan)) return rc;
        if ((long long)n_spectra * plan.cs > 2147483647LL) return fail(BISIP_ERR_UNSUPPORTED, "too many spectra per call");
        dim3 g(n_spectra * plan.cs);
#define BISIP_RC_LAUNCH(EVAL, NAME)                                                                       \
  return plan.two_per_sm ? launch_cluster(ensemble_kernel<EVAL, 2>, g, plan.cs, plan.smem, st, NAME, &P)  \
                         : launch_cluster(ensemble_kernel<EVAL, 1>, g, plan.cs, plan.smem, st, NAME, &P)
        switch (prec_planes(desc->precision)) {
          case 1: BISIP_RC_LAUNCH(DecompTF32Evaluator<1>, "ensemble_decomp_tf32");
          case 3: BISIP_RC_LAUNCH(DecompTF32Evaluator<3>, "ensemble_decomp_3xtf32");
          default: BISIP_RC_LAUNCH(DecompRCEvaluator, "ensemble_decomp_rc");
        }
#undef BISIP_RC_LAUNCH
      }
      smem += DecompEvaluator<4>::smem_doubles(*desc, rp) * 8;
      const int KC = ceil_div(desc->n_tau, 16);
      switch (KC) {
        case 1: return launch(ensemble_kernel<DecompEvaluator<1>, 2>, grid, smem, st, "ensemble_decomp", &P);
        case 2: return launch(ensemble_kernel<DecompEvaluator<2>, 2>, grid, smem, st, "ensemble_decomp", &P);
        case 3: return launch(ensemble_kernel<DecompEvaluator<3>, 2>, grid, smem, st, "ensemble_decomp", &P);
        default: return launch(ensemble_kernel<DecompEvaluator<4>, 2>, grid, smem, st, "ensemble_decomp", &P);
      }
    }
  }
}

int64_t bisip_column_stats_workspace(int n_spectra, int64_t n_samples, int n_cols) {
  if (n_spectra <= 0 || n_samples <= 0 || n_cols <= 0) return 0;
  return (int64_t)n_spectra * n_samples * n_cols * 8;
}

int bisip_column_stats(const double* data, int n_spectra, int64_t n_samples, int n_cols, int n_pct,
                       const int64_t* pct_lo, const double* pct_gamma, double* pct_out, double* mean_out,
                       double* std_out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!data || n_spectra <= 0 || n_samples <= 0 || n_cols <= 0 || n_pct < 0 || !workspace)
    return fail(BISIP_ERR_BAD_ARG, "bisip_column_stats: bad argument");
  if (n_pct > kMaxPct) return fail(BISIP_ERR_UNSUPPORTED, "bisip_column_stats: at most 16 percentiles per call");
  if (n_pct > 0 && (!pct_lo || !pct_gamma || !pct_out)) return fail(BISIP_ERR_BAD_ARG, "bisip_column_stats: percentile arrays null");
  if (workspace_bytes < bisip_column_stats_workspace(n_spectra, n_samples, n_cols))
    return fail(BISIP_ERR_BAD_ARG, "bisip_column_stats: workspace too small");
  if (n_spectra > 65535) return fail(BISIP_ERR_UNSUPPORTED, "bisip_column_stats: n_spectra > 65535 per call");
  DeviceGuard guard(data);
  cudaStream_t st = (cudaStream_t)stream;
  StatsParams P;
  P.data = data; P.keys = (unsigned long long*)workspace; P.n = n_samples; P.ncol = n_cols; P.B = n_spectra;
  P.npct = n_pct;
  for (int i = 0; i < n_pct; ++i) { P.lo[i] = pct_lo[i]; P.gamma[i] = pct_gamma[i]; }
  P.pct_out = pct_out; P.mean_out = mean_out; P.std_out = std_out;
  dim3 g1((unsigned)((n_samples + 4095) / 4096), n_spectra);
  transpose_keys_kernel<<<g1, kThreads, 0, st>>>(P);
  BISIP_CUDA(cudaGetLastError());
  dim3 g2(n_cols, n_spectra);
  column_select_kernel<<<g2, kThreads, 0, st>>>(P);
  BISIP_CUDA(cudaGetLastError());
  g_launches.fetch_add(2);
  return BISIP_OK;
}

}  // extern "C"
