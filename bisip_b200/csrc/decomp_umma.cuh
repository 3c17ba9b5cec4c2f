// Reduced-precision decomposition on the 5th-generation tensor cores (tcgen05 / TMEM), the Blackwell-native
// form of the "TF32 / 3xTF32 variant" of north_star (BASELINE config 4 compares it with FP64 DMMA).
//
//   stage 1  M[row][k] = R0 * sum_i a[row][i] L[i][k]   FP64 on the vector pipe (it is 4 % of the flops and
//            cancels heavily), one thread per (proposal row, 8-tau group); the result is split into TF32
//            hi / lo planes and written straight into TENSOR MEMORY (tcgen05.st, lane = row, column = tau):
//            the chargeability never touches shared memory;
//   stage 2  D[128 rows][NC columns] = M x (K / sigma)   tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = NC,
//            K = 8 per instruction, A from tensor memory, B = K/sigma from shared memory (canonical K-major
//            no-swizzle core-matrix layout, built once per spectrum as hi / lo planes), FP32 accumulators
//            in tensor memory; PREC = 3 issues A_lo B_hi + A_hi B_lo + A_hi B_hi ("3xTF32");
//            one elected thread issues all instructions of a half-step and tcgen05.commit arrives on an mbarrier;
//   epilogue tcgen05.ld (lane = row: one thread owns a row, no shuffles), residual and chi^2 in FP64.
//
// The whole half-step of <= 128 proposals of one spectrum is ONE M = 128 tile.  Tensor memory per CTA:
// 128 columns of accumulators + 64 (hi) + 64 (lo) columns of A = 256 columns, so two CTAs share an SM's 512.
// n_tau > 64 runs in 64-tau chunks with a second A buffer (512 columns, one CTA per SM; api.cu guarantees it).
// The operand conventions (descriptor fields, A-in-TMEM layout) are pinned on hardware by tools/umma_probe.cu.
// Restates reference Decomp_cyth (cython_funcs.pyx:75-94) + _log_likelihood (models.py:59-62) at TF32 precision.
#pragma once
#include "decomp_eval.cuh"

namespace bisip {

static_assert(kThreads == 256, "decomp_umma.cuh maps 256 threads onto 128 TMEM lanes x 2 halves");
#ifdef BISIP_PHASE_TIMING
__device__ long long g_umma_ph[8];
#define UMMA_T0 long long ut_ = clock64();
#define UMMA_MARK(i) { if (blockIdx.x == 0 && threadIdx.x == 0) { long long n_ = clock64(); g_umma_ph[i] += n_ - ut_; ut_ = n_; } }
#else
#define UMMA_T0
#define UMMA_MARK(i)
#endif
constexpr int kUmmaRows = 128;      // MMA M: proposals per tile (>= rows of a half-step: W <= 256)
constexpr int kUmmaChunk = 64;      // taus per A buffer

struct DecompUmmaShape {
  int N, S, D;
  int NCH;     // columns per part (real | imag), multiple of 16
  int NC;      // MMA N = 2 NCH, multiple of 32, <= 128
  int SP;      // taus padded to a multiple of 8 (K steps)
  int nchunks; // A chunks of <= 64 taus
  __host__ __device__ DecompUmmaShape(int n, int s, int d)
      : N(n), S(s), D(d), NCH(ceil_div(n, 16) * 16), NC(2 * ceil_div(n, 16) * 16), SP(ceil_div(s, 8) * 8),
        nchunks(ceil_div(ceil_div(s, 8) * 8, kUmmaChunk)) {}
  __host__ __device__ size_t plane_bytes() const { return (size_t)NC * SP * 4; }
  __host__ __device__ int tmem_cols() const { return nchunks > 1 ? 512 : 256; }
  __host__ __device__ static bool fits(int n_freq, int n_tau) { return n_freq <= 64 && n_tau <= 512; }
};

struct DecompUmmaSmem {
  uint8_t* Bhi;       // [NC x SP] TF32, canonical K-major core matrices (8 columns x 16 bytes)
  uint8_t* Blo;       // PREC == 3
  double* Lk;         // [SP][8]   powers of log_tau per tau (zero padded)
  float4* col;        // [NC]      (y/sigma, delta/sigma) per column as two-float pairs {ys_hi, ys_lo, ds_hi, ds_lo}
  double* part;       // [128]     partial chi^2 of the imaginary half
  uint64_t* bar;      // [2] mbarriers the MMA completions arrive on (one per A buffer)
  uint32_t* tmem;     // tensor-memory base address
  double llconst;
  uint32_t tbase;
  uint32_t phase;     // bit i: parity of the next completion of bar[i]
};

__host__ __device__ inline size_t decomp_umma_smem_doubles(const DecompUmmaShape& sh, int prec) {
  const size_t planes = prec == 3 ? 2 : 1;
  return 16 + planes * sh.plane_bytes() / 8 + (size_t)sh.SP * 8 + 2 * (size_t)sh.NC + kUmmaRows + 4;
}

template <int PREC>
__device__ inline double* decomp_umma_carve(DecompUmmaSmem& s, double* base, const DecompUmmaShape& sh) {
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 127) & ~uintptr_t(127));
  s.Bhi = p; p += sh.plane_bytes();
  s.Blo = p; if (PREC == 3) p += sh.plane_bytes();
  s.Lk = reinterpret_cast<double*>(p); p += (size_t)sh.SP * 64;
  s.col = reinterpret_cast<float4*>(p); p += (size_t)sh.NC * 16;
  s.part = reinterpret_cast<double*>(p); p += kUmmaRows * 8;
  s.bar = reinterpret_cast<uint64_t*>(p); p += 16;
  s.tmem = reinterpret_cast<uint32_t*>(p); p += 8;
  s.phase = 0;
  return base + decomp_umma_smem_doubles(sh, PREC);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// B[n][k] inside a plane: core matrix = 8 columns n x 4 taus k (128 contiguous bytes); K-adjacent core matrices
// contiguous (LBO = 128 B), n-adjacent ones SP/4 core matrices apart (SBO = 32 SP bytes)
__device__ __forceinline__ int umma_b_off(int n, int k, int SP) {
  return (n & 7) * 16 + (n >> 3) * (32 * SP) + (k >> 2) * 128 + (k & 3) * 4;
}
__device__ __forceinline__ uint64_t umma_b_desc(uint32_t saddr, int SP) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(((32 * SP) >> 4) & 0x3fff) << 32) |
         ((uint64_t)1 << 46);     // version 1, no swizzle, base offset 0
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = NC
__device__ __forceinline__ uint32_t umma_idesc(int NC) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(kUmmaRows >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tD, uint32_t tA, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tD), "r"(tA), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (warp-uniform call site): lets the compiler keep the MMA operands in uniform registers
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint32_t baddr, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(baddr), "r"(parity) : "memory");
  }
}

// TF32 hi / lo split of a double: hi = tf32(x), lo = tf32(x - hi) (22 mantissa bits in total)
__device__ __forceinline__ void split_tf32_umma(double x, uint32_t& hi, uint32_t& lo) {
  const float xf = (float)x;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(xf));
  const float r = xf - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

// Per-spectrum constants; allocates tensor memory.  All threads; ends with __syncthreads().
template <int PREC>
__device__ inline void decomp_umma_init(DecompUmmaSmem& s, const DecompUmmaShape& sh, double c_exp,
                                        const double* __restrict__ w, const double* __restrict__ taus,
                                        const double* __restrict__ log_taus, const double* __restrict__ y,
                                        const double* __restrict__ yerr, double* red) {
  const int tid = threadIdx.x;
  const int N = sh.N, S = sh.S, SP = sh.SP, NCH = sh.NCH;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s.tmem)), "r"(sh.tmem_cols())
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s.bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s.bar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(s.Bhi);
    const int nw = (int)(sh.plane_bytes() / 4) * (PREC == 3 ? 2 : 1);
    for (int i = tid; i < nw; i += kThreads) z[i] = 0u;
  }
  for (int i = tid; i < SP * 8; i += kThreads) {
    const int k = i >> 3, p = i & 7;
    s.Lk[i] = (p < sh.D && k < S) ? log_taus[(size_t)p * S + k] : 0.0;
  }
  const bool scaled = (y != nullptr);
  double csum = 0.0;
  for (int n = tid; n < sh.NC; n += kThreads) {
    const int part = n >= NCH, j = n - part * NCH;
    double ys = 0.0, ds = 0.0;
    if (j < N && scaled) {
      const double e = yerr[part * N + j];
      const double is = 1.0 / e;
      ys = y[part * N + j] * is;
      ds = part ? 0.0 : is;
      csum += 2.0 * log(e * e);
    }
    const float ysh = (float)ys, dsh = (float)ds;
    s.col[n] = make_float4(ysh, (float)(ys - (double)ysh), dsh, (float)(ds - (double)dsh));
  }
  __syncthreads();
  double cs, sn;
  sincospi(0.5 * c_exp, &sn, &cs);
  for (int i = tid; i < S * N; i += kThreads) {
    const int k = i / N, j = i - k * N;
    double kre, kim;
    debye_kernel_term(w[j], taus[k], c_exp, cs, sn, kre, kim);
    if (scaled) {
      kre *= 1.0 / yerr[j];
      kim *= 1.0 / yerr[N + j];
    }
    uint32_t hi, lo;
    split_tf32_umma(kre, hi, lo);
    *reinterpret_cast<uint32_t*>(s.Bhi + umma_b_off(j, k, SP)) = hi;
    if (PREC == 3) *reinterpret_cast<uint32_t*>(s.Blo + umma_b_off(j, k, SP)) = lo;
    split_tf32_umma(kim, hi, lo);
    *reinterpret_cast<uint32_t*>(s.Bhi + umma_b_off(NCH + j, k, SP)) = hi;
    if (PREC == 3) *reinterpret_cast<uint32_t*>(s.Blo + umma_b_off(NCH + j, k, SP)) = lo;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
  if ((tid & 31) == 0) red[tid >> 5] = csum;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // B planes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  double tot = 0.0;
  for (int i = 0; i < kWarps; ++i) tot += red[i];
  s.llconst = tot;
  s.tbase = *s.tmem;
  __syncthreads();
}

// All TMEM traffic of this CTA is complete (callers end their last evaluation with a barrier).
__device__ inline void decomp_umma_release(DecompUmmaSmem& s, const DecompUmmaShape& sh) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s.tbase), "r"(sh.tmem_cols()) : "memory");
}

__device__ __forceinline__ void lds_f64x2(uint32_t saddr, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(saddr));
}
__device__ __forceinline__ double lds_f64(uint32_t saddr) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(saddr));
  return x;
}

// Stage 1 for this thread's row over the 8-tau groups g = g0 + h, g0 + h + 2, ... < g1 of one A buffer:
// M = sum_i a_i L[i][tau] in FP64 (ascending powers, like the reference), split into TF32 planes and stored to tensor
// memory with one tcgen05.st per plane and group.  hi = round-to-nearest TF32 (magnitude + half ulp, masked — cvt.rna
// without its inf/nan handling, |M| is O(1)); lo = the FP32 remainder truncated to TF32.  Lk is read with explicit
// ld.shared (a pointer kept in a struct degrades to generic loads).
template <int ND, int PREC>
__device__ __forceinline__ void umma_stage1_part(uint32_t lk_saddr, const double (&a)[8], uint32_t ta, int g0, int g1, int h) {
  for (int g = g0 + h; g < g1; g += 2) {
    const uint32_t la = lk_saddr + (uint32_t)g * 512u;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      double L[8];
#pragma unroll
      for (int i = 0; i + 1 < ND; i += 2) lds_f64x2(la + e * 64 + i * 8, L[i], L[i + 1]);
      if (ND & 1) L[ND - 1] = lds_f64(la + e * 64 + (ND - 1) * 8);
      double m = 0.0;
#pragma unroll
      for (int i = 0; i < ND; ++i) m = fma(a[i], L[i], m);
      const float xf = (float)m;
      hi[e] = (__float_as_uint(xf) + 0x1000u) & 0xffffe000u;
      lo[e] = (__float_as_uint(xf - __uint_as_float(hi[e])) + 0x1000u) & 0xffffe000u;
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta + 8 * g), "r"(hi[0]),
                 "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7])
                 : "memory");
    if (PREC == 3)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta + 64 + 8 * g), "r"(lo[0]),
                   "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7])
                   : "memory");
  }
}

template <int PREC>
__device__ __forceinline__ void umma_stage1_dispatch(int ND, uint32_t lk, const double (&a)[8], uint32_t ta, int g0, int g1, int h) {
  switch (ND) {
    case 1: umma_stage1_part<1, PREC>(lk, a, ta, g0, g1, h); break;
    case 2: umma_stage1_part<2, PREC>(lk, a, ta, g0, g1, h); break;
    case 3: umma_stage1_part<3, PREC>(lk, a, ta, g0, g1, h); break;
    case 4: umma_stage1_part<4, PREC>(lk, a, ta, g0, g1, h); break;
    case 5: umma_stage1_part<5, PREC>(lk, a, ta, g0, g1, h); break;
    case 6: umma_stage1_part<6, PREC>(lk, a, ta, g0, g1, h); break;
    case 7: umma_stage1_part<7, PREC>(lk, a, ta, g0, g1, h); break;
    default: umma_stage1_part<8, PREC>(lk, a, ta, g0, g1, h); break;
  }
}

// One elected thread: the MMAs of K steps [j0, j1) of one A buffer.  Descriptors advance by 256 bytes (16 units) per
// K step, A by 8 tensor-memory columns.  Small terms first: A_lo B_hi, A_hi B_lo, then A_hi B_hi.
template <int PREC>
__device__ __forceinline__ void umma_issue(uint32_t tD, uint32_t tA, uint64_t dhi, uint64_t dlo, uint32_t idesc, int j0, int j1,
                                           bool overwrite) {
  uint32_t acc = overwrite ? 0u : 1u;
  if (PREC == 3) {
    for (int j = j0; j < j1; ++j) { umma_tf32_ts(tD, tA + 64 + 8 * j, dhi + 16u * j, idesc, acc); acc = 1u; }
    for (int j = j0; j < j1; ++j) umma_tf32_ts(tD, tA + 8 * j, dlo + 16u * j, idesc, 1u);
  }
  for (int j = j0; j < j1; ++j) { umma_tf32_ts(tD, tA + 8 * j, dhi + 16u * j, idesc, acc); acc = 1u; }
}

// Evaluate nrows <= 128 proposals.  WANT_Z = false: chi[q] = sum_c ((y_c - Z_c)/sigma_c)^2 (caller barriers
// before reading); WANT_Z = true: Zout[q][2][N] (forward only; init was called with y == nullptr).
// 256 threads: thread = (row r = tid & 127, half h = tid >> 7); h splits the tau groups in stage 1 and the
// real | imaginary columns in the epilogue.  Warp w may touch TMEM lanes [32 (w & 3), +32) only — exactly its rows.
// Each A buffer is produced in two parts so that the MMAs of the first part run under stage 1 of the second.
template <int PREC, bool WANT_Z>
__device__ inline void decomp_umma_eval(DecompUmmaSmem& s, const DecompUmmaShape& sh, const double* __restrict__ prop,
                                        int ndim, int nrows, double* __restrict__ chi, double* __restrict__ Zout) {
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);               // warp-uniform by construction
  const int r = tid & (kUmmaRows - 1), h = warp >> 2;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  const bool warp_live = 32 * (warp & 3) < nrows;
  const uint32_t tbase = __shfl_sync(0xffffffffu, s.tbase, 0);
  const uint32_t tD = tbase, tA0 = tbase + 128;
  const uint32_t baddr = smem_u32(s.bar);
  const uint32_t lk_saddr = smem_u32(s.Lk);
  double R0 = 0.0;
  double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (r < nrows) {
    const double* q = prop + (size_t)r * ndim;
    R0 = q[0];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (i < sh.D) ? R0 * q[1 + i] : 0.0;
  }
  const uint32_t idesc = umma_idesc(sh.NC);
  UMMA_T0
  for (int c = 0; c < sh.nchunks; ++c) {
    const int k0 = c * kUmmaChunk, kt = min(kUmmaChunk, sh.SP - k0);      // taus of this chunk (multiple of 8)
    const int ng = kt >> 3, gsplit = min(ng, ((ng + 3) >> 2) << 1);
    const uint32_t tA = tA0 + (uint32_t)(c & 1) * 128;                    // hi at tA, lo at tA + 64
    const uint64_t dhi = umma_b_desc(smem_u32(s.Bhi) + 32 * k0, sh.SP);   // 32 bytes per tau: 256 per K step
    const uint64_t dlo = umma_b_desc(smem_u32(s.Blo) + 32 * k0, sh.SP);
    if (warp_live) {
      umma_stage1_dispatch<PREC>(sh.D, lk_saddr + (uint32_t)k0 * 64u, a, tA + lane_base, 0, gsplit, h);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    UMMA_MARK(0)
    tc_fence_before();
    __syncthreads();
    UMMA_MARK(1)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        umma_issue<PREC>(tD, tA, dhi, dlo, idesc, 0, gsplit, c == 0);
        if (gsplit == ng)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(baddr + 8 * (c & 1))
                       : "memory");
      }
      __syncwarp();
    }
    UMMA_MARK(2)
    if (gsplit < ng) {
      if (warp_live) {
        umma_stage1_dispatch<PREC>(sh.D, lk_saddr + (uint32_t)k0 * 64u, a, tA + lane_base, gsplit, ng, h);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      UMMA_MARK(0)
      tc_fence_before();
      __syncthreads();
      UMMA_MARK(1)
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          umma_issue<PREC>(tD, tA, dhi, dlo, idesc, gsplit, ng, false);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(baddr + 8 * (c & 1))
                       : "memory");
        }
        __syncwarp();
      }
      UMMA_MARK(2)
    }
    // Double-buffered A: chunk c+1 is written while chunk c is multiplied, chunk c+2 re-uses the buffer of chunk c —
    // so chunk c-1 is awaited here (one chunk behind) and the last chunk below.  One mbarrier per buffer: between
    // two commits on the same barrier every thread has passed its wait and a CTA barrier.
    if (c > 0) {
      const int bi = (c - 1) & 1;
      mbar_wait(baddr + 8 * bi, (s.phase >> bi) & 1u);
      s.phase ^= 1u << bi;
    }
  }
  {
    const int bi = (sh.nchunks - 1) & 1;
    mbar_wait(baddr + 8 * bi, (s.phase >> bi) & 1u);
    s.phase ^= 1u << bi;
  }
  tc_fence_after();
  UMMA_MARK(3)
  // ---- epilogue: D row r, columns [h NCH, (h+1) NCH) ---------------------------------------------------------
  // The accumulators are FP32, so the residual is formed in FP32 too: c = y/s - R0 d/s is evaluated with two-float
  // operands (head FMA + tail corrections), which keeps its error below the accumulator's own rounding.
  double acc = 0.0;
  if (warp_live) {
    const uint32_t td = tD + lane_base + (uint32_t)(h * sh.NCH);
    const float R0h = (float)R0, R0l = (float)(R0 - (double)R0h);
    float ch[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < sh.NCH; c0 += 16) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(td + c0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (WANT_Z) {
        if (r < nrows) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j = c0 + e;
            if (j < sh.N) Zout[(size_t)r * 2 * sh.N + h * sh.N + j] = (h ? 0.0 : R0) - (double)__uint_as_float(v[e]);
          }
        }
      } else {
        const uint32_t ca = smem_u32(s.col) + (uint32_t)(h * sh.NCH + c0) * 16u;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float ysh, ysl, dsh, dsl;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ysh), "=f"(ysl), "=f"(dsh), "=f"(dsl) : "r"(ca + e * 16));
          const float d = __uint_as_float(v[e]);
          float res;
          if (h == 0) {
            const float t = fmaf(-R0h, dsh, ysh) + d;
            const float corr = fmaf(-R0l, dsh, fmaf(-R0h, dsl, ysl));
            res = t + corr;
          } else {
            res = (ysh + d) + ysl;
          }
          ch[e & 3] = fmaf(res, res, ch[e & 3]);
        }
      }
    }
    acc = ((double)ch[0] + (double)ch[1]) + ((double)ch[2] + (double)ch[3]);
  }
  tc_fence_before();
  UMMA_MARK(4)
  if (!WANT_Z) {
    if (h == 1) s.part[r] = acc;
    __syncthreads();
    if (h == 0 && r < nrows) chi[r] = acc + s.part[r];
  }
  UMMA_MARK(5)
}

}  // namespace bisip
