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
  double2* col;       // [NC]      (y/sigma, delta/sigma) per column
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
  s.col = reinterpret_cast<double2*>(p); p += (size_t)sh.NC * 16;
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
    s.col[n] = make_double2(ys, ds);
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

// stage 1 of one 8-tau group for this thread's row -> hi / lo A planes in tensor memory
template <int ND, int PREC>
__device__ __forceinline__ void umma_stage1_group(const double* __restrict__ Lk8, const double (&a)[8], uint32_t ta_hi,
                                                  uint32_t ta_lo) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const double* L = Lk8 + e * 8;
    double m = 0.0;
#pragma unroll
    for (int i = 0; i < ND; ++i) m = fma(a[i], L[i], m);
    split_tf32_umma(m, hi[e], lo[e]);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta_hi), "r"(hi[0]), "r"(hi[1]),
               "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7])
               : "memory");
  if (PREC == 3)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta_lo), "r"(lo[0]), "r"(lo[1]),
                 "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7])
                 : "memory");
}

template <int PREC>
__device__ __forceinline__ void umma_stage1_dispatch(int ND, const double* Lk8, const double (&a)[8], uint32_t th, uint32_t tl) {
  switch (ND) {
    case 1: umma_stage1_group<1, PREC>(Lk8, a, th, tl); break;
    case 2: umma_stage1_group<2, PREC>(Lk8, a, th, tl); break;
    case 3: umma_stage1_group<3, PREC>(Lk8, a, th, tl); break;
    case 4: umma_stage1_group<4, PREC>(Lk8, a, th, tl); break;
    case 5: umma_stage1_group<5, PREC>(Lk8, a, th, tl); break;
    case 6: umma_stage1_group<6, PREC>(Lk8, a, th, tl); break;
    case 7: umma_stage1_group<7, PREC>(Lk8, a, th, tl); break;
    default: umma_stage1_group<8, PREC>(Lk8, a, th, tl); break;
  }
}

// Evaluate nrows <= 128 proposals.  WANT_Z = false: chi[q] = sum_c ((y_c - Z_c)/sigma_c)^2 (caller barriers
// before reading); WANT_Z = true: Zout[q][2][N] (forward only; init was called with y == nullptr).
// 256 threads: thread = (row r = tid & 127, half h = tid >> 7); h splits the tau groups in stage 1 and the
// real | imaginary columns in the epilogue.  Warp w may touch TMEM lanes [32 (w & 3), +32) only — exactly its rows.
template <int PREC, bool WANT_Z>
__device__ inline void decomp_umma_eval(DecompUmmaSmem& s, const DecompUmmaShape& sh, const double* __restrict__ prop,
                                        int ndim, int nrows, double* __restrict__ chi, double* __restrict__ Zout) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & (kUmmaRows - 1), h = tid >> 7;
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  const bool warp_live = 32 * (warp & 3) < nrows;
  const uint32_t tD = s.tbase, tA0 = s.tbase + 128;
  const uint32_t baddr = smem_u32(s.bar);
  double R0 = 0.0;
  double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (r < nrows) {
    const double* q = prop + (size_t)r * ndim;
    R0 = q[0];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (i < sh.D) ? R0 * q[1 + i] : 0.0;
  }
  const uint32_t idesc = umma_idesc(sh.NC);
  for (int c = 0; c < sh.nchunks; ++c) {
    const int k0 = c * kUmmaChunk, kt = min(kUmmaChunk, sh.SP - k0);      // taus of this chunk (multiple of 8)
    const uint32_t tA = tA0 + (uint32_t)(c & 1) * 128;                    // hi at tA, lo at tA + 64
    if (warp_live) {
      for (int g = h; g < (kt >> 3); g += 2)
        umma_stage1_dispatch<PREC>(sh.D, s.Lk + (size_t)(k0 + 8 * g) * 8, a, tA + lane_base + 8 * g, tA + 64 + lane_base + 8 * g);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t bhi = smem_u32(s.Bhi) + 32 * k0, blo = smem_u32(s.Blo) + 32 * k0;   // 256 bytes per K step of 8 taus
      const int ks = kt >> 3;
      uint32_t acc = c > 0 ? 1u : 0u;
      if (PREC == 3) {
        for (int j = 0; j < ks; ++j) { umma_tf32_ts(tD, tA + 64 + 8 * j, umma_b_desc(bhi + 256 * j, sh.SP), idesc, acc); acc = 1u; }
        for (int j = 0; j < ks; ++j) umma_tf32_ts(tD, tA + 8 * j, umma_b_desc(blo + 256 * j, sh.SP), idesc, 1u);
      }
      for (int j = 0; j < ks; ++j) { umma_tf32_ts(tD, tA + 8 * j, umma_b_desc(bhi + 256 * j, sh.SP), idesc, acc); acc = 1u; }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(baddr + 8 * (c & 1))
                   : "memory");
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
  // ---- epilogue: D row r, columns [h NCH, (h+1) NCH) ---------------------------------------------------------
  double acc = 0.0;
  if (warp_live) {
    const uint32_t td = tD + lane_base + (uint32_t)(h * sh.NCH);
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
        const double2* col = s.col + h * sh.NCH + c0;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const double2 yd = col[e];
          const double res = fma(-R0, yd.y, yd.x) + (double)__uint_as_float(v[e]);
          acc = fma(res, res, acc);
        }
      }
    }
  }
  tc_fence_before();
  if (!WANT_Z) {
    if (h == 1) s.part[r] = acc;
    __syncthreads();
    if (h == 0 && r < nrows) chi[r] = acc + s.part[r];
  }
}

}  // namespace bisip
