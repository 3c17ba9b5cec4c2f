// Warp-private ensemble sampler: the same stretch move and the same Philox stream as sampler.cuh (and therefore as
// oracle/bisip_oracle.c), re-organised so that one half-step of one 16-proposal row tile — build the proposals, evaluate
// them, take the accept decisions — runs inside ONE WARP with __syncwarp() only.  The CTA synchronises once per half-step
// (the next half-step's proposals read walkers that other warps have just updated) instead of three times, the
// per-proposal chi^2 never travels through shared memory, and no thread waits at a barrier between its proposal and its
// evaluation.  That is what bounded every fast evaluator: at the C5 shape the collapsed FP64 kernel spent 41 % of a step
// in the serial phases between evaluations (profiles/r02_phase_cycles.md).
//
// Lane roles inside a warp (16 rows, two lanes per row: lane = 2 r + sub):
//   proposal   both lanes of a row build it, the dimensions interleaved between them (sub, sub + 2, ...)
//   evaluate   all lanes: DMMA tiles (CollapsedMmaEvaluator) or two lanes per row over the frequencies (VecWarpEvaluator)
//   accept     the sub == 0 lane takes the decision and updates the walker
//   draws      once per step every lane makes ONE Philox call: lane (r, sub) draws stretch factor, partner and acceptance
//              uniform of row r of half-step `sub` of the NEXT step (all lanes busy, no divergence); warps 0-1 also
//              draw the next step's shuffle keys
// The draws and the accept threshold are assembled on the integer / FP32 pipes (u53_int, f32_widen_int): in these phases
// every FP64-pipe instruction queues behind the other warps' tensor tiles.  The key ranking of the next split is cut into
// three parts (every warp scans the 256-bin histogram itself and hands the bin starts to its threads by shuffle) that run
// one step ahead, between the two half-step barriers: a step has TWO CTA barriers (three when it is stored) instead of
// eleven.
// CTA = 32 * ceil(W / 32) threads rounded up to a power of two (32 ... 256): every warp owns one row tile per half-step,
// a 32-walker ensemble is a single-warp CTA.  W <= 256; larger ensembles, clustered and tcgen05 evaluators use
// ensemble_kernel (sampler.cuh).
#pragma once
#include "sampler.cuh"

namespace bisip {

// Developer build (-DBISIP_PHASE_TIMING, `make dbg`): lane 0 of warp 0 of block 0 accumulates SM cycles per section of a
// half-step and prints them at the end.  Compiled out of the product library.
#ifdef BISIP_PHASE_TIMING
#define WP_T_DECL long long wt_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long wt0_ = clock64();
#define WP_T(i) { long long n_ = clock64(); wt_[i] += n_ - wt0_; wt0_ = n_; }
#define WP_T_PRINT if (blockIdx.x == 0 && threadIdx.x == 0) printf("wp cycles/step (warp 0): propose %.0f  prepare+sync %.0f  eval %.0f  accept %.0f  draws+keys+ranking %.0f  barrier-wait %.0f  store %.0f\n", \
    (double)wt_[0] / P.nsteps, (double)wt_[1] / P.nsteps, (double)wt_[2] / P.nsteps, (double)wt_[3] / P.nsteps, (double)wt_[4] / P.nsteps, \
    (double)wt_[5] / P.nsteps, (double)wt_[6] / P.nsteps);
#else
#define WP_T_DECL
#define WP_T(i)
#define WP_T_PRINT
#endif

struct WpSmem {
  double* coords;   // [W][ndim]
  double* lp;       // [W]
  double* prop;     // [rows_pad][ndim]
  double* zz;       // [2][2][rows_pad]  stretch factors: [step parity][half-step][row]
  double* red;      // [8]
  long long* bkey;  // [2][ndim]
  double* bnd;      // [2][ndim]
  float* lf;        // [2][2][rows_pad]  (ndim-1) ln zz - ln u2 in FP32 (accept filter); the rare FP64 fallback re-draws u2
  uint32_t* keys;   // [Wpad4]
  int* list;        // [2][W]            walker at rank, by step parity
  int* acc;         // [W]
  int* partner;     // [2][2][rows_pad]
  int* hist;        // [2][260]          key ranking: bucket counters, by parity of the step the keys belong to
  uint32_t* sorted; // [Wpad4]
};

__host__ __device__ inline int wp_rows_pad(int W) { return ceil_div((W + 1) / 2, 16) * 16; }

__host__ __device__ inline size_t wp_smem_bytes(int W, int ndim) {
  const int rp = wp_rows_pad(W);
  const size_t dbl = (size_t)W * ndim + W + (size_t)rp * ndim + 4 * (size_t)rp + 8 + 4 * (size_t)ndim;
  const size_t words = 4 * (size_t)rp + (size_t)(W + 4) + 2 * W + W + 4 * rp + 2 * 260 + (W + 4);
  return dbl * 8 + words * 4 + 64;
}

__device__ inline void wp_carve(WpSmem& s, double* base, int W, int ndim) {
  const int rp = wp_rows_pad(W);
  s.coords = base; base += (size_t)W * ndim;
  s.lp = base; base += W;
  s.prop = base; base += (size_t)rp * ndim;
  s.zz = base; base += 4 * rp;
  s.red = base; base += 8;
  s.bnd = base; base += 2 * ndim;
  s.bkey = reinterpret_cast<long long*>(base); base += 2 * ndim;
  s.lf = reinterpret_cast<float*>(base);
  s.keys = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(s.lf + 4 * rp) + 15) & ~uintptr_t(15));
  s.list = reinterpret_cast<int*>(s.keys + ((W + 3) & ~3));
  s.acc = s.list + 2 * W;
  s.partner = s.acc + W;
  s.hist = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(s.partner + 4 * rp) + 15) & ~uintptr_t(15));
  s.sorted = reinterpret_cast<uint32_t*>(s.hist + 2 * 260);
}

// ---- evaluators -------------------------------------------------------------------------------------------------------
// Interface: smem_doubles(desc) [host]; carve(base); init(desc, w, taus, log_taus, y, yerr, red) CTA-wide, ends with
// __syncthreads(); llconst(); prepare_row(q, theta) by the sub == 0 lane of row q; eval_warp(prop, ndim, row0, nrows,
// lane): all 32 lanes, returns chi^2 of row row0 + (lane >> 1) on both lanes of that row (rows >= nrows, beyond the
// half-step, are skipped or evaluated on whatever prop holds, and ignored by the caller).

// Collapsed decomposition z = (L K) a on FP64 tensor tiles.  The residual of proposal p at column c,
//   r_pc = (y_c - Z_pc)/s_c = y_c/s_c - R0 d_c/s_c + sum_i (R0 a_i) G_ic/s_c      (Z = R0 (d - z), z = G^T a),
// is ONE dot product of the proposal's row (1, R0, R0 a_0, ..., R0 a_{D-1}) with a per-column vector that does not
// depend on theta: a 16 x 8 tile of residuals is KS = ceil((D+2)/8) `mma.sync.m16n8k8.f64` instructions with the
// A fragment (the 16 proposals) resident in registers for all 2N/8 column tiles and one conflict-free LDS.128 per
// instruction for the B fragment.  The DMMA pipe runs at the DFMA rate, so the arithmetic costs the same as the
// register-tiled DFMA form (decomp_collapsed.cuh) — but in 1/5 of the instructions and 2/5 of the shared-memory
// wavefronts, which were what bounded that form (issue slots 57 %, LSU wavefronts 61 %, FP64 pipe 36 %).
// D <= 30 coefficients (poly_deg <= 29).
template <int KS>
struct CollapsedMmaEvaluator {
  static constexpr bool kNeedsPrepare = false;
  double* Bf;       // [nct][KS][32][2]  B fragments: (k, column) -> lane 4 (c & 7) + (k & 3), slot (k >> 2) & 1
  double llc;
  int N, S, D, nct;
  __device__ CollapsedMmaEvaluator(const bisip_model_desc& d) : N(d.n_freq), S(d.n_tau), D(d.n_coef), nct((2 * d.n_freq + 7) >> 3) {}
  static __host__ size_t smem_doubles(const bisip_model_desc& d) { return (size_t)((2 * d.n_freq + 7) >> 3) * KS * 64; }
  __device__ double* carve(double* base) { Bf = base; return base + (size_t)nct * KS * 64; }
  __device__ __forceinline__ size_t bf_index(int k, int c) const {
    return ((size_t)((c >> 3) * KS + (k >> 3)) * 32 + 4 * (c & 7) + (k & 3)) * 2 + ((k >> 2) & 1);
  }
  __device__ void init(const bisip_model_desc& d, const double* __restrict__ w, const double* __restrict__ taus,
                       const double* __restrict__ log_taus, const double* __restrict__ y, const double* __restrict__ yerr,
                       double* red) {
    const int tid = threadIdx.x, NT = blockDim.x, C = 2 * N;
    for (int i = tid; i < nct * KS * 64; i += NT) Bf[i] = 0.0;
    __syncthreads();
    double csum = 0.0;
    for (int c = tid; c < C; c += NT) {
      const double e = yerr[c];
      const double is = 1.0 / e;
      Bf[bf_index(0, c)] = y[c] * is;
      Bf[bf_index(1, c)] = (c < N) ? -is : 0.0;
      csum += 2.0 * log(e * e);
    }
    double cs, sn;
    sincospi(0.5 * d.c_exp, &sn, &cs);
    for (int idx = tid; idx < C * D; idx += NT) {
      const int c = idx / D, i = idx - c * D;
      const int j = c < N ? c : c - N;
      const double wj = w[j];
      const double is = 1.0 / yerr[c];
      const double* lt = log_taus + (size_t)i * S;
      double sum = 0.0, comp = 0.0;                  // Dot2 (Ogita-Rump-Oishi), as decomp_c_init
      for (int k = 0; k < S; ++k) {
        double kre, kim;
        debye_kernel_term(wj, taus[k], d.c_exp, cs, sn, kre, kim);
        const double kv = (c < N ? kre : kim) * is;
        const double l = lt[k];
        const double p = __dmul_rn(l, kv);
        const double pe = __fma_rn(l, kv, -p);
        const double t = __dadd_rn(sum, p);
        const double bb = __dsub_rn(t, sum);
        const double se = __dadd_rn(__dsub_rn(sum, __dsub_rn(t, bb)), __dsub_rn(p, bb));
        sum = t;
        comp = __dadd_rn(comp, __dadd_rn(se, pe));
      }
      Bf[bf_index(2 + i, c)] = sum + comp;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
    if ((tid & 31) == 0) red[tid >> 5] = csum;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < (NT >> 5); ++i) tot += red[i];
    llc = tot;
    __syncthreads();
  }
  __device__ double llconst() const { return llc; }
  __device__ __forceinline__ void prepare_row(int, const double*) {}
  __device__ __forceinline__ double eval_warp(const double* __restrict__ prop, int ndim, int row0, int, int lane) const {
    const int g = lane >> 2, t = lane & 3;
    const double* q0 = prop + (size_t)(row0 + g) * ndim;
    const double* q1 = q0 + 8 * ndim;
    const double R0a = q0[0], R0b = q1[0];
    // A fragment: a[i] -> row g + 8 (i & 1), k = 8 ks + t + 4 (i >> 1) ; row vector (1, R0, R0 a_0 ... R0 a_{D-1}, 0 ...)
    double a[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 8 * ks + t + 4 * h;
        double va, vb;
        if (k == 0) { va = 1.0; vb = 1.0; }
        else if (k == 1) { va = R0a; vb = R0b; }
        else if (k - 2 < D) { va = R0a * q0[k - 1]; vb = R0b * q1[k - 1]; }
        else { va = 0.0; vb = 0.0; }
        a[ks][2 * h] = va;
        a[ks][2 * h + 1] = vb;
      }
    }
    const double2* bf = reinterpret_cast<const double2*>(Bf) + lane;
    double chi_lo = 0.0, chi_hi = 0.0;
#pragma unroll 4
    for (int ct = 0; ct < nct; ++ct) {
      double c[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const double2 b2 = bf[(size_t)(ct * KS + ks) * 32];
        const double b[2] = {b2.x, b2.y};
        dmma_16x8x8(c, a[ks], b);
      }
      chi_lo = fma(c[0], c[0], chi_lo);
      chi_hi = fma(c[2], c[2], chi_hi);
      chi_lo = fma(c[1], c[1], chi_lo);
      chi_hi = fma(c[3], c[3], chi_hi);
    }
    chi_lo += __shfl_xor_sync(0xffffffffu, chi_lo, 1);
    chi_hi += __shfl_xor_sync(0xffffffffu, chi_hi, 1);
    chi_lo += __shfl_xor_sync(0xffffffffu, chi_lo, 2);
    chi_hi += __shfl_xor_sync(0xffffffffu, chi_hi, 2);
    // quad g holds rows g (lo) and g + 8 (hi); lane 2 r + sub wants row r
    const int r = lane >> 1;
    const double lo = __shfl_sync(0xffffffffu, chi_lo, 4 * (r & 7));
    const double hi = __shfl_sync(0xffffffffu, chi_hi, 4 * (r & 7));
    return (r >> 3) ? hi : lo;
  }
};

// The default two-stage contraction on FP64 tensor tiles (decomp_eval.cuh), one 16-proposal row tile per warp: stage 1
// leaves the chargeability in accumulator layout = the stage-2 A fragment, stage 2 streams the fragment-ordered kernel
// matrix, the epilogue squares the residuals on the accumulators.  The same instructions as decomp_eval_chi, without
// the shared-memory round trip of chi^2 and without CTA barriers around the evaluation.
template <int KC>
struct DecompMmaWarpEvaluator {
  static constexpr bool kNeedsPrepare = false;
  DecompSmem sm;
  DecompShape sh;
  __device__ DecompMmaWarpEvaluator(const bisip_model_desc& d) : sh(d.n_freq, d.n_tau, d.n_coef) {}
  static __host__ size_t smem_doubles(const bisip_model_desc& d) {
    const DecompShape h(d.n_freq, d.n_tau, d.n_coef);
    return h.kf_doubles() + h.l1_doubles() + 2 * h.col_doubles();
  }
  __device__ double* carve(double* base) {
    sm.Kf = base; base += sh.kf_doubles();
    sm.L1 = base; base += sh.l1_doubles();
    sm.ycol = base; base += sh.col_doubles();
    sm.isig = base; base += sh.col_doubles();
    sm.part = nullptr;
    return base;
  }
  __device__ void init(const bisip_model_desc& d, const double* w, const double* taus, const double* log_taus,
                       const double* y, const double* yerr, double* red) {
    decomp_init(sm, sh, d.c_exp, w, taus, log_taus, y, yerr, red);
  }
  __device__ double llconst() const { return sm.llconst; }
  __device__ __forceinline__ void prepare_row(int, const double*) {}
  __device__ __forceinline__ double eval_warp(const double* __restrict__ prop, int ndim, int row0, int, int lane) const {
    const int t = lane & 3;
    double A[KC][8];
    double R0a, R0b;
    decomp_stage1<KC>(sm, sh.D, prop, ndim, row0 >> 4, lane, A, R0a, R0b);
    double chi_lo = 0.0, chi_hi = 0.0;
#pragma unroll 2
    for (int nt = 0; nt < sh.NT2; ++nt) {
      const int col = nt * 8 + 2 * t;
      const double2 ys = *reinterpret_cast<const double2*>(sm.ycol + col);
      double c[4];
      if (nt * 8 >= sh.N) {          // tile entirely in the imaginary block: delta = 0 (warp-uniform)
        c[0] = ys.x; c[1] = ys.y; c[2] = ys.x; c[3] = ys.y;
      } else {
        const double2 ds = *reinterpret_cast<const double2*>(sm.isig + col);
        c[0] = fma(-R0a, ds.x, ys.x);
        c[1] = fma(-R0a, ds.y, ys.y);
        c[2] = fma(-R0b, ds.x, ys.x);
        c[3] = fma(-R0b, ds.y, ys.y);
      }
      decomp_stage2_tile<KC>(sm, nt, lane, A, c);      // c = (y - Z)/sigma
      chi_lo = fma(c[0], c[0], chi_lo);
      chi_lo = fma(c[1], c[1], chi_lo);
      chi_hi = fma(c[2], c[2], chi_hi);
      chi_hi = fma(c[3], c[3], chi_hi);
    }
    chi_lo += __shfl_xor_sync(0xffffffffu, chi_lo, 1);
    chi_lo += __shfl_xor_sync(0xffffffffu, chi_lo, 2);
    chi_hi += __shfl_xor_sync(0xffffffffu, chi_hi, 1);
    chi_hi += __shfl_xor_sync(0xffffffffu, chi_hi, 2);
    const int r = lane >> 1;
    const double lo = __shfl_sync(0xffffffffu, chi_lo, 4 * (r & 7));
    const double hi = __shfl_sync(0xffffffffu, chi_hi, 4 * (r & 7));
    return (r >> 3) ? hi : lo;
  }
};

// Cole-Cole / Dias / Shin: two lanes per proposal, each over every second frequency (the loop body of vec_eval_chi),
// ILP frequencies in flight per lane (their exp / reciprocal chains are latency-bound).
template <class Row, int ILP = 2>
struct VecWarpEvaluator {
  static constexpr bool kNeedsPrepare = true;
  VecSmem sm;
  int N, n_modes, rows_pad;
  __device__ VecWarpEvaluator(const bisip_model_desc& d) : N(d.n_freq), n_modes(d.n_modes) {}
  static __host__ size_t smem_doubles(const bisip_model_desc& d, int rows_pad) { return vec_smem_doubles(d.n_freq, rows_pad, Row::kRC); }
  __device__ double* carve(double* base, int rp) { rows_pad = rp; return vec_carve(sm, base, N, rp, Row::kRC); }
  __device__ void init(const bisip_model_desc&, const double* w, const double*, const double*, const double* y,
                       const double* yerr, double* red) {
    vec_init(sm, N, w, y, yerr, red);
  }
  __device__ double llconst() const { return sm.llconst; }
  __device__ __forceinline__ void prepare_row(int q, const double* th) { Row::prepare(th, n_modes, sm.rowc + (size_t)q * Row::kRC); }
  __device__ __forceinline__ double eval_warp(const double*, int, int row0, int nrows, int lane) const {
    const int row = row0 + (lane >> 1), sub = lane & 1;
    constexpr int lpr = 2, stride = lpr * kFq;
    Row rr;
    rr.load(sm.rowc + (size_t)(row < nrows ? row : row0) * Row::kRC, n_modes);
    const double* f = sm.fq + sub * kFq;
    const int j0 = row < nrows ? sub : N;                        // rows beyond the half-step: no work
    bool ok = true;
    double tot = vec_row_chi<Row, ILP, true>(rr, f, j0, N, lpr, stride, ok);
    if (!ok) tot = vec_row_chi<Row, 1, false>(rr, f, j0, N, lpr, stride, ok);    // rare: see vec_row_chi
    tot += __shfl_xor_sync(0xffffffffu, tot, 1);
    return tot;
  }
};

// carve adaptors (the two evaluator families size their shared memory differently)
template <int KS>
__device__ __forceinline__ double* wp_eval_carve(CollapsedMmaEvaluator<KS>& ev, double* base, int) { return ev.carve(base); }
template <class Row, int ILP>
__device__ __forceinline__ double* wp_eval_carve(VecWarpEvaluator<Row, ILP>& ev, double* base, int rp) { return ev.carve(base, rp); }
template <int KC>
__device__ __forceinline__ double* wp_eval_carve(DecompMmaWarpEvaluator<KC>& ev, double* base, int) { return ev.carve(base); }

template <class Eval, int MINB, int NT>
__global__ void __launch_bounds__(NT, MINB) ensemble_wp_kernel(const EnsembleParams P) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int W = P.W, ndim = P.d.ndim;
  const int rows_pad = wp_rows_pad(W);
  const int H0 = (W + 1) / 2;
  const int r = lane >> 1, sub = lane & 1;
  const int row0 = 16 * warp, q = row0 + r;         // this lane's proposal row in every half-step

  // (a start-time stagger of the two co-resident CTAs of the DMMA kernel, as ensemble_kernel has, changes nothing here:
  //  1.750e9 +- 0.1 % for 0 ... 20 us)
  Eval ev(P.d);
  WpSmem s;
  double* p = wp_eval_carve(ev, smem, rows_pad);
  wp_carve(s, p, W, ndim);
  int* const counts = s.hist;                        // [2][260] bin counters of the key ranking

  // ---- per-spectrum constants + initial ensemble ---------------------------------------
  for (int i = tid; i < 2 * ndim; i += NT) {
    s.bnd[i] = P.bounds[i];
    s.bkey[i] = ordered_key(P.bounds[i]);
  }
  const double* gc = P.coords + (size_t)b * W * ndim;
  for (int i = tid; i < W * ndim; i += NT) s.coords[i] = gc[i];
  for (int i = tid; i < W; i += NT) s.acc[i] = 0;
  for (int i = tid; i < rows_pad * ndim; i += NT) s.prop[i] = 0.0;
  for (int i = tid; i < 2 * 260; i += NT) s.hist[i] = 0;
  ev.init(P.d, P.w + (size_t)b * P.w_stride, P.taus + (size_t)b * P.tau_stride,
          P.log_taus + (size_t)b * P.tau_stride * P.d.n_coef, P.y + (size_t)b * 2 * P.d.n_freq,
          P.yerr + (size_t)b * 2 * P.d.n_freq, s.red);   // ends with __syncthreads()
  const double llc = ev.llconst();
  int flag = 0;

  // log-probability of p0, in two passes of <= H0 rows (same row -> lane map as the sampling loop)
  for (int pass = 0; pass < 2; ++pass) {
    const int off = pass ? H0 : 0, n = pass ? W - H0 : H0;
    for (int i = tid; i < n * ndim; i += NT) s.prop[i] = s.coords[off * ndim + i];
    __syncthreads();
    if (row0 < n) {
      if (Eval::kNeedsPrepare) {
        if (sub == 0 && q < n) ev.prepare_row(q, s.prop + q * ndim);
        __syncwarp();
      }
      const double chi = ev.eval_warp(s.prop, ndim, row0, n, lane);
      if (sub == 0 && q < n) {
        const double v = in_bounds(s.prop + q * ndim, s.bnd, ndim) ? -0.5 * (chi + llc) : neg_inf();
        if (v != v) flag |= 2;
        s.lp[off + q] = v;
      }
    }
    __syncthreads();
  }

  const uint32_t k0 = (uint32_t)P.seed, k1 = (uint32_t)(P.seed >> 32);
  const uint32_t spec = P.spectrum0 + (uint32_t)b;
  const int first = P.discard + P.thin - 1;
  const int Wpad4 = (W + 3) & ~3;
  const bool a_is2 = P.a == 2.0;
  const float am1f = (float)(P.a - 1.0), inv_af = (float)P.inv_a, nd1f = (float)(ndim - 1);
  int kept = 0;

  // shuffle keys of step t, chunk c (4 walkers per Philox call)
  auto gen_keys_chunk = [&](uint32_t t, int c) {
    const u32x4 rr = philox4x32_10((uint32_t)c, t, spec, 0u, k0, k1);
    const uint32_t wd[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 4 * c + e;
      s.keys[i] = i < W ? ((wd[e] & ~P.kmask) | (uint32_t)i) : 0xffffffffu;   // padding sorts last
    }
  };
  // all draws of proposal row q of half-step (t, sp = sub): one Philox call per lane, every lane busy
  auto gen_draws = [&](uint32_t t) {
    const int sp = sub;
    const int Hs = sp ? W - H0 : H0, Nc = W - Hs;
    if (q >= Hs) return;
    const u32x4 rr = philox4x32_10((uint32_t)q, t, spec, (uint32_t)(1 + sp), k0, k1);
    const double u = u53_int(rr.x, rr.y);
    double zz;
    if (a_is2) {                                      // ((a-1) u + 1)^2 / a with a = 2: (a-1) u = u and /2 are exact
      const double zr = __dadd_rn(u, 1.0);
      const double sq = __dmul_rn(zr, zr);            // in [1, 4): halving = one exponent step
      zz = __hiloint2double(__double2hiint(sq) - 0x00100000, __double2loint(sq));
    } else {
      const double zr = __dadd_rn(__dmul_rn(P.a - 1.0, u), 1.0);
      zz = P.a_pow2 ? __dmul_rn(__dmul_rn(zr, zr), P.inv_a) : __ddiv_rn(__dmul_rn(zr, zr), P.a);
    }
    const uint32_t zlow = (rr.z << 16) | 0x8000u;
    const int o = (((int)(t & 1u)) * 2 + sp) * rows_pad + q;
    s.zz[o] = zz;
    s.partner[o] = (int)__umulhi(rr.z, (uint32_t)Nc);
    // FP32 image of (ndim-1) ln zz - ln u2 straight from the random words (no FP64 conversions): |error| < 1e-5,
    // far inside the 2^-12 margin below which accept_filter() hands the decision to the FP64 logarithms
    const float uf = (float)(rr.x >> 8) * 5.9604644775390625e-08f;                    // 2^-24
    const float zrf = fmaf(am1f, uf, 1.0f);
    const float zzf = zrf * zrf * inv_af;
    const float u2f = fmaf((float)(zlow >> 6), 1.1102230246251565e-16f,                // 2^-53
                           (float)(rr.w >> 5) * 7.450580596923828e-09f);              // 2^-27
    s.lf[o] = nd1f * __logf(zzf) - __logf(u2f);
  };

  // The split of step t+1 is ranked one step AHEAD of its use, its three parts riding on the two barriers a step has anyway:
  //   step t-1, second half-step   part A of keys(t+1)  (bin slots; the keys were drawn in that step's first half-step)
  //   step t,   first half-step    part B               (scan + placement)
  //   step t,   second half-step   part C               (rank -> list of step t+1), then part A of keys(t+2)
  // counts[c] serves the keys of steps with parity c; it is zeroed in the first half-step of the step after its last use.
  // prologue: the split of step 0 in full, parts A of keys(step0 + 1), and the draws of both half-steps of step 0
  const uint32_t t0 = (uint32_t)P.step0;
  uint32_t rk_key = 0;
  int rk_bin = 0, rk_slot = 0, rk_start = 0;
  for (int c = tid; c < Wpad4 / 4; c += NT) gen_keys_chunk(t0, c);
  __syncthreads();
  {
    int* cnt0 = counts + (t0 & 1u) * 260;
    rank_part_a<NT>(s.keys, W, cnt0, rk_key, rk_bin, rk_slot);
    __syncthreads();
    rk_start = rank_part_b<NT>(W, cnt0, s.sorted, rk_key, rk_bin, rk_slot);
    __syncthreads();
    rank_part_c<NT>(W, cnt0, s.sorted, s.list + (size_t)(t0 & 1u) * W, rk_key, rk_bin, rk_start);
    for (int c = tid; c < Wpad4 / 4; c += NT) gen_keys_chunk(t0 + 1u, c);       // keys[] was last read in part A above
    __syncthreads();
    for (int i = tid; i < 260; i += NT) cnt0[i] = 0;
    rank_part_a<NT>(s.keys, W, counts + ((t0 + 1u) & 1u) * 260, rk_key, rk_bin, rk_slot);
  }
  gen_draws(t0);
  __syncthreads();

  WP_T_DECL
  for (int it = 0; it < P.nsteps; ++it) {
    const uint32_t t = (uint32_t)(P.step0 + it);
    const int par = (int)(t & 1u);
    const int* list = s.list + (size_t)par * W;               // this step's split
    int* list_next = s.list + (size_t)(par ^ 1) * W;
    int* cnt_next = counts + (par ^ 1) * 260;                 // bin counts of keys(t+1) (part A ran one half-step ago)
    int* cnt_after = counts + par * 260;                      // will count keys(t+2)
    for (int sp = 0; sp < 2; ++sp) {
      const int off = sp ? H0 : 0, Hs = sp ? W - H0 : H0;
      const int coff = sp ? 0 : H0;
      const int dq = (par * 2 + sp) * rows_pad + q;
      bool inb = false;
      int k = 0;
      if (row0 < Hs) {
        // ---- this warp's 16 rows: propose | evaluate | accept, no CTA barrier in between ---------------------
        // rows beyond the half-step (only in its last, partly filled tile) rebuild this warp's first row into their own
        // padding row: no divergent region around the proposal, and no read of a walker another warp may be updating
        const int qe = q < Hs ? q : row0;
        const int dqe = dq - q + qe;
        const int j = list[coff + s.partner[dqe]];
        k = list[off + qe];
        const bool ok = propose_pair(s.coords + j * ndim, s.coords + k * ndim, s.zz[dqe], s.prop + q * ndim, s.bkey, ndim, sub);
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        inb = ((okm >> (lane & ~1)) & 3u) == 3u;
        WP_T(0)
        if (Eval::kNeedsPrepare) {
          __syncwarp();
          if (sub == 0 && q < Hs) ev.prepare_row(q, s.prop + q * ndim);
        }
        __syncwarp();
        WP_T(1)
        const double chi = ev.eval_warp(s.prop, ndim, row0, Hs, lane);
        WP_T(2)
        if (sub == 0 && q < Hs) {
          const double lpo = s.lp[k];
          const double lpn = inb ? -0.5 * (chi + llc) : neg_inf();
          if (lpn != lpn) flag |= 1;
          // emcee: accept iff (ndim-1) ln zz + lp' - lp > ln u (see accept_filter in sampler.cuh)
          const double est = __dsub_rn(lpn, lpo) + f32_widen_int(s.lf[dq]);
          bool accept;
          if (!accept_filter(est, lpn, lpo, accept)) {
            const u32x4 rr = philox4x32_10((uint32_t)q, t, spec, (uint32_t)(1 + sp), k0, k1);    // re-draw u2
            const double lnpdiff = __dsub_rn(__dadd_rn(__dmul_rn((double)(ndim - 1), log(s.zz[dq])), lpn), lpo);
            accept = lnpdiff > log(u53_int(rr.w, (rr.z << 16) | 0x8000u));
          }
          if (accept) {
            copy_dims(s.coords + k * ndim, s.prop + q * ndim, ndim);
            s.lp[k] = lpn;
            s.acc[k] += 1;
          }
        }
      }
      WP_T(3)
      if (sp == 0) {
        // the random numbers of the NEXT step (every lane); the shuffle keys of the step after it (the first Wpad4/4
        // threads; keys[] was last read one half-step ago); scan + placement of the next step's ranking
        gen_draws(t + 1u);
        if (tid < Wpad4 / 4) gen_keys_chunk(t + 2u, tid);
        rk_start = rank_part_b<NT>(W, cnt_next, s.sorted, rk_key, rk_bin, rk_slot);
        for (int i = tid; i < 260; i += NT) cnt_after[i] = 0;
      } else {
        rank_part_c<NT>(W, cnt_next, s.sorted, list_next, rk_key, rk_bin, rk_start);      // the split of step t+1
        rank_part_a<NT>(s.keys, W, cnt_after, rk_key, rk_bin, rk_slot);                   // keys(t+2): bin slots
      }
      WP_T(4)
      __syncthreads();
      WP_T(5)
    }
    // ---- backend.save_step: chain[it] = coords ; log_prob[it] = lp; the barrier separates these reads of coords from the
    //      next half-step's accepts (a warp does not wait for the others between its proposal and its accept any more) --
    if (it >= first && (it - first) % P.thin == 0) {
      if (P.chain != nullptr) {
        double* dst = P.chain + ((size_t)b * P.nkeep + kept) * W * ndim;
        for (int i = tid; i < W * ndim; i += NT) __stcs(dst + i, s.coords[i]);
      }
      if (P.logp != nullptr) {
        double* dst = P.logp + ((size_t)b * P.nkeep + kept) * W;
        for (int i = tid; i < W; i += NT) __stcs(dst + i, s.lp[i]);
      }
      ++kept;
      __syncthreads();
    }
    WP_T(6)
  }
  WP_T_PRINT
  // ---- final state ------------------------------------------------------------------------------
  double* gco = P.coords + (size_t)b * W * ndim;
  for (int i = tid; i < W * ndim; i += NT) gco[i] = s.coords[i];
  for (int i = tid; i < W; i += NT) {
    if (P.lp) P.lp[(size_t)b * W + i] = s.lp[i];
    if (P.accepted) P.accepted[(size_t)b * W + i] = s.acc[i];
  }
  if (P.flags && flag) atomicOr(P.flags + b, flag);
}

}  // namespace bisip
