// PolynomialDecomposition evaluator for LARGE tau grids (n_tau > 64: the reference default
// n_tau = 2N for N >= 33, and the 256-tau study of BASELINE config 4).
//
// Same two-stage FP64 DMMA contraction as decomp_eval.cuh, with two changes forced by size:
//  * the stage-2 A operand (M = chargeability over the tau grid, 16 rows x S) no longer fits in
//    registers, so it is RE-COMPUTED chunk by chunk (16 taus = two m16n8k8 stage-1 tiles) inside the
//    k loop while up to 8 column tiles (32 FP64 accumulators per thread) stay resident;
//  * K (S x 2N doubles; 256 KB at S=256, N=64) no longer fits in one SM's shared memory, so the 2N
//    frequency columns are SPLIT ACROSS A THREAD-BLOCK CLUSTER: CTA r of the cluster stores and
//    contracts column tiles [r*NTC, (r+1)*NTC) only.  Every CTA of the cluster runs the same sampler
//    on an identical copy of the walkers; the per-proposal partial chi^2 are exchanged through
//    distributed shared memory once per half-step (one cluster barrier), summed in rank order so all
//    CTAs see bit-identical log-probabilities and take identical accept/reject decisions.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "decomp_eval.cuh"

namespace bisip {

namespace cg = cooperative_groups;

constexpr int kRcMaxTiles = 8;   // column tiles resident per work item (32 accumulators / thread)

struct DecompRCShape {
  int N, S, D, NT2, KC;
  int CS, rank;      // cluster size / this CTA's rank
  int NTC;           // column tiles owned per CTA = ceil(NT2 / CS)
  int nt_lo, ntc;    // first owned tile, number of owned tiles (last rank may own fewer)
  int TPG, NGC;      // tiles per work item (<= 8), work-item groups per row tile
  __host__ __device__ DecompRCShape(int n, int s, int d, int cs, int r)
      : N(n), S(s), D(d), NT2(ceil_div(2 * n, 8)), KC(ceil_div(s, 16)), CS(cs), rank(r) {
    NTC = ceil_div(NT2, CS);
    nt_lo = r * NTC;
    ntc = NT2 - nt_lo < NTC ? NT2 - nt_lo : NTC;
    if (ntc < 0) ntc = 0;
    NGC = ceil_div(NTC, kRcMaxTiles);
    TPG = ceil_div(NTC, NGC);
  }
  __host__ __device__ size_t kf_doubles() const { return (size_t)NTC * KC * 128; }
  __host__ __device__ size_t l1_doubles() const { return (size_t)KC * 2 * 2 * 32; }
  __host__ __device__ size_t col_doubles() const { return (size_t)NT2 * 8; }
};

struct DecompRCSmem {
  double* Kf;    // [NTC][KC][2][32][2]  owned column tiles only, pre-scaled by 1/sigma (likelihood mode)
  double* L1;    // [2KC][2][32]
  double* ycol;  // [NT2*8] y/sigma
  double* isig;  // [NT2*8] delta/sigma
  double* part;  // [NGC][rows_pad]
  double* xsum;  // [2][rows_pad]  this CTA's column-partial chi^2, double-buffered for the DSMEM exchange
  double llconst;
  int parity;
};

__host__ __device__ inline size_t decomp_rc_smem_doubles(const DecompRCShape& sh, int rows_pad) {
  return sh.kf_doubles() + sh.l1_doubles() + 2 * sh.col_doubles() + (size_t)(sh.NGC + 2) * rows_pad;
}

__device__ inline double* decomp_rc_carve(DecompRCSmem& s, double* base, const DecompRCShape& sh, int rows_pad) {
  s.Kf = base; base += sh.kf_doubles();
  s.L1 = base; base += sh.l1_doubles();
  s.ycol = base; base += sh.col_doubles();
  s.isig = base; base += sh.col_doubles();
  s.part = base; base += (size_t)sh.NGC * rows_pad;
  s.xsum = base; base += (size_t)2 * rows_pad;
  s.parity = 0;
  return base;
}

__device__ inline void decomp_rc_init(DecompRCSmem& s, const DecompRCShape& sh, double c_exp,
                                      const double* __restrict__ w, const double* __restrict__ taus,
                                      const double* __restrict__ log_taus, const double* __restrict__ y,
                                      const double* __restrict__ yerr, double* red) {
  const int tid = threadIdx.x;
  const int N = sh.N, S = sh.S, KC = sh.KC;
  for (int i = tid; i < (int)sh.kf_doubles(); i += kThreads) s.Kf[i] = 0.0;
  for (int i = tid; i < (int)sh.l1_doubles(); i += kThreads) {
    const int lane = i & 31, q = (i >> 5) & 1, j = i >> 6;
    const int g = lane >> 2, t = lane & 3;
    const int p = t + 4 * q, k = 8 * j + g;
    s.L1[i] = (p < sh.D && k < S) ? log_taus[(size_t)p * S + k] : 0.0;
  }
  const bool scaled = (y != nullptr);
  double csum = 0.0;
  for (int c = tid; c < (int)sh.col_doubles(); c += kThreads) {
    double ys = 0.0, ds = 0.0;
    if (c < 2 * N && scaled) {
      const double e = yerr[c];
      const double is = 1.0 / e;
      ys = y[c] * is;
      ds = (c < N) ? is : 0.0;
      csum += 2.0 * log(e * e);
    }
    s.ycol[c] = ys;
    s.isig[c] = ds;
  }
  __syncthreads();
  double cs, sn;
  sincospi(0.5 * c_exp, &sn, &cs);
  // owned columns [c_lo, c_hi): column c is Re(freq c) for c < N, Im(freq c-N) otherwise
  const int c_lo = sh.nt_lo * 8, ncols = min(2 * N, (sh.nt_lo + sh.ntc) * 8) - c_lo;
  for (int i = tid; i < S * max(ncols, 0); i += kThreads) {
    const int k = i / ncols, c = c_lo + (i - k * ncols);
    const int j = c < N ? c : c - N;
    double kre, kim;
    debye_kernel_term(w[j], taus[k], c_exp, cs, sn, kre, kim);
    double v = c < N ? kre : kim;
    if (scaled) v *= 1.0 / yerr[c];
    s.Kf[kf_index(k, c - c_lo, KC)] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
  if ((tid & 31) == 0) red[tid >> 5] = csum;
  __syncthreads();
  double tot = 0.0;
  for (int i = 0; i < kWarps; ++i) tot += red[i];
  s.llconst = tot;
  __syncthreads();
}

// One work item: row tile r x column tiles [ntl0, ntl0+ntiles) (local indices), all KC chunks.
// c[][] must be initialised by the caller; on return c = accumulators.
__device__ __forceinline__ void decomp_rc_item(const DecompRCSmem& s, const DecompRCShape& sh,
                                               const double* __restrict__ prop, int ndim, int r, int ntl0, int ntiles,
                                               int lane, double (&c)[kRcMaxTiles][4], double& R0a, double& R0b,
                                               bool init_from_data) {
  const int g = lane >> 2, t = lane & 3;
  const double* q0 = prop + (size_t)(r * 16 + g) * ndim;
  const double* q1 = q0 + 8 * ndim;
  R0a = q0[0];
  R0b = q1[0];
  double a1[4];
  a1[0] = (t < sh.D) ? R0a * q0[1 + t] : 0.0;
  a1[1] = (t < sh.D) ? R0b * q1[1 + t] : 0.0;
  a1[2] = (t + 4 < sh.D) ? R0a * q0[5 + t] : 0.0;
  a1[3] = (t + 4 < sh.D) ? R0b * q1[5 + t] : 0.0;
#pragma unroll
  for (int i = 0; i < kRcMaxTiles; ++i) {
    c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    if (init_from_data && i < ntiles) {
      const int col = (sh.nt_lo + ntl0 + i) * 8 + 2 * t;
      const double2 ys = *reinterpret_cast<const double2*>(s.ycol + col);
      const double2 ds = *reinterpret_cast<const double2*>(s.isig + col);
      c[i][0] = fma(-R0a, ds.x, ys.x);
      c[i][1] = fma(-R0a, ds.y, ys.y);
      c[i][2] = fma(-R0b, ds.x, ys.x);
      c[i][3] = fma(-R0b, ds.y, ys.y);
    }
  }
  const double2* kf = reinterpret_cast<const double2*>(s.Kf) + lane;
  for (int kc = 0; kc < sh.KC; ++kc) {
    double A[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * kc + h;
      double b1[2] = {s.L1[(j * 2 + 0) * 32 + lane], s.L1[(j * 2 + 1) * 32 + lane]};
      double m[4] = {0.0, 0.0, 0.0, 0.0};
      dmma_16x8x8(m, a1, b1);
      A[4 * h + 0] = m[0];
      A[4 * h + 1] = m[2];
      A[4 * h + 2] = m[1];
      A[4 * h + 3] = m[3];
    }
#pragma unroll
    for (int i = 0; i < kRcMaxTiles; ++i) {
      if (i < ntiles) {
        const size_t o = ((size_t)(ntl0 + i) * sh.KC + kc) * 64;
        const double2 b01 = kf[o];
        const double2 b23 = kf[o + 32];
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
        dmma_16x8x16(c[i], A, b);
      }
    }
  }
}

// chi[row] for rows [0,nrows).  Block-level; contains __syncthreads and, for CS > 1, one cluster
// barrier + DSMEM reads.  All CTAs of the cluster must call it with identical arguments.
__device__ inline void decomp_rc_eval_chi(DecompRCSmem& s, const DecompRCShape& sh, const double* __restrict__ prop,
                                          int ndim, int nrows, int rows_pad, double* chi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int RT = (nrows + 15) >> 4;
  for (int item = warp; item < RT * sh.NGC; item += kWarps) {
    const int r = item % RT, cgi = item / RT;
    const int ntl0 = cgi * sh.TPG;
    int ntiles = sh.ntc - ntl0;
    if (ntiles > sh.TPG) ntiles = sh.TPG;
    double chi0 = 0.0, chi1 = 0.0;
    if (ntiles > 0) {
      double c[kRcMaxTiles][4];
      double R0a, R0b;
      decomp_rc_item(s, sh, prop, ndim, r, ntl0, ntiles, lane, c, R0a, R0b, true);
#pragma unroll
      for (int i = 0; i < kRcMaxTiles; ++i) {
        if (i < ntiles) {
          chi0 = fma(c[i][0], c[i][0], chi0);
          chi0 = fma(c[i][1], c[i][1], chi0);
          chi1 = fma(c[i][2], c[i][2], chi1);
          chi1 = fma(c[i][3], c[i][3], chi1);
        }
      }
    }
    chi0 += __shfl_xor_sync(0xffffffffu, chi0, 1);
    chi0 += __shfl_xor_sync(0xffffffffu, chi0, 2);
    chi1 += __shfl_xor_sync(0xffffffffu, chi1, 1);
    chi1 += __shfl_xor_sync(0xffffffffu, chi1, 2);
    if (t == 0) {
      double* dst = s.part + (size_t)cgi * rows_pad;
      dst[r * 16 + g] = chi0;
      dst[r * 16 + g + 8] = chi1;
    }
  }
  __syncthreads();
  double* mine = s.xsum + (size_t)s.parity * rows_pad;
  for (int p = threadIdx.x; p < RT * 16; p += kThreads) {
    double acc = 0.0;
    for (int cgi = 0; cgi < sh.NGC; ++cgi) acc += s.part[(size_t)cgi * rows_pad + p];
    if (sh.CS == 1) chi[p] = acc; else mine[p] = acc;
  }
  if (sh.CS > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                    // every CTA's partial sums are visible
    for (int p = threadIdx.x; p < RT * 16; p += kThreads) {
      double acc = 0.0;
      for (int rk = 0; rk < sh.CS; ++rk) {             // rank order: identical bits in every CTA
        const double* peer = cluster.map_shared_rank(mine, rk);
        acc += peer[p];
      }
      chi[p] = acc;
    }
    s.parity ^= 1;   // the peer may still be reading this buffer until the NEXT cluster barrier
  }
}

// Forward only: each CTA writes its own columns of Z[row][2][N].
__device__ inline void decomp_rc_eval_Z(const DecompRCSmem& s, const DecompRCShape& sh, const double* __restrict__ prop,
                                        int ndim, int nrows, double* __restrict__ Zout) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int RT = (nrows + 15) >> 4;
  for (int item = warp; item < RT * sh.NGC; item += kWarps) {
    const int r = item % RT, cgi = item / RT;
    const int ntl0 = cgi * sh.TPG;
    int ntiles = sh.ntc - ntl0;
    if (ntiles > sh.TPG) ntiles = sh.TPG;
    if (ntiles <= 0) continue;
    double c[kRcMaxTiles][4];
    double R0a, R0b;
    decomp_rc_item(s, sh, prop, ndim, r, ntl0, ntiles, lane, c, R0a, R0b, false);
#pragma unroll
    for (int i = 0; i < kRcMaxTiles; ++i) {
      if (i < ntiles) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = (sh.nt_lo + ntl0 + i) * 8 + 2 * t + e;
          if (col < 2 * sh.N) {
            const double d = (col < sh.N) ? 1.0 : 0.0;
            const int row0 = r * 16 + g, row1 = row0 + 8;
            if (row0 < nrows) Zout[(size_t)row0 * 2 * sh.N + col] = R0a * d - c[i][e];
            if (row1 < nrows) Zout[(size_t)row1 * 2 * sh.N + col] = R0b * d - c[i][2 + e];
          }
        }
      }
    }
  }
}

}  // namespace bisip
