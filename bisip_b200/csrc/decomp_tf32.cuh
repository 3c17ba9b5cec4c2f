// Reduced-precision variants of the decomposition contraction (north_star: "a TF32/3xTF32 variant
// compared against" the FP64 DMMA path; BASELINE config 4 is the tolerance study).
//
//   stage 1 (coefficients -> chargeability M over the tau grid) stays on FP64 DMMA tiles: it is 6 % of
//           the flops and badly conditioned (powers of log10(tau) up to 6^P with cancelling
//           coefficients), so rounding it to a 10-bit mantissa would destroy M itself;
//   stage 2 (M x K, 92 % of the flops) runs on TF32 tensor tiles (mma.sync.m16n8k8.tf32, FP32
//           accumulate): PREC = 1 plain TF32 operands, PREC = 3 the error-compensated split
//           x = hi + lo (both TF32):  A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  ("3xTF32").
// Same work decomposition as decomp_rc.cuh (stage 1 recomputed per 16-tau chunk, <= 8 column tiles
// resident, columns optionally split over a CTA cluster), so every tau-grid size is covered.
// The accumulators start at (y - R0*delta)/sigma rounded to FP32 and end as the weighted residual;
// chi^2 is accumulated in FP64.
#pragma once
#include "decomp_rc.cuh"

namespace bisip {

__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
  uint32_t r;
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(double x, uint32_t& hi, uint32_t& lo) {
  const float xf = (float)x;
  hi = f32_to_tf32(xf);
  lo = f32_to_tf32(xf - __uint_as_float(hi));
}
// A (16x8, row): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B (8x8, col): b0 (k=t, n=g) b1 (k=t+4, n=g)
// C (16x8 f32): c0,c1 (g, 2t..2t+1)  c2,c3 (g+8, 2t..2t+1)
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct DecompTF32Smem {
  uint32_t* Khi;  // [NTC][KC][2][32][2]  TF32 bits of K/sigma: (kc, h) = 8-tau tile, lane, {slot t, slot t+4}
  uint32_t* Klo;  // same layout, residual plane (3xTF32 only)
  double* L1;
  double* ycol;
  double* isig;
  double* part;
  double* xsum;
  double llconst;
  int parity;
};

__host__ __device__ inline size_t decomp_tf32_smem_doubles(const DecompRCShape& sh, int rows_pad, int prec) {
  const size_t planes = prec == 3 ? 2 : 1;
  return planes * sh.kf_doubles() / 2 + sh.l1_doubles() + 2 * sh.col_doubles() + (size_t)(sh.NGC + 2) * rows_pad;
}

template <int PREC>
__device__ inline double* decomp_tf32_carve(DecompTF32Smem& s, double* base, const DecompRCShape& sh, int rows_pad) {
  s.Khi = reinterpret_cast<uint32_t*>(base); base += sh.kf_doubles() / 2;
  s.Klo = reinterpret_cast<uint32_t*>(base); if (PREC == 3) base += sh.kf_doubles() / 2;
  s.L1 = base; base += sh.l1_doubles();
  s.ycol = base; base += sh.col_doubles();
  s.isig = base; base += sh.col_doubles();
  s.part = base; base += (size_t)sh.NGC * rows_pad;
  s.xsum = base; base += (size_t)2 * rows_pad;
  s.parity = 0;
  return base;
}

template <int PREC>
__device__ inline void decomp_tf32_init(DecompTF32Smem& s, const DecompRCShape& sh, double c_exp,
                                        const double* __restrict__ w, const double* __restrict__ taus,
                                        const double* __restrict__ log_taus, const double* __restrict__ y,
                                        const double* __restrict__ yerr, double* red) {
  const int tid = threadIdx.x;
  const int N = sh.N, S = sh.S, KC = sh.KC;
  for (int i = tid; i < (int)sh.kf_doubles(); i += kThreads) {
    s.Khi[i] = 0u;
    if (PREC == 3) s.Klo[i] = 0u;
  }
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
  const int c_lo = sh.nt_lo * 8, ncols = min(2 * N, (sh.nt_lo + sh.ntc) * 8) - c_lo;
  for (int i = tid; i < S * max(ncols, 0); i += kThreads) {
    const int k = i / ncols, c = c_lo + (i - k * ncols);
    const int j = c < N ? c : c - N;
    double kre, kim;
    debye_kernel_term(w[j], taus[k], c_exp, cs, sn, kre, kim);
    double v = c < N ? kre : kim;
    if (scaled) v *= 1.0 / yerr[c];
    // tau k = 16 kc + 8 h + 2 t + e  ->  8-tau tile (kc, h), B slot t + 4 e ; column c_local = 8 ntl + g
    const int kc = k >> 4, h = (k >> 3) & 1, t = (k >> 1) & 3, e = k & 1;
    const int cl = c - c_lo, ntl = cl >> 3, g = cl & 7;
    const int idx = ((((ntl * KC + kc) * 2 + h) * 32 + (g * 4 + t)) << 1) + e;
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    s.Khi[idx] = hi;
    if (PREC == 3) s.Klo[idx] = lo;
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

template <int PREC>
__device__ __forceinline__ void decomp_tf32_item(const DecompTF32Smem& s, const DecompRCShape& sh,
                                                 const double* __restrict__ prop, int ndim, int r, int ntl0, int ntiles,
                                                 int lane, float (&c)[kRcMaxTiles][4], double& R0a, double& R0b,
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
    c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    if (init_from_data && i < ntiles) {
      const int col = (sh.nt_lo + ntl0 + i) * 8 + 2 * t;
      const double2 ys = *reinterpret_cast<const double2*>(s.ycol + col);
      const double2 ds = *reinterpret_cast<const double2*>(s.isig + col);
      c[i][0] = (float)fma(-R0a, ds.x, ys.x);
      c[i][1] = (float)fma(-R0a, ds.y, ys.y);
      c[i][2] = (float)fma(-R0b, ds.x, ys.x);
      c[i][3] = (float)fma(-R0b, ds.y, ys.y);
    }
  }
  const uint2* khi = reinterpret_cast<const uint2*>(s.Khi) + lane;
  const uint2* klo = reinterpret_cast<const uint2*>(s.Klo) + lane;
  for (int kc = 0; kc < sh.KC; ++kc) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * kc + h;
      double b1[2] = {s.L1[(j * 2 + 0) * 32 + lane], s.L1[(j * 2 + 1) * 32 + lane]};
      double m[4] = {0.0, 0.0, 0.0, 0.0};
      dmma_16x8x8(m, a1, b1);                    // FP64 stage 1: M[rows g, g+8][taus 8j+2t, 8j+2t+1]
      uint32_t ahi[4], alo[4];                   // A slots: t <- tau 8j+2t, t+4 <- tau 8j+2t+1
      split_tf32(m[0], ahi[0], alo[0]);
      split_tf32(m[2], ahi[1], alo[1]);
      split_tf32(m[1], ahi[2], alo[2]);
      split_tf32(m[3], ahi[3], alo[3]);
#pragma unroll
      for (int i = 0; i < kRcMaxTiles; ++i) {
        if (i < ntiles) {
          const size_t o = (((size_t)(ntl0 + i) * sh.KC + kc) * 2 + h) * 32;
          const uint2 bh = khi[o];
          if (PREC == 3) {
            const uint2 bl = klo[o];
            mma_tf32_16x8x8(c[i], alo, bh.x, bh.y);
            mma_tf32_16x8x8(c[i], ahi, bl.x, bl.y);
          }
          mma_tf32_16x8x8(c[i], ahi, bh.x, bh.y);
        }
      }
    }
  }
}

template <int PREC>
__device__ inline void decomp_tf32_eval_chi(DecompTF32Smem& s, const DecompRCShape& sh, const double* __restrict__ prop,
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
      float c[kRcMaxTiles][4];
      double R0a, R0b;
      decomp_tf32_item<PREC>(s, sh, prop, ndim, r, ntl0, ntiles, lane, c, R0a, R0b, true);
#pragma unroll
      for (int i = 0; i < kRcMaxTiles; ++i) {
        if (i < ntiles) {
          const double c0 = c[i][0], c1 = c[i][1], c2 = c[i][2], c3 = c[i][3];
          chi0 = fma(c0, c0, chi0);
          chi0 = fma(c1, c1, chi0);
          chi1 = fma(c2, c2, chi1);
          chi1 = fma(c3, c3, chi1);
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
    cluster.sync();
    for (int p = threadIdx.x; p < RT * 16; p += kThreads) {
      double acc = 0.0;
      for (int rk = 0; rk < sh.CS; ++rk) acc += cluster.map_shared_rank(mine, rk)[p];
      chi[p] = acc;
    }
    s.parity ^= 1;
  }
}

template <int PREC>
__device__ inline void decomp_tf32_eval_Z(const DecompTF32Smem& s, const DecompRCShape& sh,
                                          const double* __restrict__ prop, int ndim, int nrows, double* __restrict__ Zout) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int RT = (nrows + 15) >> 4;
  for (int item = warp; item < RT * sh.NGC; item += kWarps) {
    const int r = item % RT, cgi = item / RT;
    const int ntl0 = cgi * sh.TPG;
    int ntiles = sh.ntc - ntl0;
    if (ntiles > sh.TPG) ntiles = sh.TPG;
    if (ntiles <= 0) continue;
    float c[kRcMaxTiles][4];
    double R0a, R0b;
    decomp_tf32_item<PREC>(s, sh, prop, ndim, r, ntl0, ntiles, lane, c, R0a, R0b, false);
#pragma unroll
    for (int i = 0; i < kRcMaxTiles; ++i) {
      if (i < ntiles) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = (sh.nt_lo + ntl0 + i) * 8 + 2 * t + e;
          if (col < 2 * sh.N) {
            const double d = (col < sh.N) ? 1.0 : 0.0;
            const int row0 = r * 16 + g, row1 = row0 + 8;
            if (row0 < nrows) Zout[(size_t)row0 * 2 * sh.N + col] = R0a * d - (double)c[i][e];
            if (row1 < nrows) Zout[(size_t)row1 * 2 * sh.N + col] = R0b * d - (double)c[i][2 + e];
          }
        }
      }
    }
  }
}

}  // namespace bisip
