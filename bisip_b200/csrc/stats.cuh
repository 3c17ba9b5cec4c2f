// Column statistics of data[B][n][ncol]: exact order statistics (NumPy 'linear' percentile),
// mean and population std.  Replaces np.percentile / np.mean / np.std over the flat chain
// (reference utils.py:35, :53, :69, :85).
//
// Kernel 1 (transpose_keys): rows are read coalesced and scattered into a column-major array
//   of order-preserving uint64 keys, keys[b][col][n].
// Kernel 2 (column_select): one CTA per (col, b).  Multi-rank MSB radix select: 8 passes of
//   8-bit digits over the contiguous key column; every requested rank carries its own
//   (prefix, remaining-rank) pair and its own 256-bin shared-memory histogram, so all ranks
//   are resolved in the same 8 sweeps.  Mean / std use fixed-order block reductions
//   (two-pass variance, like NumPy).
#pragma once
#include "common.cuh"

namespace bisip {

constexpr int kMaxPct = 16;              // percentiles per call
constexpr int kMaxRanks = 2 * kMaxPct;   // each needs the order statistics lo and lo+1

struct StatsParams {
  const double* data;
  unsigned long long* keys;   // workspace [B][ncol][n]
  long long n;
  int ncol, B, npct;
  long long lo[kMaxPct];
  double gamma[kMaxPct];
  double* pct_out;
  double* mean_out;
  double* std_out;
};

__device__ __forceinline__ unsigned long long f64_key(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// grid (row_chunks, B); each CTA transposes a [64 rows][ncol] slab through registers.
__global__ void __launch_bounds__(kThreads) transpose_keys_kernel(const StatsParams P) {
  const int b = blockIdx.y;
  const long long n = P.n;
  const int ncol = P.ncol;
  const double* src = P.data + (size_t)b * n * ncol;
  unsigned long long* dst = P.keys + (size_t)b * n * ncol;
  const long long chunk = 4096;   // rows per CTA
  const long long r0 = (long long)blockIdx.x * chunk;
  const long long r1 = min(n, r0 + chunk);
  // thread -> (row, col) with col fastest on the read side; writes are strided by column
  // but each warp writes runs of consecutive rows for the same column after the swap below.
  for (int c = 0; c < ncol; ++c) {
    for (long long r = r0 + threadIdx.x; r < r1; r += kThreads)
      dst[(size_t)c * n + r] = f64_key(src[(size_t)r * ncol + c]);
  }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < kWarps; ++i) t += red[i];
  return t;
}

// grid (ncol, B)
__global__ void __launch_bounds__(kThreads) column_select_kernel(const StatsParams P) {
  __shared__ unsigned int hist[kMaxRanks][256];
  __shared__ unsigned long long prefix[kMaxRanks];
  __shared__ long long krem[kMaxRanks];
  __shared__ double red[kWarps];
  const int col = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const long long n = P.n;
  const unsigned long long* keys = P.keys + ((size_t)b * P.ncol + col) * n;
  const int R = 2 * P.npct;

  if (tid < R) {
    const int q = tid >> 1;
    long long k = P.lo[q] + (tid & 1);
    if (k > n - 1) k = n - 1;
    if (k < 0) k = 0;
    krem[tid] = k;
    prefix[tid] = 0ull;
  }
  // ---- mean / std (two-pass) -----------------------------------------------------------
  if (P.mean_out != nullptr || P.std_out != nullptr) {
    double s = 0.0;
    for (long long i = tid; i < n; i += kThreads) s += key_f64(keys[i]);
    const double mean = block_sum(s, red) / (double)n;
    double v = 0.0;
    for (long long i = tid; i < n; i += kThreads) {
      const double d = key_f64(keys[i]) - mean;
      v = fma(d, d, v);
    }
    const double var = block_sum(v, red) / (double)n;
    if (tid == 0) {
      if (P.mean_out) P.mean_out[(size_t)b * P.ncol + col] = mean;
      if (P.std_out) P.std_out[(size_t)b * P.ncol + col] = sqrt(var);
    }
  }
  // ---- multi-rank radix select ---------------------------------------------------------------
  // dprefix[0..ndist): the distinct prefixes among the R ranks; owner[r]: which one rank r follows
  __shared__ unsigned long long dprefix[kMaxRanks];
  __shared__ int owner[kMaxRanks];
  __shared__ int ndist_s;
  if (tid == 0) { ndist_s = 1; dprefix[0] = 0ull; }
  if (tid < R) owner[tid] = 0;
  __syncthreads();
  for (int pass = 0; pass < 8; ++pass) {
    const int ndist = ndist_s;
    for (int i = tid; i < ndist * 256; i += kThreads) (&hist[0][0])[i] = 0u;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    for (long long i0 = 0; i0 < n; i0 += kThreads) {
      const long long i = i0 + tid;
      const bool valid = i < n;
      const unsigned long long key = valid ? keys[i] : 0ull;
      const unsigned int digit = (unsigned int)(key >> shift) & 255u;
      const unsigned long long hi = pass == 0 ? 0ull : (key >> (shift + 8));
      for (int u = 0; u < ndist; ++u) {
        const bool m = valid && (hi == dprefix[u]);
        const unsigned int mask = __ballot_sync(0xffffffffu, m);
        if (m) {
          const unsigned int peers = __match_any_sync(mask, digit);
          if ((int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[u][digit], __popc(peers));
        }
      }
    }
    __syncthreads();
    if (tid < R) {
      const int src = owner[tid];
      long long k = krem[tid];
      unsigned int d = 0;
      for (; d < 255u; ++d) {
        const unsigned int c = hist[src][d];
        if (k < (long long)c) break;
        k -= c;
      }
      krem[tid] = k;
      prefix[tid] = (prefix[tid] << 8) | d;
    }
    __syncthreads();
    if (tid == 0) {
      int nd = 0;
      for (int r = 0; r < R; ++r) {
        int o = -1;
        for (int u = 0; u < nd; ++u) if (dprefix[u] == prefix[r]) o = u;
        if (o < 0) { o = nd; dprefix[nd++] = prefix[r]; }
        owner[r] = o;
      }
      ndist_s = nd;
    }
    __syncthreads();
  }
  // ---- NumPy _lerp: a + (b-a)*t, and b - (b-a)*(1-t) where t >= 0.5 (no FMA contraction) ------
  if (tid < P.npct) {
    const double a = key_f64(prefix[2 * tid]), bb = key_f64(prefix[2 * tid + 1]);
    const double t = P.gamma[tid];
    const double diff = __dsub_rn(bb, a);
    double v = __dadd_rn(a, __dmul_rn(diff, t));
    if (t >= 0.5) v = __dsub_rn(bb, __dmul_rn(diff, __dsub_rn(1.0, t)));
    P.pct_out[((size_t)b * P.npct + tid) * P.ncol + col] = v;
  }
}

}  // namespace bisip
