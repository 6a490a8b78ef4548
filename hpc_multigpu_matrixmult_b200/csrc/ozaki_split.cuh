/*
 * ozaki_split.cuh — error-free splitting of FP64 operands into signed 7-bit digits.
 *
 * For a row i of A (a column j of B): e = 1 + floor(log2(max |x|)) so that |x| * 2^-e < 1, then
 *     r_0 = x * 2^-e;   d_t = trunc(r_{t-1} * 128) in [-127, 127];   r_t = r_{t-1} * 128 - d_t
 * Every step is exact in FP64 (scaling by powers of two, subtracting the integer part), so
 *     x = 2^e * ( sum_{t=1..S} d_t * 2^(-7t) + r_S * 2^(-7S) ),   |r_S| < 1,
 * i.e. S digits carry the top 7*S bits below the row/column maximum.  Digits are stored as
 * K-major int8 matrices, one per t, back to back:
 *     SA[t][i][k]  (m rows, row pitch kp bytes)        SB[u][j][k]  (n rows: B is transposed)
 * kp = K rounded up to 128 with zero digits, which is what the UMMA K-major tiles want.
 * HBM-bound streaming kernels: 8 bytes read, S bytes written per element.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ozaki_gemm.cuh"

namespace phpc {
namespace oz {

__device__ __forceinline__ int exp_above(double x) { /* smallest e with |x| < 2^e; ZERO_EXP for 0 */
  const int hi = __double2hiint(fabs(x));
  const int lo = __double2loint(x);
  if ((hi | lo) == 0) return ZERO_EXP;
  const int biased = hi >> 20;
  if (biased == 0x7ff) return NONFINITE_EXP;       /* Inf / NaN poisons the row / column */
  return biased == 0 ? -1022 : biased - 1023 + 1; /* denormals share the smallest normal exponent */
}

__global__ void exp_init_kernel(int *e, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) e[i] = ZERO_EXP;
}

/* eA[i] = max over the k columns of row i.  One warp per (row, 1024-column segment). */
__global__ void row_exp_kernel(const double *__restrict__ A, long long lda, int m, int k, int *__restrict__ eA) {
  const int warps_per_block = blockDim.x >> 5;
  const int segs = (k + 1023) / 1024;
  const long long unit = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (unit >= (long long)m * segs) return;
  const int row = (int)(unit / segs), seg = (int)(unit % segs);
  const int lane = threadIdx.x & 31;
  const double *p = A + (long long)row * lda;
  int e = ZERO_EXP;
  const int k_end = min(k, (seg + 1) * 1024);
  for (int c = seg * 1024 + lane; c < k_end; c += 32) e = max(e, exp_above(p[c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
  if (lane == 0 && e != ZERO_EXP) atomicMax(eA + row, e);
}

/* eB[j] = max over the k rows of column j.  Thread = column, block = 256 columns x 64-row band. */
__global__ void col_exp_kernel(const double *__restrict__ B, long long ldb, int k, int n, int *__restrict__ eB) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  const int r0 = blockIdx.y * 64, r1 = min(k, r0 + 64);
  int e = ZERO_EXP;
  for (int r = r0; r < r1; ++r) e = max(e, exp_above(B[(long long)r * ldb + col]));
  if (e != ZERO_EXP) atomicMax(eB + col, e);
}

/* ---- tiled digit stores: store[row tile][k step][digit][4 KiB canonical tile] ---- */

/* A: thread = 16 consecutive k of one (padded) row = one 16-byte chunk of a core matrix per digit */
__global__ void split_a_tiled_kernel(const double *__restrict__ A, long long lda, int m, int k, int kp, const int *__restrict__ eA,
                                     int8_t *__restrict__ TA, int S) {
  const int chunks = kp / 16;
  const int m_pad = (m + 127) / 128 * 128;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)m_pad * chunks) return;
  /* consecutive threads walk down the rows of one k chunk: the 8 rows of a core matrix are 128 contiguous bytes */
  const int chunk = (int)(idx / m_pad), row = (int)(idx % m_pad);
  const int c0 = chunk * 16;
  const int e = row < m ? eA[row] : ZERO_EXP;
  double r[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int c = c0 + j;
    r[j] = (row < m && c < k && e != ZERO_EXP && e != NONFINITE_EXP) ? scalbn(A[(long long)row * lda + c], -e) : 0.0;
  }
  const int ksteps = kp / 32;
  const size_t base = (((size_t)(row >> 7) * ksteps + (c0 >> 5)) * S) * 4096 + tile_offset(row & 127, c0 & 31);
  for (int t = 0; t < S; ++t) {
    union {
      int8_t b[16];
      int4 v;
    } out;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const double s = r[j] * 128.0;
      const int d = (int)s;
      r[j] = s - (double)d;
      out.b[j] = (int8_t)d;
    }
    *reinterpret_cast<int4 *>(TA + base + (size_t)t * 4096) = out.v;
  }
}

/* B (transposed): thread = 32 consecutive k of one (padded) column; warp = 32 adjacent columns */
__global__ void split_b_tiled_kernel(const double *__restrict__ B, long long ldb, int k, int n, int kp, const int *__restrict__ eB,
                                     int8_t *__restrict__ TB, int S) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_pad = (n + 127) / 128 * 128;
  if (col >= n_pad) return;
  const int ks = blockIdx.y; /* k step of 32 */
  const int k0 = ks * 32;
  const int e = col < n ? eB[col] : ZERO_EXP;
  double r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int row = k0 + j;
    r[j] = (col < n && row < k && e != ZERO_EXP && e != NONFINITE_EXP) ? scalbn(B[(long long)row * ldb + col], -e) : 0.0;
  }
  const int ksteps = kp / 32;
  const size_t base = (((size_t)(col >> 7) * ksteps + ks) * S) * 4096 + tile_offset(col & 127, 0);
  for (int t = 0; t < S; ++t) {
    union {
      int8_t b[32];
      int4 v[2];
    } out;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const double s = r[j] * 128.0;
      const int d = (int)s;
      r[j] = s - (double)d;
      out.b[j] = (int8_t)d;
    }
    int8_t *dst = TB + base + (size_t)t * 4096;
    *reinterpret_cast<int4 *>(dst) = out.v[0];       /* k bytes 0..15  */
    *reinterpret_cast<int4 *>(dst + 128) = out.v[1]; /* k bytes 16..31: next core matrix along k */
  }
}


/* ---- EXPERIMENTAL balanced base-256 digits (see BAL_BITS in ozaki_gemm.cuh) ----
 * q = rint(x * 2^(BAL_BITS - e)) (|q| <= 2^54), written in base 256 with digits in [-128, 127] by carrying
 * from the least significant end; digit slot 0 is the most significant.  Same tiled store layout. */
__device__ __forceinline__ void balanced_digits(double x, int e, int S, int8_t *out /* [S], most significant first */) {
  long long q = (e == ZERO_EXP || e == NONFINITE_EXP) ? 0ll : __double2ll_rn(scalbn(x, BAL_BITS - e));
  for (int i = S - 1; i >= 0; --i) {
    const long long d = ((q + 128) & 255) - 128;
    q = (q - d) >> 8;
    out[i] = (int8_t)d;
  }
}

__global__ void split_a_tiled_balanced_kernel(const double *__restrict__ A, long long lda, int m, int k, int kp, const int *__restrict__ eA,
                                              int8_t *__restrict__ TA, int S) {
  const int chunks = kp / 16;
  const int m_pad = (m + 127) / 128 * 128;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)m_pad * chunks) return;
  const int chunk = (int)(idx / m_pad), row = (int)(idx % m_pad);
  const int c0 = chunk * 16;
  const int e = row < m ? eA[row] : ZERO_EXP;
  union {
    int8_t b[MAX_SLICES][16];
    int4 v[MAX_SLICES];
  } out;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int c = c0 + j;
    int8_t dg[MAX_SLICES];
    balanced_digits((row < m && c < k) ? A[(long long)row * lda + c] : 0.0, e, S, dg);
    for (int t = 0; t < S; ++t) out.b[t][j] = dg[t];
  }
  const int ksteps = kp / 32;
  const size_t base = (((size_t)(row >> 7) * ksteps + (c0 >> 5)) * S) * 4096 + tile_offset(row & 127, c0 & 31);
  for (int t = 0; t < S; ++t) *reinterpret_cast<int4 *>(TA + base + (size_t)t * 4096) = out.v[t];
}

__global__ void split_b_tiled_balanced_kernel(const double *__restrict__ B, long long ldb, int k, int n, int kp, const int *__restrict__ eB,
                                              int8_t *__restrict__ TB, int S) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_pad = (n + 127) / 128 * 128;
  if (col >= n_pad) return;
  const int ks = blockIdx.y;
  const int k0 = ks * 32;
  const int e = col < n ? eB[col] : ZERO_EXP;
  const int ksteps = kp / 32;
  int8_t *dst0 = TB + (((size_t)(col >> 7) * ksteps + ks) * S) * 4096 + tile_offset(col & 127, 0);
  for (int half = 0; half < 2; ++half) { /* 16 k bytes = one core-matrix row per half */
    union {
      int8_t b[MAX_SLICES][16];
      int4 v[MAX_SLICES];
    } out;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int row = k0 + half * 16 + j;
      int8_t dg[MAX_SLICES];
      balanced_digits((col < n && row < k) ? B[(long long)row * ldb + col] : 0.0, e, S, dg);
      for (int t = 0; t < S; ++t) out.b[t][j] = dg[t];
    }
    for (int t = 0; t < S; ++t) *reinterpret_cast<int4 *>(dst0 + (size_t)t * 4096 + half * 128) = out.v[t];
  }
}

}  // namespace oz
}  // namespace phpc
