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


/* ======================================================================================================
 * EXPERIMENTAL split kernels (opt-in through PHPC_OZAKI_DIGITS / PHPC_OZAKI_KERNEL, see phpc_launch_ozaki):
 *   - balanced base-256 digits (BAL_BITS in ozaki_gemm.cuh): 7 digits, 28 digit products instead of 36;
 *   - the B store in "half-major" order for the 2-CTA kernel (ozaki_gemm2.cuh): each CTA of a pair reads
 *     the 64 B^T rows (output columns) of its half of every digit tile as one contiguous range.
 * Their bodies are __host__ __device__ so the very same lines run on the CPU in tests/test_ozaki_split_host.py
 * (tests/csrc/oz_host_probe.cu) against the integer model in oracle/ozaki_model.py.
 * ====================================================================================================== */

/* q = rint(x * 2^(BAL_BITS - e)) (|q| <= 2^54), written in base 256 with digits in [-128, 127] by carrying
 * from the least significant end; digit slot 0 is the most significant. */
__host__ __device__ __forceinline__ void balanced_digits(double x, int e, int S, int8_t *out /* [S], most significant first */) {
#ifdef __CUDA_ARCH__
  long long q = (e == ZERO_EXP || e == NONFINITE_EXP) ? 0ll : __double2ll_rn(scalbn(x, BAL_BITS - e));
#else
  long long q = (e == ZERO_EXP || e == NONFINITE_EXP) ? 0ll : llrint(scalbn(x, BAL_BITS - e));
#endif
  for (int i = S - 1; i >= 0; --i) {
    const long long d = ((q + 128) & 255) - 128;
    q = (q - d) >> 8;
    out[i] = (int8_t)d;
  }
}

/* digits of one value, most significant first: BAL = false is the truncating 7-bit scheme of the kernels above */
template <bool BAL, int S>
__host__ __device__ __forceinline__ void digits_of(double x, int e, int8_t *out) {
  if (BAL) {
    balanced_digits(x, e, S, out);
    return;
  }
  double r = (e == ZERO_EXP || e == NONFINITE_EXP) ? 0.0 : scalbn(x, -e);
#pragma unroll
  for (int t = 0; t < S; ++t) {
    const double s = r * 128.0;
    const int d = (int)s;
    r = s - (double)d;
    out[t] = (int8_t)d;
  }
}

/* byte offset of (row, global k byte, digit t) in a tiled digit store.  halves = 1: store[row tile][k step][digit][4 KiB]
 * (the layout of the kernels above); halves = 2: store[row tile][k step][half][digit][2 KiB], rows 0..63 / 64..127 */
__host__ __device__ __forceinline__ size_t store_offset(int row, int kbyte, int t, int S, int ksteps, int halves) {
  const int tile = row >> 7, r = row & 127, ks = kbyte >> 5, kb = kbyte & 31;
  const int rows_per_half = 128 / halves, h = r / rows_per_half, rh = r % rows_per_half;
  return ((((size_t)tile * ksteps + ks) * halves + h) * S + t) * ((size_t)rows_per_half * 32) + tile_offset(rh, kb);
}

/* A: work item = 16 consecutive k of one padded row (m_pad = rows of the store: a multiple of 128, of 256 for the 2-CTA kernel) */
template <bool BAL, int S>
__host__ __device__ __forceinline__ void split_a_body(long long idx, const double *__restrict__ A, long long lda, int m, int m_pad, int k,
                                                      int kp, const int *__restrict__ eA, int8_t *__restrict__ TA) {
  const int chunks = kp / 16;
  if (idx >= (long long)m_pad * chunks) return;
  const int chunk = (int)(idx / m_pad), row = (int)(idx % m_pad);
  const int c0 = chunk * 16;
  const int e = row < m ? eA[row] : ZERO_EXP;
  union {
    int8_t b[S][16];
    int4 v[S];
  } out;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int c = c0 + j;
    int8_t dg[S];
    digits_of<BAL, S>((row < m && c < k) ? A[(long long)row * lda + c] : 0.0, e, dg);
#pragma unroll
    for (int t = 0; t < S; ++t) out.b[t][j] = dg[t];
  }
#pragma unroll
  for (int t = 0; t < S; ++t) *reinterpret_cast<int4 *>(TA + store_offset(row, c0, t, S, kp / 32, 1)) = out.v[t];
}

/* B (transposed): work item = the 32 k of one k step of one padded column */
template <bool BAL, int S>
__host__ __device__ __forceinline__ void split_b_body(int col, int ks, const double *__restrict__ B, long long ldb, int k, int n, int n_pad,
                                                      int kp, const int *__restrict__ eB, int8_t *__restrict__ TB, int halves) {
  if (col >= n_pad) return;
  const int k0 = ks * 32;
  const int e = col < n ? eB[col] : ZERO_EXP;
  for (int half16 = 0; half16 < 2; ++half16) { /* 16 k bytes = one core-matrix row */
    union {
      int8_t b[S][16];
      int4 v[S];
    } out;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int row = k0 + half16 * 16 + j;
      int8_t dg[S];
      digits_of<BAL, S>((col < n && row < k) ? B[(long long)row * ldb + col] : 0.0, e, dg);
#pragma unroll
      for (int t = 0; t < S; ++t) out.b[t][j] = dg[t];
    }
#pragma unroll
    for (int t = 0; t < S; ++t) *reinterpret_cast<int4 *>(TB + store_offset(col, k0 + half16 * 16, t, S, kp / 32, halves)) = out.v[t];
  }
}

template <bool BAL, int S>
__global__ void split_a_tiled_v2_kernel(const double *__restrict__ A, long long lda, int m, int m_pad, int k, int kp,
                                        const int *__restrict__ eA, int8_t *__restrict__ TA) {
  split_a_body<BAL, S>((long long)blockIdx.x * blockDim.x + threadIdx.x, A, lda, m, m_pad, k, kp, eA, TA);
}

template <bool BAL, int S>
__global__ void split_b_tiled_v2_kernel(const double *__restrict__ B, long long ldb, int k, int n, int n_pad, int kp,
                                        const int *__restrict__ eB, int8_t *__restrict__ TB, int halves) {
  split_b_body<BAL, S>(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, B, ldb, k, n, n_pad, kp, eB, TB, halves);
}

}  // namespace oz
}  // namespace phpc
