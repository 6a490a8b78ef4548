/*
 * ozaki_split.cuh — exponents, guard and error-free digit split of one K chunk for ozaki_gemm.cuh.
 *
 * For a row i of A (a column j of B) inside the chunk: e = 1 + floor(log2(max |x|)), so |x| * 2^-e < 1;
 *     q = rint(x * 2^(54 - e))            (|q| <= 2^54: the value rounded to 54 bits below the row/column scale)
 *     q = sum_{t=1..7} d_t * 256^(7-t)    with BALANCED digits d_t in [-128, 127], carried from the least significant end
 * Digits are stored as int8 in the tensor core's own tile order (ozaki_gemm.cuh): store[row tile][k step][digit][4 KiB],
 * B transposed (its "rows" are output columns), k padded to a multiple of 128 with zero digits.
 * HBM-bound streaming kernels: 8 bytes read, 7 bytes written per element; the exponent kernels read the chunk once more.
 *
 * Guard (one int per launch, read by both GEMM kernels): the tcgen05 path only takes a chunk when
 *   - every input is finite (an Inf/NaN must propagate as in FP64: native kernel),
 *   - eA[i] + eB[j] stays inside [EXP_SUM_MIN, EXP_SUM_MAX] for every pair, so that no scale factor or scaled
 *     partial sum leaves the normal FP64 range (underflow / overflow then behave as in FP64: native kernel),
 *   - the nonzero entries of every row / column span at most 2^MAX_SPREAD (an entry 2^-56 below its row maximum
 *     would lose all its bits; beyond 2^40 the native kernel takes over and the componentwise FP64 bound holds).
 * Otherwise the guard is set and the DMMA kernel computes the chunk (phpc_launch_ozaki).
 *
 * The split bodies are __host__ __device__ so the very same lines run on the CPU in tests/test_ozaki_split_host.py
 * (tests/csrc/oz_host_probe.cu) against the integer model in oracle/ozaki_model.py.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ozaki_gemm.cuh"

namespace phpc {
namespace oz {

constexpr int GUARD_NONFINITE = 1, GUARD_RANGE = 2, GUARD_SPREAD = 4;
constexpr int EXP_NONE = 2147483647; /* "no nonzero entry seen" for the minimum exponent */

__device__ __forceinline__ int exp_above(double x) { /* smallest e with |x| < 2^e; ZERO_EXP for 0 */
  const int hi = __double2hiint(fabs(x));
  const int lo = __double2loint(x);
  if ((hi | lo) == 0) return ZERO_EXP;
  const int biased = hi >> 20;
  if (biased == 0x7ff) return NONFINITE_EXP;       /* Inf / NaN */
  return biased == 0 ? -1022 : biased - 1023 + 1; /* denormals share the smallest normal exponent */
}

/* e[0 .. n) = ZERO_EXP (maxima), e[n .. 2n) = EXP_NONE (minima), *guard = 0 */
__global__ void exp_init_kernel(int *e, int n, int *guard) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += gridDim.x * blockDim.x) e[i] = i < n ? ZERO_EXP : EXP_NONE;
  if (blockIdx.x == 0 && threadIdx.x == 0) *guard = 0;
}

/* eA[i] / eminA[i] = max / min exponent over the k columns of row i.  One warp per (row, 1024-column segment). */
__global__ void row_exp_kernel(const double *__restrict__ A, long long lda, int m, int k, int *__restrict__ eA, int *__restrict__ eminA,
                               int *__restrict__ guard) {
  const int warps_per_block = blockDim.x >> 5;
  const int segs = (k + 1023) / 1024;
  const long long unit = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (unit >= (long long)m * segs) return;
  const int row = (int)(unit / segs), seg = (int)(unit % segs);
  const int lane = threadIdx.x & 31;
  const double *p = A + (long long)row * lda;
  int e = ZERO_EXP, emin = EXP_NONE;
  const int k_end = min(k, (seg + 1) * 1024);
  for (int c = seg * 1024 + lane; c < k_end; c += 32) {
    const int ex = exp_above(p[c]);
    e = max(e, ex);
    if (ex != ZERO_EXP) emin = min(emin, ex);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
    emin = min(emin, __shfl_xor_sync(0xffffffffu, emin, o));
  }
  if (lane == 0 && e != ZERO_EXP) {
    atomicMax(eA + row, e);
    atomicMin(eminA + row, emin);
    if (e == NONFINITE_EXP) atomicOr(guard, GUARD_NONFINITE);
  }
}

/* eB[j] / eminB[j] over the k rows of column j.  Thread = column, block = 256 columns x 64-row band. */
__global__ void col_exp_kernel(const double *__restrict__ B, long long ldb, int k, int n, int *__restrict__ eB, int *__restrict__ eminB,
                               int *__restrict__ guard) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  const int r0 = blockIdx.y * 64, r1 = min(k, r0 + 64);
  int e = ZERO_EXP, emin = EXP_NONE;
  for (int r = r0; r < r1; ++r) {
    const int ex = exp_above(B[(long long)r * ldb + col]);
    e = max(e, ex);
    if (ex != ZERO_EXP) emin = min(emin, ex);
  }
  if (e != ZERO_EXP) {
    atomicMax(eB + col, e);
    atomicMin(eminB + col, emin);
    if (e == NONFINITE_EXP) atomicOr(guard, GUARD_NONFINITE);
  }
}

/* one block: range of eA and of eB over the nonzero rows / columns, and the largest spread -> guard bits */
__global__ void guard_kernel(const int *__restrict__ eA, const int *__restrict__ eminA, int m, const int *__restrict__ eB,
                             const int *__restrict__ eminB, int n, int *__restrict__ guard) {
  __shared__ int s_max[2], s_min[2], s_spread;
  if (threadIdx.x == 0) {
    s_max[0] = s_max[1] = ZERO_EXP;
    s_min[0] = s_min[1] = EXP_NONE;
    s_spread = 0;
  }
  __syncthreads();
  for (int which = 0; which < 2; ++which) {
    const int *e = which ? eB : eA, *em = which ? eminB : eminA;
    const int cnt = which ? n : m;
    int mx = ZERO_EXP, mn = EXP_NONE, sp = 0;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int v = e[i];
      if (v == NONFINITE_EXP) atomicOr(guard, GUARD_NONFINITE); /* also set by the exponent kernels; exponents of a cached B chunk only pass here */
      if (v == ZERO_EXP || v == NONFINITE_EXP) continue;
      mx = max(mx, v);
      mn = min(mn, v);
      sp = max(sp, v - em[i]);
    }
    atomicMax(&s_max[which], mx);
    atomicMin(&s_min[which], mn);
    atomicMax(&s_spread, sp);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int g = 0;
    if (s_max[0] != ZERO_EXP && s_max[1] != ZERO_EXP) { /* otherwise A or B is all zero: nothing is added, any scale is fine */
      if (s_max[0] + s_max[1] > EXP_SUM_MAX || s_min[0] + s_min[1] < EXP_SUM_MIN) g |= GUARD_RANGE;
    }
    if (s_spread > MAX_SPREAD) g |= GUARD_SPREAD;
    if (g) atomicOr(guard, g);
  }
}

/* ---- digits ---- */

/* balanced base-256 digits of q = rint(x * 2^(BAL_BITS - e)), most significant first */
__host__ __device__ __forceinline__ void balanced_digits(double x, int e, int8_t *out /* [S] */) {
#ifdef __CUDA_ARCH__
  long long q = (e == ZERO_EXP || e == NONFINITE_EXP) ? 0ll : __double2ll_rn(scalbn(x, BAL_BITS - e));
#else
  long long q = (e == ZERO_EXP || e == NONFINITE_EXP) ? 0ll : llrint(scalbn(x, BAL_BITS - e));
#endif
#pragma unroll
  for (int i = S - 1; i >= 0; --i) {
    const long long d = ((q + 128) & 255) - 128;
    q = (q - d) >> 8;
    out[i] = (int8_t)d;
  }
}

/* byte offset of (row, global k byte, digit t) in a tiled digit store: store[row tile][k step][digit][4 KiB] */
__host__ __device__ __forceinline__ size_t store_offset(int row, int kbyte, int t, int ksteps) {
  const int tile = row >> 7, r = row & 127, ks = kbyte >> 5, kb = kbyte & 31;
  return (((size_t)tile * ksteps + ks) * S + t) * (size_t)SLOT_BYTES + tile_offset(r, kb);
}

/* A: work item = 16 consecutive k of one padded row (m_pad = rows of the store: a multiple of 128); consecutive items walk
 * down the rows of one k chunk: the 8 rows of a core matrix are 128 contiguous bytes */
__host__ __device__ __forceinline__ void split_a_body(long long idx, const double *__restrict__ A, long long lda, int m, int m_pad, int k, int kp,
                                                      const int *__restrict__ eA, int8_t *__restrict__ TA) {
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
    balanced_digits((row < m && c < k) ? A[(long long)row * lda + c] : 0.0, e, dg);
#pragma unroll
    for (int t = 0; t < S; ++t) out.b[t][j] = dg[t];
  }
#pragma unroll
  for (int t = 0; t < S; ++t) *reinterpret_cast<int4 *>(TA + store_offset(row, c0, t, kp / 32)) = out.v[t];
}

/* B (transposed): work item = the 32 k of one k step of one padded column; a warp = 32 adjacent columns */
__host__ __device__ __forceinline__ void split_b_body(int col, int ks, const double *__restrict__ B, long long ldb, int k, int n, int n_pad, int kp,
                                                      const int *__restrict__ eB, int8_t *__restrict__ TB) {
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
      balanced_digits((col < n && row < k) ? B[(long long)row * ldb + col] : 0.0, e, dg);
#pragma unroll
      for (int t = 0; t < S; ++t) out.b[t][j] = dg[t];
    }
#pragma unroll
    for (int t = 0; t < S; ++t) *reinterpret_cast<int4 *>(TB + store_offset(col, k0 + half16 * 16, t, kp / 32)) = out.v[t];
  }
}

__global__ void split_a_kernel(const double *__restrict__ A, long long lda, int m, int m_pad, int k, int kp, const int *__restrict__ eA,
                               int8_t *__restrict__ TA, const int *__restrict__ guard) {
  if (*guard != 0) return; /* the native kernel takes this chunk: no digits needed */
  split_a_body((long long)blockIdx.x * blockDim.x + threadIdx.x, A, lda, m, m_pad, k, kp, eA, TA);
}

__global__ void split_b_kernel(const double *__restrict__ B, long long ldb, int k, int n, int n_pad, int kp, const int *__restrict__ eB,
                               int8_t *__restrict__ TB, const int *__restrict__ guard /* may be NULL: split unconditionally */) {
  if (guard && *guard != 0) return;
  split_b_body(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, B, ldb, k, n, n_pad, kp, eB, TB);
}

}  // namespace oz
}  // namespace phpc
