/*
 * ozaki_gemm.cuh — FP64 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM).
 *
 * tcgen05.mma has no FP64 kind, so the FP64 product is rebuilt EXACTLY from integer products (the
 * Ozaki scheme): every row of A and every column of B is scaled by a power of two and cut into S
 * signed 7-bit digits (ozaki_split.cuh),
 *     a_ik = 2^eA[i] * sum_t A_t[i][k] * 2^(-7t),   b_kj = 2^eB[j] * sum_u B_u[k][j] * 2^(-7u),
 * the digit matrices are multiplied on the int8 tensor pipe with exact int32 accumulation in TMEM
 * (|A_t.B_u| <= K * 127^2, no rounding at all), and
 *     C[i][j] += 2^(eA[i]+eB[j]) * sum_g 2^(-7g) * P_g[i][j],   P_g = sum_{t+u=g} A_t.B_u
 * is applied in FP64 by the epilogue warps.  Groups with g > S+1 are dropped (their weight is below
 * 2^(-7(S+1)) of the row/column scale), so S(S+1)/2 int8 MMAs stand for one FP64 MMA; S = 8 carries
 * 56 bits.  The reference kernel this replaces is gemm_kernel of src/phpc_gemm.cu:6-57 (same
 * C += A.B contract); the arithmetic differs from it only in the order of exact partial sums.
 *
 * Kernel (one CTA per SM, persistent over 128 x 128 output tiles, static round robin):
 *   K-outer schedule  per 32-byte k step ALL needed digit tiles of A and B are staged once (one
 *               4 KiB slot per digit matrix) and every pair (t,u) of up to four groups is issued from
 *               them, one TMEM accumulator (128 columns) per group = all 512 TMEM columns:
 *                 pass 1  groups S+1 .. S-2  (the 4 least significant; needs every digit)
 *                 pass 2  groups S-3 .. 2    (digits 1..S-4 only; two k steps per stage)
 *   digit stores  written by the split kernels ALREADY in the shared-memory order the tensor core
 *               wants (UMMA canonical K-major, no swizzle: 8-row x 16-byte core matrices; a 128-row x
 *               32-byte tile = 4 KiB, k chunks 128 B apart, 8-row groups 256 B apart), tile after tile:
 *               store[row tile][k step][digit][4 KiB], so a k step of a pass is ONE contiguous global
 *               range per operand.
 *   warp 0      producer: two cp.async.bulk copies per k step into a 3-stage mbarrier ring
 *   warp 1      TMEM allocator + MMA issuer: the whole warp walks warp-uniform, fully unrolled loops
 *               and one elected lane issues tcgen05.mma.kind::i8 128x128x32 / tcgen05.commit (with the
 *               loops inside `if (lane == 0)` every MMA cost 140-180 cycles of register -> uniform
 *               register moves instead of 65; tools/umma_rate.cu, profiles/umma_rate*_r01.jsonl)
 *   warps 2-5   epilogue: tcgen05.ld the four int32 accumulators of a pass, combine them exactly in
 *               FP64 (<= 52 significant bits), transpose through shared memory, one coalesced
 *               read-modify-write of C per pass with 32 loads in flight per lane
 * Earlier variants (pair-outer 128x256 tiles with TMA; K-outer with 16 TMA boxes per step) and what
 * ncu said about them are in profiles/ozaki_experiments_r01.md.
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmma_gemm.cuh" /* mbarrier wrappers, smem_u32, tile_coords */

namespace phpc {
namespace oz {

constexpr int DIGIT_BITS = 7;
/* EXPERIMENTAL (PHPC_OZAKI_DIGITS=balanced, not validated on hardware in round 1): balanced base-256
 * digits in [-128,127] of the value rounded to BAL_BITS bits below its row/column scale; 7 digits,
 * 28 digit products, same accuracy in the integer model (oracle/ozaki_model.py, gemm_balanced). */
constexpr int BAL_BITS = 54;
constexpr int MAX_SLICES = 8;
constexpr int ZERO_EXP = -2147483647 - 1; /* exponent of an all-zero row / column */
constexpr int NONFINITE_EXP = 2147483647; /* the row / column holds an Inf or NaN: its C elements become NaN */

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BKB = 32;                /* bytes of k per step = one int8 MMA (K = 32) */
constexpr int SLOT_BYTES = BM * BKB;   /* one digit tile: 128 rows x 32 B */
constexpr int TILE_BYTES = SLOT_BYTES;
constexpr int MAX_S = MAX_SLICES;
constexpr int STAGE_BYTES = 2 * MAX_S * SLOT_BYTES; /* A digit slots then B digit slots */
constexpr int STAGES = 3;
constexpr int THREADS = 192;           /* warp 0 producer, warp 1 MMA, warps 2-5 epilogue */
constexpr int EPI_WARP_BYTES = 32 * 33 * 8; /* per epilogue warp: 32 x 32 doubles, padded */
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 4 * EPI_WARP_BYTES;
constexpr int GROUPS_PER_PASS = 4;
constexpr int TMEM_COLS = GROUPS_PER_PASS * BN; /* 512 */

struct Params {
  double *C;
  long long ldc;
  int M, N;
  int ksteps; /* padded K / 32 */
  int S;      /* digits per operand */
  const int *eA;
  const int *eB;
  int tiles_m, tiles_n;
  const int8_t *TA; /* [tiles_m][ksteps][S][4096] */
  const int8_t *TB; /* [tiles_n][ksteps][S][4096] */
  int prefetch;     /* k steps of L2 prefetch ahead of the shared-memory ring (0 = off) */
  int flags;        /* diagnostics: 1 = epilogue skips the C read-modify-write, 2 = no operand loads (MMA rate only) */
  unsigned int *progress; /* bring-up aid of the experimental 2-CTA kernel (PHPC_OZ_PROGRESS=1): host-mapped words, 8 per CTA,
                           * where every warp role records how far it got, readable from the host WHILE a kernel hangs */
};

/* instruction descriptor: s8 x s8 -> s32, A and B K-major */
__device__ __forceinline__ uint32_t idesc_i8(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

/* one lane of a fully converged warp (the others skip): keeps the surrounding code warp-uniform so
 * descriptors and addresses stay in uniform registers instead of being moved there per MMA */
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, int (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

/* 2^e as a double (e clamped to the normal range; INT_MIN exponents mean "all zero") */
__device__ __forceinline__ double pow2d(int e) {
  e = max(-1022, min(1023, e));
  return __hiloint2double((e + 1023) << 20, 0);
}

/* byte offset of element (row r < 128, k byte kb < 32) inside a canonical 4 KiB tile */
__host__ __device__ __forceinline__ int tile_offset(int r, int kb) { return (r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15); }

/* UMMA descriptor, K-major, no swizzle: LBO = 128 B between the two k chunks, SBO = 256 B between 8-row groups */
__device__ __forceinline__ uint64_t smem_desc_kmajor_noswz(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

template <int S_T, bool BAL = false>
__global__ void __launch_bounds__(THREADS, 1) ozaki_gemm_kernel(const Params p) {
  constexpr int DB = BAL ? 8 : DIGIT_BITS; /* bits between consecutive digit groups */
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES;
  const uint32_t tfull = bars + 16 * STAGES, tempty = tfull + 8;
  const uint32_t tmem_slot = tempty + 8;
  const uint32_t epi0 = bars + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;
  const int S = S_T > 0 ? S_T : p.S; /* compile-time digit count: the MMA issue loops unroll into straight-line uniform code */
  const int npass = (S + GROUPS_PER_PASS - 1) / GROUPS_PER_PASS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    /* ===== producer: two contiguous bulk copies per k step ===== */
    if (lane == 0 && !(p.flags & 2)) {
      int stage = 0;
      uint32_t phase = 0;
      const size_t step_bytes = (size_t)S * TILE_BYTES;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int tm, tn;
        tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
        const int8_t *ta = p.TA + (size_t)tm * p.ksteps * step_bytes;
        const int8_t *tb = p.TB + (size_t)tn * p.ksteps * step_bytes;
        for (int ps = 0; ps < npass; ++ps) {
          const int g_hi = S + 1 - GROUPS_PER_PASS * ps;
          const int d_hi = min(S, g_hi - 1);                      /* digits 1 .. d_hi take part in this pass */
          const uint32_t bytes = (uint32_t)d_hi * TILE_BYTES;
          const int sub = (2 * d_hi <= MAX_S) ? 2 : 1;            /* a pass that needs <= half the digit slots packs 2 k steps per stage */
          for (int ks = 0; ks < p.ksteps; ks += sub) {
            const int nsub = min(sub, p.ksteps - ks);
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t full = full0 + 8 * stage;
            mbar_expect_tx(full, 2 * bytes * nsub);
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            for (int h = 0; h < nsub; ++h) {
              bulk_load(sa + h * d_hi * SLOT_BYTES, ta + (size_t)(ks + h) * step_bytes, bytes, full);
              bulk_load(sa + (MAX_S + h * d_hi) * SLOT_BYTES, tb + (size_t)(ks + h) * step_bytes, bytes, full);
            }
            if (p.prefetch > 0 && ks + p.prefetch < p.ksteps) { /* pull the digits of a later k step into L2 */
              for (int h = 0; h < nsub; ++h) {
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ta + (size_t)(ks + h + p.prefetch) * step_bytes), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tb + (size_t)(ks + h + p.prefetch) * step_bytes), "r"(bytes) : "memory");
              }
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    /* ===== MMA issuer: the whole warp walks the loops (uniform control flow and addresses), one
     * elected lane issues tcgen05.mma / tcgen05.commit.  With the loops inside `if (lane == 0)` every
     * MMA cost ~140-180 cycles of register->uniform-register traffic (tools/umma_rate.cu). ===== */
    {
      const uint32_t idesc = idesc_i8(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t unit = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int ps = 0; ps < npass; ++ps, ++unit) {
          const int g_hi = S + 1 - GROUPS_PER_PASS * ps;
          const int g_lo = max(2, g_hi - GROUPS_PER_PASS + 1);
          mbar_wait(tempty, (unit & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int d_hi = min(S, g_hi - 1);
          const int sub = (2 * d_hi <= MAX_S) ? 2 : 1;
          for (int ks = 0; ks < p.ksteps; ks += sub) {
            const int nsub = min(sub, p.ksteps - ks);
            if (!(p.flags & 2)) mbar_wait(full0 + 8 * stage, phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            if (elect_one()) {
              for (int h = 0; h < nsub; ++h) {
                const uint64_t da0 = smem_desc_kmajor_noswz(sa + h * d_hi * SLOT_BYTES);
                const uint64_t db0 = smem_desc_kmajor_noswz(sa + (MAX_S + h * d_hi) * SLOT_BYTES);
                const uint32_t first = (ks + h) > 0 ? 1u : 0u;
                if (S_T > 0) {
                  /* fully unrolled: every (group, digit pair) of the pass with immediate descriptor offsets */
#pragma unroll
                  for (int gi = 0; gi < GROUPS_PER_PASS; ++gi) {
#pragma unroll
                    for (int t = 1; t <= MAX_S; ++t) {
                      const int g = g_hi - gi;
                      const int u = g - t;
                      if (g >= g_lo && t <= S && u >= 1 && u <= S)
                        umma_i8(tmem_base + (uint32_t)(g - g_lo) * BN, da0 + (uint64_t)((t - 1) * (SLOT_BYTES >> 4)),
                                    db0 + (uint64_t)((u - 1) * (SLOT_BYTES >> 4)), idesc, (t > max(1, g - S)) ? 1u : first);
                    }
                  }
                } else {
                  for (int g = g_hi; g >= g_lo; --g) {
                    const uint32_t tacc = tmem_base + (uint32_t)(g - g_lo) * BN;
                    const int t_lo = max(1, g - S), t_hi = min(S, g - 1);
                    for (int t = t_lo; t <= t_hi; ++t) {
                      const int u = g - t;
                      umma_i8(tacc, da0 + (uint64_t)((t - 1) * (SLOT_BYTES >> 4)), db0 + (uint64_t)((u - 1) * (SLOT_BYTES >> 4)), idesc,
                                  (t > t_lo) ? 1u : first);
                    }
                  }
                }
              }
              if (!(p.flags & 2)) umma_commit(empty0 + 8 * stage);
            }
            __syncwarp();
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (elect_one()) umma_commit(tfull);
          __syncwarp();
        }
      }
    }
  } else {
    /* ===== epilogue: combine the groups of a pass exactly, then one C += per element ===== */
    const int quarter = warp & 3;
    const uint32_t tr = epi0 + (uint32_t)(warp - 2) * EPI_WARP_BYTES;
    uint32_t unit = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int tm, tn;
      tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
      const int row0 = tm * BM + quarter * 32;
      const int my_row = row0 + lane;
      const int ea = (my_row < p.M) ? p.eA[my_row] : ZERO_EXP;
      const int rows_here = min(32, p.M - row0);
      for (int ps = 0; ps < npass; ++ps, ++unit) {
        const int g_hi = S + 1 - GROUPS_PER_PASS * ps;
        const int g_lo = max(2, g_hi - GROUPS_PER_PASS + 1);
        mbar_wait(tfull, unit & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (tn * BN + c0 >= p.N || rows_here <= 0) break;
          double acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = 0.0;
          for (int g = g_hi; g >= g_lo; --g) {
            int v[32];
            tmem_ld_32x32b_x32(tlane + (uint32_t)(g - g_lo) * BN + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const double w = pow2d(DB * (g_hi - g));
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fma((double)v[j], w, acc[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(tr + (uint32_t)(lane * 33 + j) * 8), "d"(acc[j]) : "memory");
          __syncwarp();
          const int col = tn * BN + c0 + lane;
          const int eb = (col < p.N) ? __ldg(p.eB + col) : ZERO_EXP;
          double *cptr = p.C + (long long)row0 * p.ldc + col;
          const bool col_ok = eb != ZERO_EXP && !(p.flags & 1);
          /* all 32 row loads of this lane's column are issued before any is used: one memory
           * round trip per 32x32 block instead of four */
          double cold[32];
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) cold[rr] = (col_ok && rr < rows_here) ? cptr[(long long)rr * p.ldc] : 0.0;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            double x;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(tr + (uint32_t)(rr * 33 + lane) * 8) : "memory");
            const int er = __shfl_sync(0xffffffffu, ea, rr);
            if (col_ok && rr < rows_here) {
              if (er == NONFINITE_EXP || eb == NONFINITE_EXP)
                cptr[(long long)rr * p.ldc] = __longlong_as_double(0x7ff8000000000000ll); /* Inf/NaN in the row or column */
              else if (er != ZERO_EXP && x != 0.0)
                cptr[(long long)rr * p.ldc] = cold[rr] + x * pow2d(BAL ? er + eb - 2 * BAL_BITS + 8 * (2 * S - g_hi) : er + eb - DIGIT_BITS * g_hi);
            }
          }
          __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


}  // namespace oz
}  // namespace phpc
