/*
 * ozaki_gemm.cuh — FP64 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM).
 *
 * tcgen05.mma has no FP64 kind, so the FP64 product is rebuilt EXACTLY from integer products (the
 * Ozaki scheme): every row of A and every column of B of a K chunk is scaled by a power of two,
 * rounded to BAL_BITS = 54 bits below its row / column maximum and written in base 256 with
 * BALANCED digits in [-128, 127] (ozaki_split.cuh),
 *     a_ik = 2^(eA[i]-54) * sum_{t=1..7} A_t[i][k] * 256^(7-t),   b_kj likewise with eB[j], B_u,
 * the digit matrices are multiplied on the int8 tensor pipe with exact int32 accumulation in TMEM
 * (|sum over K <= 8192 of up to 7 products| < 2^31, no rounding at all), and
 *     C[i][j] += 2^(eA[i]+eB[j]-108) * sum_g 256^(14-g) * P_g[i][j],   P_g = sum_{t+u=g} A_t.B_u
 * is applied in FP64 by the epilogue warps.  Groups g > 8 are dropped: their weight is below 2^-56 of
 * the row/column scale and balanced digits keep their sign, so what is dropped still cancels.
 * 28 int8 MMAs stand for one FP64 MMA.  The reference kernel this replaces is gemm_kernel of
 * src/phpc_gemm.cu:6-57 (same C += A.B contract); the arithmetic differs from it only in the
 * order of exact partial sums and in the two FP64 roundings per K chunk (joining the two passes, adding to C).
 *
 * Kernel (one CTA per SM, persistent over 128 x 128 output tiles, static round robin):
 *   K-outer schedule  per 32-byte k step ALL needed digit tiles of A and B are staged once (one
 *               4 KiB slot per digit matrix) and every pair (t,u) of up to four groups is issued from
 *               them, one TMEM accumulator (128 columns) per group = all 512 TMEM columns:
 *                 pass 1  groups 8 .. 5  (22 digit products, needs every digit)
 *                 pass 2  groups 4 .. 2  (6 digit products, digits 1..3 only; two k steps per stage)
 *   paired MMAs       digits u and u+1 of B lie back to back in a stage, and 8-row groups are 256 B apart
 *               in the canonical layout, so ONE tcgen05.mma with N = 256 multiplies A_t by
 *               [B_u | B_u+1] into the ADJACENT accumulators of groups t+u and t+u+1: 16 instructions
 *               instead of 28 per k step and A is read from shared memory once per two digit
 *               products (96 instead of 128 B/clk of shared-memory reads: room for the bulk copies
 *               that refill the ring; measured +8 % and -38 % DRAM traffic, profiles/ozaki_knobs_r02.jsonl)
 *   digit stores  written by the split kernels ALREADY in the shared-memory order the tensor core
 *               wants (UMMA canonical K-major, no swizzle: 8-row x 16-byte core matrices; a 128-row x
 *               32-byte tile = 4 KiB, k chunks 128 B apart, 8-row groups 256 B apart), tile after tile:
 *               store[row tile][k step][digit][4 KiB], so a k step of a pass is ONE contiguous global
 *               range per operand.
 *   wave start  the CTAs working on tiles i*grid .. (i+1)*grid-1 (16 tile rows x ~9 tile columns) share
 *               A row panels and B column panels, which only hit in L2 if they are streamed at the same
 *               time.  Free-running CTAs drift apart by more than a tile (measured 470-580 us after 110
 *               tiles of 416 us) and the panels are then fetched from DRAM again and again (623 GB for a
 *               32768 x 8192 x 32768 launch).  The producers therefore start every tile together: one
 *               atomic counter per wave (all CTAs are resident: grid <= SM count, one CTA per SM).
 *   warp 0      producer: two cp.async.bulk copies per k step into a 4-stage mbarrier ring (4 x 56 KiB)
 *   warp 1      TMEM allocator + MMA issuer: the whole warp walks warp-uniform, fully unrolled code
 *               and one elected lane issues tcgen05.mma.kind::i8 / tcgen05.commit (with the loops inside
 *               `if (lane == 0)` every MMA cost 140-180 cycles of register -> uniform register moves
 *               instead of 65; tools/umma_rate.cu, profiles/umma_rate*_r01.jsonl)
 *   warps 4-11  epilogue (two warpgroups; setmaxnreg moves registers from warps 0-3 to them): thread = one tile row x 64
 *               columns.  After each pass the int32 accumulators are combined exactly in FP64 into 64 registers per
 *               thread (tcgen05.ld) and TMEM is handed back to the MMA warp at once; after pass 2 the two partial
 *               results are joined with one rounding and C gets ONE read-modify-write per element and K chunk
 *               (32-byte accesses, each thread owns 512 contiguous bytes of its row) while the MMA warp is already
 *               working on the next tile: no accumulator double buffering needed, and half the C traffic of a
 *               per-pass update
 *   guard       *p.guard != 0 (set by the exponent kernels: non-finite input, exponents near the FP64
 *               range limits, rows/columns spanning more than MAX_SPREAD binary orders of magnitude)
 *               makes the kernel return at once; the native-FP64 DMMA kernel launched right after it
 *               with the opposite predicate computes that K chunk instead (phpc_launch_ozaki).
 * Earlier variants (pair-outer 128x256 tiles with TMA; K-outer with 16 TMA boxes per step; truncated
 * 7-bit digits, 36 products; a 2-CTA cta_group::2 kernel; clusters of two CTAs with the A digits multicast: same speed,
 * profiles/ozaki_knobs_pair_multicast_r02.jsonl) and what was measured on them are in
 * profiles/ozaki_experiments_r01.md, profiles/ozaki_variants_r02.jsonl and profiles/ozaki_knobs_r02.jsonl.
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dmma_gemm.cuh" /* mbarrier wrappers, smem_u32, tile_coords */

namespace phpc {
namespace oz {

constexpr int S = 7;          /* digits per operand */
constexpr int DIGIT_BITS = 8; /* balanced base-256 digits */
constexpr int BAL_BITS = 54;  /* bits kept below the row / column scale */
constexpr int PRODUCTS = S * (S + 1) / 2;
constexpr int KC_MAX = 8192;  /* K chunk: 7 products x 8192 x 128^2 < 2^31 (exact up to K = 18724) */
constexpr int ZERO_EXP = -2147483647 - 1; /* exponent of an all-zero row / column */
constexpr int NONFINITE_EXP = 2147483647; /* the row / column holds an Inf or NaN (the guard sends the chunk to the DMMA kernel) */
constexpr int MAX_SPREAD = 40; /* guard: nonzero entries of one row / column of a K chunk may span at most 2^40 (every entry keeps >= 16 bits) */
constexpr int EXP_SUM_MIN = -960, EXP_SUM_MAX = 960; /* guard: eA[i] + eB[j] stays where 2^(eA+eB-108+..) and its product are normal numbers */

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BKB = 32;                /* bytes of k per step = one int8 MMA (K = 32) */
constexpr int SLOT_BYTES = BM * BKB;   /* one digit tile: 128 rows x 32 B */
constexpr int TILE_BYTES = SLOT_BYTES;
constexpr int STAGE_BYTES = 2 * S * SLOT_BYTES; /* A digit slots then B digit slots: 56 KiB */
constexpr int STAGES = 4;
constexpr int THREADS = 384;           /* warpgroup 0: warp 0 producer, warp 1 MMA issuer (warps 2, 3 idle); warpgroups 1-2: 8 epilogue warps */
constexpr int EPI_WARPS = 8;
constexpr int CTRL_REGS = 40, EPI_REGS = 232; /* setmaxnreg: the epilogue threads each hold 64 FP64 partial results across a pass */
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int GROUPS_PER_PASS = 4;
constexpr int NPASS = 2;
constexpr int TMEM_COLS = GROUPS_PER_PASS * BN; /* 512 */
static_assert(SMEM_BYTES <= 232448, "dynamic shared memory per CTA");

struct Params {
  double *C;
  long long ldc;
  int M, N;
  int ksteps; /* padded K / 32 */
  const int *eA;
  const int *eB;
  int tiles_m, tiles_n;
  const int8_t *TA; /* [tiles_m][ksteps][S][4096] */
  const int8_t *TB; /* [tiles_n][ksteps][S][4096] */
  const int *guard;        /* != 0: this K chunk belongs to the native-FP64 kernel, return at once */
  unsigned int *wave_sync; /* one zeroed counter per wave of gridDim.x tiles */
  int flags;               /* diagnostics (tools/ozaki_knobs.py): 1 = epilogue skips the C read-modify-write, 2 = no operand loads,
                            * 4 = no wave synchronisation */
  unsigned long long *tstamp; /* diagnostics: globaltimer at the start of every tile's loads [2*tile] and end of its epilogue [2*tile+1] */
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

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

/* 2^e as a double for e in the normal range [-1022, 1023] (the guard keeps the kernel inside it) */
__device__ __forceinline__ double pow2d(int e) { return __hiloint2double((e + 1023) << 20, 0); }

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

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

/* groups of pass PS: G_HI .. G_LO, digits 1 .. D_HI take part */
template <int PS>
struct Pass {
  static constexpr int G_HI = S + 1 - GROUPS_PER_PASS * PS;
  static constexpr int G_LO = (G_HI - GROUPS_PER_PASS + 1) > 2 ? (G_HI - GROUPS_PER_PASS + 1) : 2;
  static constexpr int D_HI = (G_HI - 1) < S ? (G_HI - 1) : S;
  static constexpr int SUB = (2 * D_HI <= S) ? 2 : 1; /* a pass that needs <= half the digit slots packs 2 k steps per stage */
};

/* All MMAs of one 32-byte k step of pass PS: straight-line code with immediate descriptor offsets, t-major so that the first
 * MMA into every accumulator of the pass is a t = 1 product (with g <= S + 1 every group starts at t = 1). */
template <int PS>
__device__ __forceinline__ void issue_kstep(uint32_t tmem_base, uint64_t da0, uint64_t db0, uint32_t first) {
  constexpr int G_HI = Pass<PS>::G_HI, G_LO = Pass<PS>::G_LO;
  const uint32_t idesc1 = idesc_i8(BM, BN), idesc2 = idesc_i8(BM, 2 * BN);
#pragma unroll
  for (int t = 1; t <= S; ++t) {
    const int u_lo = (G_LO - t) > 1 ? (G_LO - t) : 1;
    const int u_hi = (G_HI - t) < S ? (G_HI - t) : S;
#pragma unroll
    for (int u = 1; u <= S; ++u) {
      if (u < u_lo || u > u_hi) continue;
      if ((u - u_lo) & 1) continue; /* covered by the N = 256 MMA issued for u - 1 */
      const bool pair = u + 1 <= u_hi;
      umma_i8(tmem_base + (uint32_t)(t + u - G_LO) * BN, da0 + (uint64_t)((t - 1) * (SLOT_BYTES >> 4)), db0 + (uint64_t)((u - 1) * (SLOT_BYTES >> 4)),
              pair ? idesc2 : idesc1, t > 1 ? 1u : first);
    }
  }
}

/* producer side of one pass of one tile */
template <int PS>
__device__ __forceinline__ void load_pass(const Params &p, const int8_t *ta, const int8_t *tb, uint32_t smem_base, uint32_t full0, uint32_t empty0,
                                          int &stage, uint32_t &phase) {
  constexpr int D_HI = Pass<PS>::D_HI, SUB = Pass<PS>::SUB;
  constexpr uint32_t bytes = (uint32_t)D_HI * TILE_BYTES;
  constexpr size_t step_bytes = (size_t)S * TILE_BYTES;
  for (int ks = 0; ks < p.ksteps; ks += SUB) {
    const int nsub = min(SUB, p.ksteps - ks);
    mbar_wait(empty0 + 8 * stage, phase ^ 1);
    const uint32_t full = full0 + 8 * stage;
    mbar_expect_tx(full, 2 * bytes * nsub);
    const uint32_t sa = smem_base + stage * STAGE_BYTES;
    for (int h = 0; h < nsub; ++h) {
      bulk_load(sa + h * D_HI * SLOT_BYTES, ta + (size_t)(ks + h) * step_bytes, bytes, full);
      bulk_load(sa + (S + h * D_HI) * SLOT_BYTES, tb + (size_t)(ks + h) * step_bytes, bytes, full);
    }
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
}

/* MMA side of one pass of one tile */
template <int PS>
__device__ __forceinline__ void mma_pass(const Params &p, uint32_t tmem_base, uint32_t smem_base, uint32_t full0, uint32_t empty0, uint32_t tfull,
                                         uint32_t tempty, uint32_t unit, int &stage, uint32_t &phase) {
  constexpr int D_HI = Pass<PS>::D_HI, SUB = Pass<PS>::SUB;
  mbar_wait(tempty, (unit & 1) ^ 1); /* the epilogue has drained the accumulators of the previous pass */
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int ks = 0; ks < p.ksteps; ks += SUB) {
    const int nsub = min(SUB, p.ksteps - ks);
    if (!(p.flags & 2)) mbar_wait(full0 + 8 * stage, phase);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t sa = smem_base + stage * STAGE_BYTES;
    if (elect_one()) {
      for (int h = 0; h < nsub; ++h) {
        const uint64_t da0 = smem_desc_kmajor_noswz(sa + h * D_HI * SLOT_BYTES);
        const uint64_t db0 = smem_desc_kmajor_noswz(sa + (S + h * D_HI) * SLOT_BYTES);
        issue_kstep<PS>(tmem_base, da0, db0, (ks + h) > 0 ? 1u : 0u);
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

/* Epilogue, part 1 (per pass): drain the accumulators of pass PS into registers.  Thread = one row of the tile (TMEM lane) and
 * 64 of its 128 columns; part[j] = sum_g 256^(G_HI-g) P_g exactly (|sum| < 2^53); pass 1 (groups 8..5) starts acc, pass 2
 * (groups 4..2, weight 2^32 above it) is added with ONE rounding: acc = fl(part2 * 2^32 + part1).  The accumulators are handed
 * back to the MMA warp as soon as the loads have completed, i.e. before anything touches global memory. */
template <int PS>
__device__ __forceinline__ void epilogue_collect(uint32_t tmem_base, uint32_t tfull, uint32_t tempty, uint32_t unit, int quarter, int half, int lane,
                                                 double (&acc)[64]) {
  constexpr int G_HI = Pass<PS>::G_HI, G_LO = Pass<PS>::G_LO;
  mbar_wait(tfull, unit & 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    double part[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) part[j] = 0.0;
#pragma unroll
    for (int g = G_HI; g >= G_LO; --g) {
      int v[16];
      tmem_ld_32x32b_x16(tlane + (uint32_t)(g - G_LO) * BN + c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const double w = pow2d(DIGIT_BITS * (G_HI - g));
#pragma unroll
      for (int j = 0; j < 16; ++j) part[j] = fma((double)v[j], w, part[j]); /* exact */
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[c0 + j] = (PS == 0) ? part[j] : fma(part[j], pow2d(DIGIT_BITS * GROUPS_PER_PASS), acc[c0 + j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (lane == 0) mbar_arrive(tempty);
}

/* Epilogue, part 2 (per tile): C[row][col0 .. col0+63] += acc * 2^(eA[row] + eB[col] - 60), one read-modify-write of C per K
 * chunk.  Every thread owns 512 contiguous bytes of its row and moves them with 32-byte accesses (whole sectors); this runs
 * while the MMA warp is already working on the next tile. */
__device__ __forceinline__ void epilogue_update(const Params &p, const double (&acc)[64], int row, int col0, int ea) {
  constexpr int SCALE = -2 * BAL_BITS + DIGIT_BITS * (2 * S - Pass<0>::G_HI); /* weight of group 8 relative to 2^(eA+eB) */
  if (row >= p.M || ea == ZERO_EXP || (p.flags & 1)) return;
  double *crow = p.C + (long long)row * p.ldc;
  const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 31) == 0);
#pragma unroll
  for (int j0 = 0; j0 < 64; j0 += 16) { /* 4 x 32 bytes in flight per thread */
    double c[16];
    int eb[16];
    bool vec[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = col0 + j0 + 4 * q;
      vec[q] = vec_ok && col + 3 < p.N;
#pragma unroll
      for (int e = 0; e < 4; ++e) eb[4 * q + e] = (col + e < p.N) ? __ldg(p.eB + col + e) : ZERO_EXP;
      if (vec[q]) {
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c[4 * q]), "=d"(c[4 * q + 1]), "=d"(c[4 * q + 2]), "=d"(c[4 * q + 3]) : "l"(crow + col) : "memory");
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) c[4 * q + e] = (col + e < p.N) ? crow[col + e] : 0.0;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = col0 + j0 + 4 * q;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double x = acc[j0 + 4 * q + e];
        if (x != 0.0 && eb[4 * q + e] != ZERO_EXP) c[4 * q + e] += x * pow2d(ea + eb[4 * q + e] + SCALE);
      }
      if (vec[q]) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(crow + col), "d"(c[4 * q]), "d"(c[4 * q + 1]), "d"(c[4 * q + 2]), "d"(c[4 * q + 3]) : "memory");
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < p.N && acc[j0 + 4 * q + e] != 0.0 && eb[4 * q + e] != ZERO_EXP) crow[col + e] = c[4 * q + e];
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) ozaki_gemm_kernel(const Params p) {
  if (p.guard && *p.guard != 0) return; /* uniform over the grid: this chunk goes to the native-FP64 kernel */
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES;
  const uint32_t tfull = bars + 16 * STAGES, tempty = tfull + 8;
  const uint32_t tmem_slot = tempty + 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;
  const int worker = (int)blockIdx.x, workers = (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, EPI_WARPS);
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

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CTRL_REGS));
    if (warp == 0) {
      /* ===== producer: two contiguous bulk copies per k step ===== */
      if (lane == 0 && !(p.flags & 2)) {
        int stage = 0;
        uint32_t phase = 0;
        constexpr size_t step_bytes = (size_t)S * TILE_BYTES;
        int wave = 0;
        for (int tile = worker; tile < total_tiles; tile += workers, ++wave) {
          int tm, tn;
          tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
          const int8_t *ta = p.TA + (size_t)tm * p.ksteps * step_bytes;
          const int8_t *tb = p.TB + (size_t)tn * p.ksteps * step_bytes;
          if (p.tstamp) p.tstamp[2 * (size_t)tile] = globaltimer_ns();
          if (!(p.flags & 4)) {
            /* all CTAs of this wave start streaming their panels together (every CTA is resident: grid <= SM count) */
            unsigned int *ctr = p.wave_sync + wave;
            const long long left = (long long)total_tiles - (long long)wave * workers;
            const unsigned int expect = (unsigned int)(left < workers ? left : workers);
            atomicAdd(ctr, 1u);
            unsigned int seen;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
            } while (seen < expect);
          }
          load_pass<0>(p, ta, tb, smem_base, full0, empty0, stage, phase);
          load_pass<1>(p, ta, tb, smem_base, full0, empty0, stage, phase);
        }
      }
    } else if (warp == 1) {
      /* ===== MMA issuer: the whole warp walks the loops (uniform control flow and addresses), one
       * elected lane issues tcgen05.mma / tcgen05.commit ===== */
      int stage = 0;
      uint32_t phase = 0;
      uint32_t unit = 0;
      for (int tile = worker; tile < total_tiles; tile += workers) {
        mma_pass<0>(p, tmem_base, smem_base, full0, empty0, tfull, tempty, unit++, stage, phase);
        mma_pass<1>(p, tmem_base, smem_base, full0, empty0, tfull, tempty, unit++, stage, phase);
      }
    }
  } else {
    /* ===== epilogue (8 warps): drain both passes into registers, then ONE C += per element and K chunk ===== */
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    const int ew = warp - 4;
    const int quarter = warp & 3; /* the TMEM lanes a warp may read: 32 * (warp id % 4) */
    const int half = ew >> 2;     /* columns 0..63 or 64..127 of the tile */
    uint32_t unit = 0;
    for (int tile = worker; tile < total_tiles; tile += workers) {
      int tm, tn;
      tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
      const int row = tm * BM + quarter * 32 + lane;
      const int ea = (row < p.M) ? p.eA[row] : ZERO_EXP;
      double acc[64];
      epilogue_collect<0>(tmem_base, tfull, tempty, unit++, quarter, half, lane, acc);
      epilogue_collect<1>(tmem_base, tfull, tempty, unit++, quarter, half, lane, acc);
      epilogue_update(p, acc, row, tn * BN + half * 64, ea);
      if (p.tstamp && ew == 0 && lane == 0) p.tstamp[2 * (size_t)tile + 1] = globaltimer_ns();
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
