/*
 * ozaki_gemm2.cuh — EXPERIMENTAL 2-CTA variant of ozaki_gemm.cuh (PHPC_OZAKI_KERNEL=2cta).
 * Written at the end of round 1 with the GPU budget spent: it compiles for sm_100a but has NOT run on
 * hardware yet; it is never selected unless the environment asks for it.  tools/ozaki_variants.py is the
 * one-call validation (parity against the DMMA kernel and the integer model, then timing).
 *
 * Why: profiles/ozaki_experiments_r01.md — the 1-CTA kernel needs 96 KiB of digit tiles per 32-byte k step
 * for 36 MMAs of 65 cycles, i.e. 41 B/clk/SM out of L2, which is the L2->SM ceiling (~42 B/clk/SM, TMA
 * chip throughput in the microarchitecture notes); "no operand loads" ran 12-23 % faster.  A CTA pair
 * (cta_group::2) multiplies a 256 x 128 tile with M = 256 MMAs: each CTA stages its own 128 A rows but only
 * HALF of every B digit tile (64 of the 128 output columns), so the L2->SM and the shared-memory->tensor
 * core traffic per MMA drop by a quarter (6 KiB instead of 8 KiB per 128x128x32 per SM).
 *
 * Same arithmetic, schedule and epilogue as ozaki_gemm.cuh (K-outer, four group accumulators = all 512
 * TMEM columns of each CTA, two passes); what changes is the plumbing between the two CTAs of a pair:
 *   tile        pair tile = 256 rows x 128 columns; CTA rank r owns rows [256*tm2 + 128*r, +128) of it and
 *               lanes 0..127 of ITS OWN tensor memory hold them
 *   operands    every CTA: A digits of its rows (d x 4 KiB) and the half-major B digits of its rank
 *               (d x 2 KiB, ozaki_split.cuh store_offset(halves = 2)) per k step, 4-stage ring
 *   full        each CTA's loads complete ITS full[stage]; warp 1 of the peer CTA (which issues no MMAs)
 *               forwards that to the leader's peer_full[stage] with a remote mbarrier arrive
 *   MMA         leader CTA (rank 0) only: tcgen05.mma.cta_group::2, M = 256, N = 128
 *   empty/tfull tcgen05.commit.cta_group::2 ... multicast::cluster, mask 0b11: arrives in both CTAs
 *   tempty      lives in the leader: its 4 epilogue warps arrive locally, the peer's 4 remotely
 *
 * Second load path, template parameter TMA (PHPC_OZAKI_KERNEL=2cta-tma): the digit stores are described by tensor
 * maps (rows of 2 KiB) and every CTA loads with cp.async.bulk.tensor.2d.cta_group::2, whose completion bytes may be
 * sent to the mbarrier of the PEER CTA: both CTAs' copies complete the LEADER's full[stage] directly (the leader
 * expects the bytes of both), no relay warp and no peer_full barrier.  Same data, same shared-memory image.
 */
#pragma once
#include "ozaki_gemm.cuh"

namespace phpc {
namespace oz {

constexpr int B_HALF_BYTES = SLOT_BYTES / 2;                              /* 64 B^T rows x 32 B */
constexpr int STAGE2_BYTES = MAX_S * SLOT_BYTES + MAX_S * B_HALF_BYTES;    /* 48 KiB */
constexpr int STAGES2 = 4;
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256 + 4 * EPI_WARP_BYTES;
static_assert(SMEM2_BYTES <= 232448, "dynamic shared memory per CTA");

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
/* shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster */
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
/* wait that also acquires what CTAs of the cluster released before arriving */
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, 0x989680;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
/* arrives on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once all MMAs issued so far are done */
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

/* tensor maps of the digit stores, one per pass shape: box = {256 x 8 B = one 2 KiB row, 2*d rows (A) / d rows (B half)} */
struct StoreMaps {
  CUtensorMap a[2];
  CUtensorMap b[2];
};
/* the completion bytes go to `leader_bar`, a shared::cluster address that may lie in the other CTA of the pair */
__device__ __forceinline__ void tma2_load_rows(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(leader_bar), "r"(0), "r"(row)
      : "memory");
}

/* progress word of one warp role: (stage marker << 28) | (units done & 0xfffffff); slot 0 producer, 1 MMA issuer / relay,
 * 2..5 epilogue warps, 6 set-up / tear-down.  Markers: 1 waiting for a barrier, 2 past it, 3 role finished. */
__device__ __forceinline__ void progress_mark(const Params &p, int slot, uint32_t marker, uint32_t count) {
  if (p.progress) {
    volatile unsigned int *w = p.progress + (size_t)blockIdx.x * 8 + slot;
    *w = (marker << 28) | (count & 0x0fffffffu);
  }
}

/* Params as in ozaki_gemm.cuh with: tiles_m = number of 128-row tiles rounded up to EVEN (TA holds zero digits
 * for the padding tile), TB in half-major order. */
template <int S_T, bool BAL, bool TMA = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
    ozaki_gemm_2cta_kernel(const Params p, const __grid_constant__ StoreMaps maps) {
  static_assert(S_T >= 2 && S_T <= MAX_S, "digit count is a compile-time constant in this kernel");
  static_assert(!TMA || (S_T + GROUPS_PER_PASS - 1) / GROUPS_PER_PASS <= 2, "two tensor-map shapes: at most two passes");
  constexpr int S = S_T;
  constexpr int DB = BAL ? 8 : DIGIT_BITS;
  constexpr int NPASS = (S + GROUPS_PER_PASS - 1) / GROUPS_PER_PASS;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + STAGES2 * STAGE2_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES2, pfull0 = bars + 16 * STAGES2;
  const uint32_t tfull = bars + 24 * STAGES2, tempty = tfull + 8;
  const uint32_t tmem_slot = tempty + 8;
  const uint32_t epi0 = bars + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int tiles_m2 = p.tiles_m >> 1;
  const int total_tiles = tiles_m2 * p.tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
      mbar_init(pfull0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 8); /* 4 epilogue warps of each CTA of the pair */
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) progress_mark(p, 6, 1, 0);
  cluster_sync_all(); /* the peer's barriers exist before anything arrives on them remotely */
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  if (threadIdx.x == 0) progress_mark(p, 6, 2, tmem_base);

  if (warp == 0) {
    /* ===== producer (both CTAs): own A rows + own half of the B digit tiles ===== */
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, loads = 0;
      const size_t a_step = (size_t)S * SLOT_BYTES, b_step = (size_t)2 * S * B_HALF_BYTES;
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
        int tm2, tn;
        tile_coords(tile, tiles_m2, p.tiles_n, tm2, tn);
        const int8_t *ta = p.TA + (size_t)(2 * tm2 + (int)rank) * p.ksteps * a_step;
        const int8_t *tb = p.TB + (size_t)tn * p.ksteps * b_step + (size_t)rank * S * B_HALF_BYTES;
#pragma unroll 1
        for (int ps = 0; ps < NPASS; ++ps) {
          const int g_hi = S + 1 - GROUPS_PER_PASS * ps;
          const int d_hi = min(S, g_hi - 1);
          const int sub = (2 * d_hi <= MAX_S) ? 2 : 1;
          const uint32_t a_bytes = (uint32_t)d_hi * SLOT_BYTES, b_bytes = (uint32_t)d_hi * B_HALF_BYTES;
          for (int ks = 0; ks < p.ksteps; ks += sub) {
            const int nsub = min(sub, p.ksteps - ks);
            progress_mark(p, 0, 1, loads);
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            progress_mark(p, 0, 2, loads++);
            const uint32_t full = full0 + 8 * stage;
            const uint32_t sa = smem_base + stage * STAGE2_BYTES;
            const uint32_t sb = sa + MAX_S * SLOT_BYTES;
            if (TMA) {
              /* both CTAs' copies complete the LEADER's barrier, which expects the bytes of both */
              if (leader) mbar_expect_tx(full, 2 * (a_bytes + b_bytes) * nsub);
              const uint32_t leader_full = map_to_cta(full, 0);
              for (int h = 0; h < nsub; ++h) {
                /* store rows are 2 KiB: an A digit tile is 2 rows, a B half tile 1 row */
                const long long a_row = ((long long)(2 * tm2 + (int)rank) * p.ksteps + (ks + h)) * (2 * S);
                const long long b_row = (((long long)tn * p.ksteps + (ks + h)) * 2 + (int)rank) * S;
                tma2_load_rows(sa + h * d_hi * SLOT_BYTES, &maps.a[ps], leader_full, (int)a_row);
                tma2_load_rows(sb + h * d_hi * B_HALF_BYTES, &maps.b[ps], leader_full, (int)b_row);
              }
            } else {
              mbar_expect_tx(full, (a_bytes + b_bytes) * nsub);
              for (int h = 0; h < nsub; ++h) {
                bulk_load(sa + h * d_hi * SLOT_BYTES, ta + (size_t)(ks + h) * a_step, a_bytes, full);
                bulk_load(sb + h * d_hi * B_HALF_BYTES, tb + (size_t)(ks + h) * b_step, b_bytes, full);
              }
            }
            if (++stage == STAGES2) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      progress_mark(p, 0, 3, loads);
    }
  } else if (warp == 1 && !leader) {
    /* ===== peer CTA: forward "my operands of this stage have landed" to the leader ===== */
    if (lane == 0 && !TMA) {
      int stage = 0;
      uint32_t phase = 0, fwd = 0;
      const uint32_t leader_pfull0 = map_to_cta(pfull0, 0);
      for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
#pragma unroll 1
        for (int ps = 0; ps < NPASS; ++ps) {
          const int g_hi = S + 1 - GROUPS_PER_PASS * ps;
          const int d_hi = min(S, g_hi - 1);
          const int sub = (2 * d_hi <= MAX_S) ? 2 : 1;
          for (int ks = 0; ks < p.ksteps; ks += sub) {
            progress_mark(p, 1, 1, fwd);
            mbar_wait(full0 + 8 * stage, phase);
            mbar_arrive_cluster(leader_pfull0 + 8 * stage);
            progress_mark(p, 1, 2, fwd++);
            if (++stage == STAGES2) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      progress_mark(p, 1, 3, fwd);
    }
  } else if (warp == 1) {
    /* ===== leader CTA: MMA issuer for the pair (warp-uniform loops, one elected lane issues) ===== */
    const uint32_t idesc = idesc_i8(2 * BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t unit = 0, steps = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
#pragma unroll
      for (int ps = 0; ps < NPASS; ++ps, ++unit) {
        const int g_hi = S + 1 - GROUPS_PER_PASS * ps;
        const int g_lo = max(2, g_hi - GROUPS_PER_PASS + 1);
        const int d_hi = min(S, g_hi - 1);
        const int sub = (2 * d_hi <= MAX_S) ? 2 : 1;
        if (lane == 0) progress_mark(p, 1, 4, unit);
        mbar_wait_cluster(tempty, (unit & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int ks = 0; ks < p.ksteps; ks += sub) {
          const int nsub = min(sub, p.ksteps - ks);
          if (lane == 0) progress_mark(p, 1, 1, steps);
          if (TMA)
            mbar_wait_cluster(full0 + 8 * stage, phase); /* the peer's copies complete this barrier too */
          else
            mbar_wait(full0 + 8 * stage, phase);
          if (lane == 0) progress_mark(p, 1, 5, steps);
          if (!TMA) mbar_wait_cluster(pfull0 + 8 * stage, phase);
          if (lane == 0) progress_mark(p, 1, 2, steps++);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + stage * STAGE2_BYTES;
          const uint32_t sb = sa + MAX_S * SLOT_BYTES;
          if (elect_one()) {
            for (int h = 0; h < nsub; ++h) {
              const uint64_t da0 = smem_desc_kmajor_noswz(sa + h * d_hi * SLOT_BYTES);
              const uint64_t db0 = smem_desc_kmajor_noswz(sb + h * d_hi * B_HALF_BYTES);
              const uint32_t first = (ks + h) > 0 ? 1u : 0u;
#pragma unroll
              for (int gi = 0; gi < GROUPS_PER_PASS; ++gi) {
#pragma unroll
                for (int t = 1; t <= MAX_S; ++t) {
                  const int g = g_hi - gi;
                  const int u = g - t;
                  if (g >= g_lo && t <= S && u >= 1 && u <= S)
                    umma2_i8(tmem_base + (uint32_t)(g - g_lo) * BN, da0 + (uint64_t)((t - 1) * (SLOT_BYTES >> 4)),
                             db0 + (uint64_t)((u - 1) * (B_HALF_BYTES >> 4)), idesc, (t > max(1, g - S)) ? 1u : first);
                }
              }
            }
            umma2_commit_both(empty0 + 8 * stage);
          }
          __syncwarp();
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma2_commit_both(tfull);
        __syncwarp();
      }
    }
    if (lane == 0) progress_mark(p, 1, 3, steps);
  } else {
    /* ===== epilogue (both CTAs, own 128 rows): as ozaki_gemm.cuh, tempty lives in the leader ===== */
    const int quarter = warp & 3;
    const uint32_t tr = epi0 + (uint32_t)(warp - 2) * EPI_WARP_BYTES;
    const uint32_t leader_tempty = map_to_cta(tempty, 0);
    uint32_t unit = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += nclusters) {
      int tm2, tn;
      tile_coords(tile, tiles_m2, p.tiles_n, tm2, tn);
      const int row0 = (2 * tm2 + (int)rank) * BM + quarter * 32;
      const int my_row = row0 + lane;
      const int ea = (my_row < p.M) ? p.eA[my_row] : ZERO_EXP;
      const int rows_here = min(32, p.M - row0);
#pragma unroll 1
      for (int ps = 0; ps < NPASS; ++ps, ++unit) {
        const int g_hi = S + 1 - GROUPS_PER_PASS * ps;
        const int g_lo = max(2, g_hi - GROUPS_PER_PASS + 1);
        if (lane == 0) progress_mark(p, warp, 1, unit);
        mbar_wait(tfull, unit & 1);
        if (lane == 0) progress_mark(p, warp, 2, unit);
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
                cptr[(long long)rr * p.ldc] = __longlong_as_double(0x7ff8000000000000ll);
              else if (er != ZERO_EXP && x != 0.0)
                cptr[(long long)rr * p.ldc] = cold[rr] + x * pow2d(BAL ? er + eb - 2 * BAL_BITS + 8 * (2 * S - g_hi) : er + eb - DIGIT_BITS * g_hi);
            }
          }
          __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_tempty);
      }
    }
    if (lane == 0) progress_mark(p, warp, 3, unit);
  }

  /* both CTAs are done with each other's tensor memory and barriers before either one leaves */
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) progress_mark(p, 6, 4, 0);
  cluster_sync_all();
  if (threadIdx.x == 0) progress_mark(p, 6, 3, 0);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace oz
}  // namespace phpc
