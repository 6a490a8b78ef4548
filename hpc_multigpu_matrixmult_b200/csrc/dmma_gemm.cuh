/*
 * dmma_gemm.cuh — the local block GEMM of the SUMMA hot path for sm_100a:
 *     C[M x N] += A[M x K] * B[K x N]      (FP64, all row-major, C accumulating)
 *
 * Replaces the reference's gemm_kernel (reference src/phpc_gemm.cu:6-57: one
 * thread per C element, 32x32 shared tile, 2 barriers per 32-deep phase).
 * Same contract: zero-padded edge tiles (reference :38-46) and a final
 * "C += sum" read-modify-write (reference :54-55).
 *
 * Design (B200-first, not a translation):
 *   - persistent CTAs, one per SM, pulling 128x128 output tiles from a global
 *     atomic tile counter (dynamic, so SMs borrowed by an overlapping NCCL
 *     broadcast just take fewer tiles), tiles rasterised in 16-row bands for
 *     L2 reuse of the A/B panels;
 *   - warp specialised: 1 producer warpgroup (one elected lane; setmaxnreg
 *     hands its registers to the consumers) feeds a 4-stage
 *     shared-memory ring with TMA (cp.async.bulk.tensor.2d, 128-byte swizzle,
 *     hardware zero fill outside the matrix = the reference's zero padding),
 *     8 consumer warps (2x4, 64x32 warp tiles) wait on mbarriers and issue
 *     FP64 tensor-core MMAs (mma.sync m8n8k4.f64 -> SASS DMMA.8x8x4) with the
 *     accumulators in registers.  tcgen05.mma has no FP64 kind (kind::f16,
 *     tf32, i8, f8f6f4, mxf*: none is 64-bit), so DMMA is the only tensor-core
 *     path in the reference's element type on sm_100a; see DESIGN.md.
 *   - fragment loads are bank-conflict free 16-byte LDS: the k index inside a
 *     group of 8 and the row/column inside a group of 8/16 are permuted
 *     consistently between A, B and C (a GEMM is invariant under a common
 *     permutation of k, and a permutation of rows/cols only relabels C).
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace phpc {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;      /* 16 doubles = one 128-byte swizzle row of A */
constexpr int STAGES = 4;
constexpr int WARPS_M = 2;  /* warp tile 64 x 32 */
constexpr int WARPS_N = 4;
constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
constexpr int PRODUCER_WARPS = 4; /* one warpgroup so setmaxnreg can hand its registers to the consumers */
constexpr int GEMM_THREADS = (PRODUCER_WARPS + CONSUMER_WARPS) * 32;
constexpr int PRODUCER_REGS = 40;
constexpr int CONSUMER_REGS = 232;
constexpr int A_TILE_BYTES = BM * BK * 8;       /* 16 KiB, [128 rows][16 k]     */
constexpr int B_BOX_COLS = 16;                  /* one TMA box = [16 k][16 n]   */
constexpr int B_BOX_BYTES = BK * B_BOX_COLS * 8;
constexpr int B_BOXES = BN / B_BOX_COLS;
constexpr int B_TILE_BYTES = B_BOXES * B_BOX_BYTES;
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* align slack */ + 256 /* barriers */;
constexpr int GROUP_M = 16; /* raster band height in tiles */

struct GemmParams {
  double *C;
  long long ldc;
  int M, N, K;
  int tiles_m, tiles_n, k_iters;
  unsigned int *sched; /* [0] next tile, [1] CTAs finished (self-resetting) */
  const int *guard;    /* fallback launches of the tcgen05 path (phpc_launch_ozaki): run only when *guard != 0; NULL = always run */
};

/* ---- thin PTX wrappers ------------------------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

/* tile index -> (tile row, tile col), 16-row bands walked column by column */
__device__ __forceinline__ void tile_coords(int idx, int tiles_m, int tiles_n, int &tm, int &tn) {
  const int band_tiles = GROUP_M * tiles_n;
  const int band = idx / band_tiles;
  const int first_m = band * GROUP_M;
  const int rows = min(GROUP_M, tiles_m - first_m);
  const int r = idx - band * band_tiles;
  tm = first_m + r % rows;
  tn = r / rows;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
    dmma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  if (p.guard && *p.guard == 0) return; /* uniform over the grid, before the tile counter is touched */
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u; /* 128B swizzle atoms are 1 KiB */
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;           /* full[STAGES], empty[STAGES], tile[STAGES] */
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tile0 = bars + 16 * STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp < PRODUCER_WARPS) {
    /* ===== producer warpgroup: gives its registers away; one lane drives the tile scheduler and TMA ===== */
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
    if (warp == 0 && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (;;) {
        const int tile = (int)atomicAdd(p.sched, 1u);
        const bool done = tile >= total_tiles;
        int tm = 0, tn = 0;
        if (!done) tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
        const int iters = done ? 1 : p.k_iters;
        for (int kit = 0; kit < iters; ++kit) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          if (kit == 0) {
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(tile0 + 4 * stage), "r"(done ? -1 : tile) : "memory");
          }
          const uint32_t full = full0 + 8 * stage;
          if (done) {
            mbar_arrive(full); /* sentinel stage: no data */
          } else {
            mbar_expect_tx(full, STAGE_BYTES);
            const uint32_t sa = smem_base + stage * STAGE_BYTES;
            const uint32_t sb = sa + A_TILE_BYTES;
            const int k0 = kit * BK;
            tma_load_2d(sa, &tmA, full, k0, tm * BM);
#pragma unroll
            for (int jb = 0; jb < B_BOXES; ++jb) tma_load_2d(sb + jb * B_BOX_BYTES, &tmB, full, tn * BN + jb * B_BOX_COLS, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (done) break;
      }
    }
  } else {
    /* ===== consumers: 8 warps, 64x32 warp tiles, DMMA.8x8x4 ===== */
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
    const int cw = warp - PRODUCER_WARPS;
    const int wm = cw / WARPS_N, wn = cw % WARPS_N;
    const int g = lane >> 2, t = lane & 3;
    const int rho = (g >> 1) | ((g & 1) << 2); /* fragment row g <-> tile row rho within each group of 8 */
    /* A: row = wm*64 + i*8 + rho (row&7 == rho); 16B chunk (4j+t)^rho holds A[row][8j+2t .. 8j+2t+1] */
    const uint32_t a_off = (uint32_t)((wm * 64 + rho) * 128);
    const uint32_t a_chunk0 = (uint32_t)(((0 + t) ^ rho) << 4), a_chunk1 = (uint32_t)(((4 + t) ^ rho) << 4);
    /* B: box = wn*2+q, row kk = 8j+2t+e (kk&7 == 2t+e), 16B chunk g^(2t+e) holds B[kk][n0+2g .. n0+2g+1] */
    const uint32_t b_off = (uint32_t)(A_TILE_BYTES + wn * 2 * B_BOX_BYTES + 2 * t * 128);
    const uint32_t b_chunk0 = (uint32_t)((g ^ (2 * t)) << 4), b_chunk1 = (uint32_t)((g ^ (2 * t + 1)) << 4);

    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
      mbar_wait(full0 + 8 * stage, phase);
      int tile;
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(tile) : "r"(tile0 + 4 * stage) : "memory");
      if (tile < 0) break;

      double acc[8][2][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][q][e] = 0.0;

      for (int kit = 0; kit < p.k_iters; ++kit) {
        if (kit > 0) mbar_wait(full0 + 8 * stage, phase);
        const uint32_t sbase = smem_base + stage * STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          double2 af[8], bf[2][2];
          const uint32_t aaddr = sbase + a_off + (j ? a_chunk1 : a_chunk0);
          const uint32_t baddr = sbase + b_off + j * 1024;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            bf[q][0] = lds128(baddr + q * B_BOX_BYTES + b_chunk0);
            bf[q][1] = lds128(baddr + q * B_BOX_BYTES + 128 + b_chunk1);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) af[i] = lds128(aaddr + i * 1024);
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const double a = e ? af[i].y : af[i].x;
                dmma_8x8x4(acc[i][q][0], acc[i][q][1], a, bf[q][e].x); /* columns n0+2g   -> C cols 4t, 4t+2   */
                dmma_8x8x4(acc[i][q][2], acc[i][q][3], a, bf[q][e].y); /* columns n0+2g+1 -> C cols 4t+1, 4t+3 */
              }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stage);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }

      /* epilogue: C += acc (reference src/phpc_gemm.cu:54-55), each thread owns 4 consecutive columns */
      int tm, tn;
      tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
      const int row0 = tm * BM + wm * 64 + rho;
      const int col0 = tn * BN + wn * 32 + 4 * t;
      const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row0 + i * 8;
        if (row >= p.M) continue;
        double *crow = p.C + (long long)row * p.ldc;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int col = col0 + q * 16;
          const double v0 = acc[i][q][0], v1 = acc[i][q][2], v2 = acc[i][q][1], v3 = acc[i][q][3];
          if (vec_ok && col + 3 < p.N) {
            double2 *ptr = reinterpret_cast<double2 *>(crow + col);
            double2 lo = ptr[0], hi = ptr[1];
            lo.x += v0;
            lo.y += v1;
            hi.x += v2;
            hi.y += v3;
            ptr[0] = lo;
            ptr[1] = hi;
          } else {
            if (col < p.N) crow[col] += v0;
            if (col + 1 < p.N) crow[col + 1] += v1;
            if (col + 2 < p.N) crow[col + 2] += v2;
            if (col + 3 < p.N) crow[col + 3] += v3;
          }
        }
      }
    }
  }

  /* last CTA out resets the scheduler words so the next launch needs no memset */
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(p.sched + 1, 1u);
    if (prev == gridDim.x - 1) {
      p.sched[0] = 0;
      p.sched[1] = 0;
      __threadfence();
    }
  }
}

}  // namespace phpc
