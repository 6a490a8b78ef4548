/*
 * umma_rate.cu — issue-rate microbenchmark of tcgen05.mma (operands in shared memory, accumulators
 * in TMEM, no loads, zeros everywhere): cycles per MMA for several kinds and shapes, one CTA per SM.
 * Gives the int8 tensor peak the Ozaki kernel is measured against ("int8 tensor peak:
 * not measured").  Output: one JSON line per configuration.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"%s failed: %s line %d\n",#x,cudaGetErrorString(e),__LINE__); exit(1);} } while(0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { int kind; int M; int N; int accs; int reps; int swz; int rnd; };  /* rnd: pseudo-random operand bytes instead of zeros */   /* kind 0 = i8, 1 = f8f6f4 (e4m3), 2 = f16 (bf16) */

__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr) {
  uint64_t d = 0; d |= (uint64_t)((saddr & 0x3FFFF) >> 4); d |= (uint64_t)(128 >> 4) << 16; d |= (uint64_t)(256 >> 4) << 32; d |= (uint64_t)1 << 46; return d;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0; d |= (uint64_t)((saddr & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}

__global__ void __launch_bounds__(128, 1) umma_rate_kernel(Cfg c, long long *cycles) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, slot = bar + 16;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) {
    /* zeros do not toggle the datapath (a zero-operand run is not a sustained peak): rnd fills with a hash.  For the
     * float kinds the hash is masked so that no byte pattern is an Inf/NaN (exponent bits never all ones). */
    uint32_t v = 0;
    if (c.rnd) {
      v = (uint32_t)(i + 1) * 2654435761u + blockIdx.x * 40503u;
      v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
      if (c.kind == 1) v &= 0xB7B7B7B7u;       /* e4m3: clear one exponent bit per byte */
      else if (c.kind == 2) v &= 0xBF7FBF7Fu;  /* bf16: clear the top exponent bit of each half word */
    }
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4 * i), "r"(v));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 1) {
    /* whole warp converged, one elected lane issues (see ozaki_gemm.cuh): issue cost is then a few cycles per MMA */
    uint32_t idesc;
    if (c.kind == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10);       /* s32 <- s8 x s8 */
    else if (c.kind == 1) idesc = (1u << 4);                            /* f32 <- e4m3 x e4m3 */
    else idesc = (1u << 4) | (1u << 7) | (1u << 10);                    /* f32 <- bf16 x bf16 */
    idesc |= ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint64_t da = c.swz ? desc_sw128(base) : desc_noswz(base);
    const uint64_t db = c.swz ? desc_sw128(base + 32 * 1024) : desc_noswz(base + 32 * 1024);
    const uint32_t d0 = tmem, d1 = tmem + (c.accs > 1 ? c.N : 0);
    const long long t0 = clock64();
    uint32_t elected;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(elected));
    for (int r = 0; r < c.reps; r += 8) {
      if (elected) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t d = (j & 1) ? d1 : d0;
          if (c.kind == 0)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
          else if (c.kind == 1)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
        }
      }
      __syncwarp();
    }
    if (elected) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int smem = 96 * 1024 + 1024 + 64;
  CK(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long *d; CK(cudaMalloc(&d, sms * sizeof(long long)));
  long long *h = (long long *)malloc(sms * sizeof(long long));
  const char *kinds[3] = {"i8", "f8f6f4(e4m3)", "f16(bf16)"};
  Cfg cfgs[] = {
    {0,128,128,1,4000,0},{0,128,128,4,4000,0},{0,128,256,1,4000,0},{0,128,256,2,4000,0},{0,128,64,1,4000,0},{0,64,128,1,4000,0},{0,64,256,1,4000,0},
    {0,128,128,4,4000,1},{0,128,256,2,4000,1},
    {1,128,128,4,4000,0},{1,128,256,2,4000,0},{1,128,256,2,4000,1},
    {2,128,128,4,4000,0},{2,128,256,2,4000,0},{2,128,256,2,4000,1},
  };
  if (getenv("UMMA_SUSTAIN")) { /* seconds-long runs of the best int8 shape: the power-capped (sustained) int8 peak.
                                 * zero operands first (round 1: 4.4 POP/s, the datapath does not toggle), then random bytes */
    for (int rnd = 0; rnd < 2; ++rnd) {
      Cfg c = {0, 128, 256, 2, 40000000, 0, rnd};
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0)); umma_rate_kernel<<<sms, 128, smem>>>(c, d); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("{\"kind\": \"i8\", \"M\": 128, \"N\": 256, \"K\": 32, \"operands\": \"%s\", \"sustained_seconds\": %.2f, \"chip_tops_sustained\": %.1f}\n",
             rnd ? "random" : "zero", ms * 1e-3, 2.0 * 128 * 256 * 32 * (double)c.reps * sms / (ms * 1e-3) / 1e12);
      fflush(stdout);
    }
    return 0;
  }
  const int rnd_all = getenv("UMMA_RANDOM") ? 1 : 0; /* burst table with random operand bytes */
  for (auto &c : cfgs) {
    c.rnd = rnd_all;
    for (int one_sm = 0; one_sm < 2; ++one_sm) {
      const int grid = one_sm ? 1 : sms;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      umma_rate_kernel<<<grid, 128, smem>>>(c, d); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); umma_rate_kernel<<<grid, 128, smem>>>(c, d); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
      double avg = 0; long long mx = 0; for (int i = 0; i < grid; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; } avg /= grid;
      const int K = c.kind == 2 ? 16 : 32;
      const double macs = (double)c.M * c.N * K;
      printf("{\"kind\": \"%s\", \"M\": %d, \"N\": %d, \"K\": %d, \"accumulators\": %d, \"layout\": \"%s\", \"operands\": \"%s\", \"sms\": %d, \"cycles_per_mma\": %.1f, "
             "\"mac_per_clk_per_sm\": %.0f, \"chip_tops_at_event_time\": %.1f}\n", kinds[c.kind], c.M, c.N, K, c.accs, c.swz ? "sw128" : "noswz", c.rnd ? "random" : "zero", grid,
             avg / c.reps, macs / (avg / c.reps), 2.0 * macs * c.reps * grid / (ms * 1e-3) / 1e12);
      fflush(stdout);
    }
  }
  return 0;
}
