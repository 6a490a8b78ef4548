/* phpc_internal.h — shared by the translation units of libphpc_b200.so (not installed). */
#pragma once
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define PHPC_B200_VERSION 200 /* 200 (round 2): one tcgen05 kernel (7 balanced digits, paired MMAs, wave start) + guard -> DMMA; phpc_gemm_device_ozaki lost `slices` */

#define PHPC_MAX_DEVICES 16

[[noreturn]] void phpc_die(const char *what, const char *detail, const char *file, int line);

#define CUDA_CHECK(call)                                                           \
  do {                                                                             \
    cudaError_t err__ = (call);                                                    \
    if (err__ != cudaSuccess) phpc_die(#call, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

#define CUBLAS_CHECK(call)                                                \
  do {                                                                    \
    cublasStatus_t st__ = (call);                                         \
    if (st__ != CUBLAS_STATUS_SUCCESS) {                                  \
      char msg__[64];                                                     \
      snprintf(msg__, sizeof msg__, "cublas status %d", (int)st__);       \
      phpc_die(#call, msg__, __FILE__, __LINE__);                         \
    }                                                                     \
  } while (0)

#define PHPC_REQUIRE(cond, msg)                                  \
  do {                                                           \
    if (!(cond)) phpc_die(#cond, msg, __FILE__, __LINE__);       \
  } while (0)

/* grow-only cached device buffer (the reference cudaMallocAsync/cudaFreeAsync's
 * three buffers on every k-step, src/phpc_gemm.cu:106-108,123-125) */
struct DevBuf {
  void *ptr = nullptr;
  size_t bytes = 0;
};

struct DeviceCtx {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t compute = nullptr; /* GEMM launches */
  cudaStream_t comm = nullptr;    /* NCCL broadcasts / A-panel pulls (high priority) */
  cudaStream_t comm2 = nullptr;   /* B-panel pulls (high priority) */
  cudaStream_t copy = nullptr;    /* H2D / D2H staging */
  cublasHandle_t blas = nullptr;
  unsigned int *sched = nullptr; /* tile-scheduler words, self-resetting */
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  DevBuf bufA, bufB, bufC;
  DevBuf ozA, ozB, ozE; /* Ozaki digit matrices of the current K chunk and the row/column exponents */
  DevBuf ozSync, ozT;   /* wave-sync counters; diagnostics: per-tile timestamps */
  DevBuf ozG;           /* [0] guard of the current K chunk, [1] chunks handed to the native kernel, [2] OR of their reasons */
  cudaEvent_t gemm_done = nullptr; /* recorded after every GEMM launch: launches of one device never overlap (shared scratch) */
  bool gemm_in_flight = false;
  cudaStream_t gemm_last_stream = nullptr;
};

DeviceCtx *phpc_ctx(int device); /* lazily created; makes `device` current */
DeviceCtx *phpc_cur_ctx(void);   /* context of the process' bound device */
void *phpc_buf_reserve(DevBuf *b, size_t bytes);

/* enqueue the DMMA kernel on `stream`; returns launches (0 or 1) */
int phpc_launch_dmma(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                     int k, int n, int ctas, cudaStream_t stream);
/* FP64 GEMM rebuilt from int8 tcgen05 MMAs (Ozaki scheme, 7 balanced base-256 digits per operand; K chunks the guard
 * rejects run on the DMMA kernel); returns the number of kernels launched */
/* Exponents and digits of ONE B chunk (k <= 8192) kept by the caller across calls that multiply different A rows by the same
 * chunk (the row bands of the host-sourced runs): the first call fills it (`ready` false), later calls skip the exponent and
 * split kernels of B.  TB: phpc_ozaki_bcache_bytes(k, n, &ints) bytes, eB: `ints` ints.  All calls on one stream. */
struct OzBCache {
  signed char *TB = nullptr;
  int *eB = nullptr;
  bool ready = false;
};
size_t phpc_ozaki_bcache_bytes(int k, int n, size_t *exp_ints);
int phpc_launch_ozaki(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                      int k, int n, int ctas, cudaStream_t stream, OzBCache *bcache = nullptr);
bool phpc_use_ozaki(void); /* env PHPC_GEMM=ozaki */
void phpc_launch_cublas(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                        int k, int n, cudaStream_t stream);

/* shared host allocations (phpc_host_malloc_shared): owner-side lookup and importer-side mapping, used by the gather */
int phpc_host_shared_lookup(const void *p, char *name, unsigned long long *offset, unsigned long long *bytes);
void *phpc_host_shared_map(const char *name, unsigned long long bytes, unsigned long long off, unsigned long long len);
void phpc_host_shared_release_imports(void);

static inline long long phpc_pad_ld(long long cols) { return (cols + 15) / 16 * 16; }
