/*
 * phpc_b200.h — ADDITIONS to the reference's C interface (nothing here replaces a
 * reference symbol; phpc_gemm.cuh and phpc_summa.h hold the drop-in entry points
 * and are implemented on top of these).  Plain C ABI: pointers, sizes, ints.
 *
 * Device-resident entry points exist because the reference's host-pointer API
 * (src/phpc_gemm.cu:93-121: pin, cudaMallocAsync, H2D, kernel, D2H, free on
 * EVERY k-step) cannot express "operands already in HBM", which is where a B200
 * SUMMA keeps them (decision D3).
 *
 * All functions abort the process with a message on CUDA/NCCL/cuBLAS errors.
 */
#ifndef _PHPC_B200_H
#define _PHPC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library / device ---------------------------------------------------- */
/* Library ABI version (also proves the .so loads without a GPU). */
int phpc_b200_version(void);
/* Number of visible CUDA devices; 0 when there is no driver/GPU (never aborts). */
int phpc_b200_device_count(void);
/* Bind the calling process to `device` (one process per GPU); creates the
 * per-device context (streams, cuBLAS handle, tile-scheduler words). */
void phpc_b200_set_device(int device);
int phpc_b200_get_device(void);
int phpc_b200_sm_count(void);
/* Release every cached device buffer, stream and handle. */
void phpc_b200_finalize(void);

/* ---- memory -------------------------------------------------------------- */
void *phpc_device_malloc(size_t bytes);
void phpc_device_free(void *p);
void *phpc_host_malloc_pinned(size_t bytes);
void phpc_host_free_pinned(void *p);
/* Page-lock / unlock a caller-owned host range so copies from it are asynchronous DMA. */
void phpc_host_register(void *p, size_t bytes);
void phpc_host_unregister(void *p);
/* Page-locked host memory that every rank of the node can map (POSIX shared memory).  When rank 0's result matrix C lives in
 * such an allocation, the gather of phpc_gemm_summa_cuda / phpc_summa_download_c lets every rank write its C block into it
 * over its OWN PCIe link (all links in parallel) instead of sending all blocks through rank 0's GPU (the reference's serial
 * gather, src/phpc_summa.c:97-110).  Returns NULL when the shared-memory file system cannot hold `bytes` (use malloc then:
 * the result is the same, only the gather is serial). */
void *phpc_host_malloc_shared(size_t bytes);
void phpc_host_free_shared(void *p);
void phpc_device_memset(void *p, int value, size_t bytes);
void phpc_device_synchronize(void);
/* rows x cols doubles between a host matrix (ld_host) and a device matrix (ld_dev); synchronous. */
void phpc_copy2d_to_host(double *host, long long ld_host, const double *dev, long long ld_dev, long long rows, long long cols);
void phpc_copy2d_to_device(double *dev, long long ld_dev, const double *host, long long ld_host, long long rows, long long cols);

/* ---- local block GEMM on device pointers --------------------------------- */
/*
 * dC[m x n, ldc] += dA[m x k, lda] * dB[k x n, ldb] with the sm_100a DMMA kernel,
 * enqueued on `stream` (a cudaStream_t; NULL = the library's compute stream)
 * and NOT synchronised.  GEMM launches of one device are ordered one after the other
 * whatever stream they are given (they share per-device scratch); K is walked in
 * chunks of 4096, one launch each.  dA, dB must be 16-byte aligned with even lda, ldb (TMA
 * global-stride rule); the call aborts otherwise.  `ctas` <= 0 means one
 * persistent CTA per SM.  Returns the number of kernels launched (0 for an
 * empty problem).
 */
int phpc_gemm_device(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m, int k, int n,
                     int ctas, void *stream);
/* Same contraction through cublasDgemm (alpha = beta = 1) on the same stream. */
void phpc_gemm_device_cublas(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m, int k,
                             int n, void *stream);
/*
 * Same contraction on the tcgen05 tensor cores (the default local GEMM of the reference-named entry points): per K chunk of
 * 8192, rows of A and columns of B are scaled by a power of two and written as 7 balanced base-256 digits (54 bits below the
 * row / column maximum, error free up to that rounding), the 28 digit products with t + u <= 8 run as int8 MMAs with exact
 * int32 accumulators in TMEM, and the FP64 result is reassembled in the epilogue (Ozaki scheme).  Error model: normwise per
 * row of A / column of B, |dC_ij| <~ k * 2^-55 * max|A_i,:| * max|B_:,j| plus two FP64 roundings per chunk.  A guard computed
 * with the exponents sends a K chunk to the native-FP64 DMMA kernel instead when it holds an Inf/NaN, when exponents come
 * near the FP64 range limits (underflow / overflow then behave as in FP64), or when the nonzero entries of one row / column
 * span more than 2^40 — so special values propagate exactly as in the reference.  Returns the number of kernels launched.
 */
int phpc_gemm_device_ozaki(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m, int k,
                           int n, void *stream);
/* K chunks the guard handed to the native-FP64 kernel since the previous call (synchronises the device). */
long long phpc_ozaki_fallback_chunks(void);
/* Run phpc_gemm_device `reps` times back to back and return the mean device
 * milliseconds per launch (CUDA events on the launching stream). */
float phpc_gemm_device_timed(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m, int k,
                             int n, int ctas, int reps, int backend /* 0 DMMA, 1 cuBLAS, 2 Ozaki */);

/* ---- synthetic inputs (device side) --------------------------------------- */
/*
 * Fill the rows x cols window whose top-left element is global (row0, col0) of
 * an N x N matrix, stored with leading dimension ld at d.
 *   PHPC_FILL_INDEX   d[r][c] = (double)((row0+r)*N + (col0+c))   (reference
 *                     src/main.c:85-86, src/iterative.c:30-31)
 *   PHPC_FILL_SEEDED  uniform in (-1,1) from splitmix64(seed, global flat index):
 *                     regenerable on any rank, host or device (oracle/fill.py).
 */
#define PHPC_FILL_INDEX 0
#define PHPC_FILL_SEEDED 1
void phpc_fill_device(double *d, long long ld, long long rows, long long cols, long long row0, long long col0, long long N, int kind,
                      unsigned long long seed, void *stream);
/* Host version of the same generators (used by main.out and tests). */
void phpc_fill_host(double *h, long long ld, long long rows, long long cols, long long row0, long long col0, long long N, int kind,
                    unsigned long long seed);

/* The local GEMM the reference-named entry points (phpc_gemm_cuda, phpc_gemm_summa_cuda) run in this process, as a backend
 * number for phpc_summa_run: PHPC_BACKEND_OZAKI (tcgen05, default) or PHPC_BACKEND_DMMA (environment PHPC_GEMM=dmma). */
int phpc_default_backend(void);

/* The fixed arithmetic of the tcgen05 path: digits per operand (7), int8 digit products per FP64 product (28), K chunk (8192)
 * and the largest exponent spread inside one row / column the guard accepts (40).  Any pointer may be NULL. */
void phpc_ozaki_config(int *digits, int *products, int *k_chunk, int *max_spread);

/* ---- diagnostics (tools/ozaki_knobs.py; environment PHPC_OZ_TSTAMP=1, PHPC_OZ_FLAGS) ---- */
/* Per tile of the last tcgen05 launch: globaltimer ns at the start of its loads [2*tile] and the end of its epilogue [2*tile+1]. */
long long phpc_oz_tstamp_read(unsigned long long *out, long long max_tiles);

#ifdef __cplusplus
}
#endif

#endif
