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
void phpc_device_memset(void *p, int value, size_t bytes);
void phpc_device_synchronize(void);
/* rows x cols doubles between a host matrix (ld_host) and a device matrix (ld_dev); synchronous. */
void phpc_copy2d_to_host(double *host, long long ld_host, const double *dev, long long ld_dev, long long rows, long long cols);
void phpc_copy2d_to_device(double *dev, long long ld_dev, const double *host, long long ld_host, long long rows, long long cols);

/* ---- local block GEMM on device pointers --------------------------------- */
/*
 * dC[m x n, ldc] += dA[m x k, lda] * dB[k x n, ldb] with the sm_100a DMMA kernel,
 * enqueued on `stream` (a cudaStream_t; NULL = the library's compute stream)
 * and NOT synchronised.  dA, dB must be 16-byte aligned with even lda, ldb (TMA
 * global-stride rule); the call aborts otherwise.  `ctas` <= 0 means one
 * persistent CTA per SM.  Returns the number of kernels launched (1, or 0 for an
 * empty problem).
 */
int phpc_gemm_device(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m, int k, int n,
                     int ctas, void *stream);
/* Same contraction through cublasDgemm (alpha = beta = 1) on the same stream. */
void phpc_gemm_device_cublas(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m, int k,
                             int n, void *stream);
/*
 * Same contraction on the tcgen05 tensor cores: A and B are cut into `slices` signed 7-bit
 * digit matrices (error-free, per-row / per-column power-of-two scaling), the digit products run
 * as int8 MMAs with exact int32 accumulators in TMEM, and the FP64 result is reassembled in the
 * epilogue (Ozaki scheme; slices <= 0 = PHPC_OZAKI_SLICES or 8, i.e. 56 bits below each row /
 * column maximum).  Inputs must be finite.  Returns the number of kernels launched.
 */
int phpc_gemm_device_ozaki(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m, int k,
                           int n, int slices, void *stream);
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

/* The configuration the tcgen05 (Ozaki) path of this process runs with (environment PHPC_OZAKI_DIGITS / PHPC_OZAKI_KERNEL /
 * PHPC_OZAKI_SLICES, else the built-in defaults): digits per operand, int8 digit products per FP64 product
 * (digits*(digits+1)/2), kernel 0 = 1-CTA, 1 = 2-CTA (relay), 2 = 2-CTA (tensor-map loads), balanced = 1 for balanced
 * base-256 digits (0: truncated 7-bit digits).  Any pointer may be NULL. */
void phpc_ozaki_config(int *digits, int *products, int *kernel, int *balanced);

/* ---- bring-up diagnostics of the experimental 2-CTA kernel (PHPC_OZAKI_KERNEL=2cta, PHPC_OZ_PROGRESS=1) ---- */
/* 1 when everything enqueued on the library's compute stream has finished, 0 while work is pending (never blocks). */
int phpc_compute_stream_idle(void);
/* Copies the host-mapped progress words (8 per CTA: producer, MMA issuer / relay, 4 epilogue warps, set-up) of the last
 * 2-CTA launch; works WHILE that kernel runs or hangs.  Returns the number of words written (0 when not enabled). */
int phpc_oz_progress_read(unsigned int *out, int max_words);

#ifdef __cplusplus
}
#endif

#endif
