/*
 * phpc_gemm.cuh — C-ABI of the local block GEMM (drop-in for the reference's
 * src/phpc_gemm.cuh:4-14; same names, argument order and meaning).
 *
 * Both functions compute, on HOST pointers, the accumulating row-major product
 *     c[m x n, ldc] += a[m x k, lda] * b[k x n, ldb]          (FP64)
 * and return only when the result is in host `c` (synchronous, like the
 * reference).  Leading dimensions are in elements.  They are the `gemm_t`
 * plugin the SUMMA loop calls once per k-step (reference src/phpc_summa.c:7,93).
 *
 * Error behaviour: the reference ignores every CUDA status; this library never
 * returns a silently wrong result — any CUDA/cuBLAS failure, a missing GPU or a
 * missing sm_100a kernel image prints "phpc: ... failed at file:line" to stderr
 * and aborts the process.  There is no CPU fallback.
 */
#ifndef _PHPC_GEMM_CUH
#define _PHPC_GEMM_CUH

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Replaces reference src/phpc_gemm.cu:59-156 (phpc_gemm_cuda).
 *   gpu_count     local GPUs to split the n columns over (device 0..gpu_count-1,
 *                 dev_n = n/g + (gpu < n%g), as reference :98); 1 in the
 *                 one-rank-per-GPU deployment.
 *   grid_width x grid_height
 *                 number of persistent CTAs requested; 0 or 1 in total (the
 *                 reference's default 1x1) means "one per SM"; clamped to the
 *                 SM count.
 *   block_width   the reference's tile edge; accepted for CLI/CSV
 *                 compatibility, every value runs the same 128x128 tiles.
 *   Kernel: the tcgen05/TMEM kernel (FP64 rebuilt exactly from int8 MMAs, see
 *   phpc_gemm_device_ozaki in phpc_b200.h) unless the environment says
 *   PHPC_GEMM=dmma (native-FP64 DMMA kernel).  grid_width x grid_height only
 *   applies to the DMMA kernel.
 *   compute_time  out: device seconds of the GEMM kernel(s), mean over the local
 *                 GPUs (reference :145).
 */
void phpc_gemm_cuda(const double *a, int lda, const double *b, int ldb, double *c, int ldc, int m, int k, int n, int gpu_count,
                    int grid_width, int grid_height, int block_width, float *compute_time);

/*
 * Replaces reference src/phpc_gemm.cu:158-174 (phpc_gemm_cublas): the same
 * contraction through cuBLAS Dgemm (column-major trick n,m,k,b,a as reference
 * :170), alpha = beta = 1.  Comparison baseline and device-side oracle.
 * *gpu_time is set to 0 exactly like the reference (:173).
 */
void phpc_gemm_cublas(const double *a, int lda, const double *b, int ldb, double *c, int ldc, int m, int k, int n, int gpu_count,
                      int grid_width, int grid_height, int block_width, float *gpu_time);

#ifdef __cplusplus
}
#endif

#endif
