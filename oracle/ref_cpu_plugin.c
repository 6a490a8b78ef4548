/*
 * ref_cpu_plugin.c — CPU stand-ins for the two gemm_t plugins so that the
 * reference's UNMODIFIED src/phpc_summa.c can run in a container without a GPU
 * (oracle/_ref/ref_summa_cpu.out).  TEST INFRASTRUCTURE ONLY.
 *
 * Same signatures as reference src/phpc_gemm.cuh:8,14; arithmetic follows the
 * reference kernel (src/phpc_gemm.cu:33-55): per-element sum in ascending k into
 * a zero-initialised local, then C += sum.
 */
void oracle_gemm_block(const double *a, long lda, const double *b, long ldb, double *c, long ldc, int m, int k, int n);

void phpc_gemm_cuda(const double *a, int lda, const double *b, int ldb, double *c, int ldc, int m, int k, int n, int gpu_count,
                    int grid_width, int grid_height, int block_width, float *compute_time) {
  (void)gpu_count;
  (void)grid_width;
  (void)grid_height;
  (void)block_width;
  oracle_gemm_block(a, lda, b, ldb, c, ldc, m, k, n);
  *compute_time = 0;
}

void phpc_gemm_cublas(const double *a, int lda, const double *b, int ldb, double *c, int ldc, int m, int k, int n, int gpu_count,
                      int grid_width, int grid_height, int block_width, float *gpu_time) {
  phpc_gemm_cuda(a, lda, b, ldb, c, ldc, m, k, n, gpu_count, grid_width, grid_height, block_width, gpu_time);
}
