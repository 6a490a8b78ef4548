/* dropin_wrap.c — TEST INFRASTRUCTURE ONLY (tests/test_dropin_link.py).
 *
 * The reference's main.c never checks or outputs C (src/main.c:88-114) and does not zero it (:66), so a link test of the
 * UNMODIFIED reference driver against libphpc_b200.so needs a way to look at the result without touching the reference's
 * sources.  The drop-in binaries are linked with
 *     -Wl,--wrap=phpc_gemm_summa_cuda -Wl,--wrap=phpc_gemm_summa_cublas
 * so the calls in the reference's main.o (src/main.c:94,106) arrive here first: zero C, call the real entry point (the
 * library's own SUMMA in recipe (b) of INTEGRATION.md; the reference's phpc_summa.c with the library's phpc_gemm_cuda as
 * its gemm_t plugin in recipe (a)), then rank 0 writes C to $PHPC_DROPIN_DUMP.{cuda,cublas} for the test to compare with
 * the oracle. */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void __real_phpc_gemm_summa_cuda(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int gpu_count, int grid_width,
                                 int grid_height, int block_width, float *compute_time);
void __real_phpc_gemm_summa_cublas(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int gpu_count, float *compute_time);

static void dump(MPI_Comm comm, const double *C, int n, const char *suffix) {
  int rank;
  MPI_Comm_rank(comm, &rank);
  const char *base = getenv("PHPC_DROPIN_DUMP");
  if (rank != 0 || !base) return;
  char path[512];
  snprintf(path, sizeof path, "%s.%s", base, suffix);
  FILE *f = fopen(path, "wb");
  if (!f || fwrite(C, sizeof(double), (size_t)n * n, f) != (size_t)n * n) {
    fprintf(stderr, "dropin_wrap: cannot write %s\n", path);
    abort();
  }
  fclose(f);
}

void __wrap_phpc_gemm_summa_cuda(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int gpu_count, int grid_width,
                                 int grid_height, int block_width, float *compute_time) {
  memset(C, 0, (size_t)n * n * sizeof(double));
  __real_phpc_gemm_summa_cuda(grid_comm, A, B, C, n, gpu_count, grid_width, grid_height, block_width, compute_time);
  dump(grid_comm, C, n, "cuda");
}

void __wrap_phpc_gemm_summa_cublas(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int gpu_count, float *compute_time) {
  memset(C, 0, (size_t)n * n * sizeof(double));
  __real_phpc_gemm_summa_cublas(grid_comm, A, B, C, n, gpu_count, compute_time);
  dump(grid_comm, C, n, "cublas");
}
