/*
 * ref_summa_driver.c — runs the reference's own phpc_gemm_summa_cuda (reference
 * src/phpc_summa.c, compiled unchanged) under the MPI shim and dumps rank 0's
 * gathered C.  TEST INFRASTRUCTURE ONLY (pins oracle_summa and the shim).
 *
 *   mpirun -n P ref_summa_cpu.out <N> <fill: 0 index | 1 seeded> <out.bin | ->
 *
 * Prints "N,P,r,c,seconds" on rank 0: seconds = wall time of the SUMMA call between two barriers (what
 * bench.py reports as the reference's SUMMA on all host cores); "-" skips the dump of C.
 *
 * Process grid exactly as reference src/main.c:38-62 (MPI_Dims_create, periodic
 * Cartesian grid, reorder 0); C is zeroed first (the reference's main.c does not,
 * SURVEY F6 — without it no result check is possible).
 */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "phpc_summa.h"

void oracle_fill(double *h, long ld, long rows, long cols, long long row0, long long col0, long long N, int kind, unsigned long long seed);

int main(int argc, char **argv) {
  int rank, size, dims[2] = {1, 1}, period[2] = {1, 1};
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  if (argc != 4) {
    if (rank == 0) fprintf(stderr, "Usage: %s <N> <fill> <out.bin>\n", argv[0]);
    MPI_Abort(MPI_COMM_WORLD, 1);
  }
  const int N = atoi(argv[1]), fill = atoi(argv[2]);
  if (size > 1) {
    dims[0] = dims[1] = 0;
    MPI_Dims_create(size, 2, dims);
  }
  if (N % dims[0] || N % dims[1]) MPI_Abort(MPI_COMM_WORLD, 1);
  MPI_Comm grid;
  MPI_Cart_create(MPI_COMM_WORLD, 2, dims, period, 0, &grid);
  double *A = malloc(sizeof(double) * N * N), *B = malloc(sizeof(double) * N * N), *C = calloc((size_t)N * N, sizeof(double));
  oracle_fill(A, N, N, N, 0, 0, N, fill, 1234);
  oracle_fill(B, N, N, N, 0, 0, N, fill, 5678);
  float t = 0;
  struct timespec t0, t1;
  MPI_Barrier(MPI_COMM_WORLD);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  phpc_gemm_summa_cuda(grid, A, B, C, N, 1, 1, 1, 32, &t);
  MPI_Barrier(MPI_COMM_WORLD);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (rank == 0) {
    if (strcmp(argv[3], "-")) {
      FILE *f = fopen(argv[3], "wb");
      if (!f || fwrite(C, sizeof(double), (size_t)N * N, f) != (size_t)N * N) MPI_Abort(MPI_COMM_WORLD, 1);
      fclose(f);
    }
    printf("%d,%d,%d,%d,%.6f\n", N, size, dims[0], dims[1], (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
  }
  free(A);
  free(B);
  free(C);
  MPI_Finalize();
  return 0;
}
