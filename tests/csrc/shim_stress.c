/* shim_stress.c — hammers the MPI shim the way the SUMMA control plane uses it: broadcasts with a
 * rotating root, barriers, all-reduces, a strided-datatype gather to rank 0 (reference
 * src/phpc_summa.c:53-59,97-110) and the MPI_IN_PLACE reduce of reference src/main.c:97.
 * Exit code 0 iff every rank saw consistent data.  Built and run by tests/test_multiprocess_cpu.py. */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

int main(int argc, char **argv) {
  int rank, size;
  MPI_Init(&argc, &argv);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  int dims[2] = {0, 0}, per[2] = {1, 1}, coords[2];
  if (size == 1) dims[0] = dims[1] = 1; else MPI_Dims_create(size, 2, dims);
  MPI_Comm grid, row, col;
  MPI_Cart_create(MPI_COMM_WORLD, 2, dims, per, 0, &grid);
  MPI_Cart_coords(grid, rank, 2, coords);
  int keep_row[2] = {0, 1}, keep_col[2] = {1, 0};
  MPI_Cart_sub(grid, keep_row, &row);
  MPI_Cart_sub(grid, keep_col, &col);
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  long bad = 0;
  for (int it = 0; it < iters; ++it) {
    int v = (rank == it % size) ? it : -1;
    MPI_Bcast(&v, 1, MPI_INT, it % size, grid);
    if (v != it) ++bad;
    int rv = (coords[1] == it % dims[1]) ? it + coords[0] : -1; /* row broadcast, root = column it % c */
    MPI_Bcast(&rv, 1, MPI_INT, it % dims[1], row);
    if (rv != it + coords[0]) ++bad;
    int cv = (coords[0] == it % dims[0]) ? it + coords[1] : -1; /* column broadcast, root = row it % r */
    MPI_Bcast(&cv, 1, MPI_INT, it % dims[0], col);
    if (cv != it + coords[1]) ++bad;
    MPI_Barrier(grid);
    if (rank == it % size && (it % 97) == 0) usleep(200);
    double x = rank + it, y = 0;
    MPI_Allreduce(&x, &y, 1, MPI_DOUBLE, MPI_SUM, grid);
    if (y != (double)size * it + size * (size - 1) / 2.0) ++bad;
  }
  /* strided gather of an (m x n) block per rank into rank 0's N x N matrix */
  const int m = 6, n = 5, N0 = dims[0] * m, N1 = dims[1] * n;
  double *C = calloc((size_t)N0 * N1, sizeof(double));
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) C[(coords[0] * m + i) * N1 + coords[1] * n + j] = 1000.0 * rank + i * n + j;
  MPI_Datatype blk;
  MPI_Type_vector(m, n, N1, MPI_DOUBLE, &blk);
  MPI_Type_commit(&blk);
  if (rank == 0) {
    for (int r = 1; r < size; ++r) {
      int co[2];
      MPI_Cart_coords(grid, r, 2, co);
      MPI_Recv(C + (co[0] * m) * N1 + co[1] * n, 1, blk, r, 0, grid, MPI_STATUS_IGNORE);
    }
    for (int r = 0; r < size; ++r) {
      int co[2];
      MPI_Cart_coords(grid, r, 2, co);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j)
          if (C[(co[0] * m + i) * N1 + co[1] * n + j] != 1000.0 * r + i * n + j) ++bad;
    }
  } else {
    MPI_Send(C + (coords[0] * m) * N1 + coords[1] * n, 1, blk, 0, 0, grid);
  }
  MPI_Type_free(&blk);
  float t = (float)(rank + 1);
  MPI_Reduce(rank == 0 ? MPI_IN_PLACE : &t, &t, 1, MPI_FLOAT, MPI_SUM, 0, MPI_COMM_WORLD);
  if (rank == 0 && t != size * (size + 1) / 2.0f) ++bad;
  MPI_Comm_free(&row);
  MPI_Comm_free(&col);
  free(C);
  if (bad) fprintf(stderr, "rank %d: %ld inconsistencies\n", rank, bad);
  MPI_Finalize();
  return bad != 0;
}
