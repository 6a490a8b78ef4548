/*
 * main.c — `main.out <matrix_size> <tile_width> <grid_width> <grid_height> <test_name>`
 *
 * Drop-in for the reference driver (src/main.c:17-123): same five positional
 * arguments, same abort conditions and messages, same CSV file name and record.
 * Times the SUMMA with the library's default local GEMM (the tcgen05 kernel; PHPC_GEMM=dmma selects native FP64), then
 * the same SUMMA with cuBLAS Dgemm.
 *
 * Deliberate differences:
 *   - C is zeroed before each pass (the reference accumulates the cuBLAS pass on
 *     top of the CUDA pass and never checks results, src/main.c:66,94,106);
 *   - one rank drives ONE GPU (rank -> LOCAL_RANK -> device); the CSV gpu_count
 *     column is therefore 1 per rank instead of "every visible device";
 *   - PHPC_MODE=device (default when N >= 16384): no N x N host matrices at all;
 *     each rank generates its owned blocks in HBM with the reference's fill
 *     A[i] = B[i] = i (src/main.c:85-86) and C stays distributed.  PHPC_MODE=host
 *     keeps the reference's "full A, B, C on every rank" through the host entry
 *     points phpc_gemm_summa_cuda / phpc_gemm_summa_cublas.
 *   - PHPC_PGRID=RxC overrides MPI_Dims_create (which gives 8 -> 4x2; BASELINE asks
 *     for 2x4); PHPC_FILL=seeded switches the synthetic input; PHPC_VERIFY=1 checks
 *     sampled elements of C against the closed form of the index fill;
 *   - a JSON sidecar next to the CSV carries TFLOP/s, GEMM and exposed-broadcast
 *     times (the 9-column CSV record is unchanged).
 */
#include <math.h>
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/phpc_b200.h"
#include "../../include/phpc_summa.h"
#include "../../include/utils.h"

#define MPI_ASSERT(check)                                              \
  if (!(check)) {                                                      \
    fprintf(stderr, "Check at " __FILE__ " line %d failed", __LINE__); \
    MPI_Abort(MPI_COMM_WORLD, EXIT_FAILURE);                           \
  }

/* closed form of C = A*B for A[i] = B[i] = i, exact in 128-bit integers */
static double index_fill_product(long long i, long long j, long long N) {
  const __int128 n = N, s1 = n * (n - 1) / 2, s2 = (n - 1) * n * (2 * n - 1) / 6;
  return (double)((__int128)i * n * n * s1 + (__int128)i * j * n * n + n * s2 + (__int128)j * s1);
}

static double verify_block(phpc_summa *s, int N, double scale) {
  int dims[2], coords[2], block[2];
  phpc_summa_geometry(s, dims, coords, block);
  double worst = 0.0;
  enum { SAMPLES = 64 };
  double row[1];
  unsigned long long state = 0x1234567ull + (unsigned long long)coords[0] * 7919 + coords[1];
  for (int t = 0; t < SAMPLES; ++t) {
    state = state * 6364136223846793005ull + 1442695040888963407ull;
    const int r = (int)((state >> 33) % (unsigned long long)block[0]);
    state = state * 6364136223846793005ull + 1442695040888963407ull;
    const int c = (int)((state >> 33) % (unsigned long long)block[1]);
    phpc_summa_read_c_block(s, row, 1, r, c, 1, 1);
    const double want = scale * index_fill_product((long long)coords[0] * block[0] + r, (long long)coords[1] * block[1] + c, N);
    const double err = fabs(row[0] - want) / (fabs(want) > 0 ? fabs(want) : 1.0);
    if (err > worst) worst = err;
  }
  return worst;
}

int main(int argc, char *argv[]) {
  int dims[2], period[2], coord[2], rank, size;
  double start_time, cuda_time, cublas_time;

  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);

  if (argc != 6) {
    if (rank == 0) fprintf(stderr, "Usage: %s <matrix_size> <tile_width> <grid_width> <grid_height> <test_name>\n", argv[0]);
    MPI_Abort(MPI_COMM_WORLD, EXIT_FAILURE);
  }
  const int N = atoi(argv[1]);
  const int tile_width = atoi(argv[2]);
  const int grid_width = atoi(argv[3]);
  const int grid_height = atoi(argv[4]);
  const char *test_name = argv[5];

  dims[0] = dims[1] = 1;
  const char *pgrid = getenv("PHPC_PGRID");
  if (pgrid && sscanf(pgrid, "%dx%d", &dims[0], &dims[1]) == 2) {
    MPI_ASSERT(dims[0] * dims[1] == size);
  } else if (size > 1) {
    dims[0] = dims[1] = 0;
    MPI_Dims_create(size, 2, dims);
  }
  if (N <= 0 || N % dims[0] != 0 || N % dims[1] != 0) {
    if (rank == 0)
      fprintf(stderr, "Error: Matrix size N (%d) must be divisible by process grid dimensions (%d x %d).\n", N, dims[0], dims[1]);
    MPI_Abort(MPI_COMM_WORLD, EXIT_FAILURE);
  }

  MPI_ASSERT(phpc_b200_device_count() > 0);
  const int gpu_count = 1; /* GPUs driven by one rank */
  { /* bind the rank to its GPU (and to that GPU's NUMA node) BEFORE the host matrices are allocated and first touched */
    const char *lr = getenv("PHPC_DEVICE") ? getenv("PHPC_DEVICE") : getenv("LOCAL_RANK");
    phpc_b200_set_device((lr ? atoi(lr) : rank) % phpc_b200_device_count());
  }

  period[0] = period[1] = 1;
  MPI_Comm grid_comm;
  MPI_Cart_create(MPI_COMM_WORLD, 2, dims, period, 0, &grid_comm);
  MPI_Cart_coords(grid_comm, rank, 2, coord);

  const char *mode_env = getenv("PHPC_MODE");
  const int device_mode = mode_env ? !strcmp(mode_env, "device") : (N >= 16384);
  const char *fill_env = getenv("PHPC_FILL");
  const int fill = (fill_env && !strcmp(fill_env, "seeded")) ? PHPC_FILL_SEEDED : PHPC_FILL_INDEX;
  const int verify = getenv("PHPC_VERIFY") && atoi(getenv("PHPC_VERIFY"));

  FILE *csv_file = NULL;
  char filename[256];
  if (rank == 0) {
    snprintf(filename, sizeof filename, "csv/%s_N%d_T%d_G%d_TW%d_GW%d_GH%d.csv", test_name, N, size, gpu_count, tile_width, grid_width,
             grid_height);
    csv_file = fopen(filename, "w");
    if (csv_file == NULL) {
      fprintf(stderr, "Error: Could not create CSV file %s\n", filename);
      MPI_Abort(MPI_COMM_WORLD, EXIT_FAILURE);
    }
  }

  float cuda_gpu_time = 0.f, cublas_gpu_time = 0.f;
  double worst_err = -1.0, exposed_ms = 0.0, total_ms = 0.0;

  if (device_mode) {
    phpc_summa *s = phpc_summa_create(grid_comm, N, 0);
    phpc_summa_fill(s, fill, 1234, 5678);
    phpc_summa_stats st;

    MPI_Barrier(MPI_COMM_WORLD);
    start_time = get_cur_time();
    phpc_summa_run(s, phpc_default_backend(), grid_width * grid_height, NULL, &st);
    MPI_Barrier(MPI_COMM_WORLD);
    cuda_time = get_cur_time() - start_time;
    cuda_gpu_time = st.gemm_ms / 1000.f;
    exposed_ms = st.exposed_ms;
    total_ms = st.total_ms;
    if (verify && fill == PHPC_FILL_INDEX) worst_err = verify_block(s, N, 1.0);

    phpc_summa_zero_c(s);
    MPI_Barrier(MPI_COMM_WORLD);
    start_time = get_cur_time();
    phpc_summa_run(s, PHPC_BACKEND_CUBLAS, 0, NULL, &st);
    MPI_Barrier(MPI_COMM_WORLD);
    cublas_time = get_cur_time() - start_time;
    phpc_summa_destroy(s);
  } else {
    const size_t elems = (size_t)N * (size_t)N;
    double *A = (double *)malloc(elems * sizeof(double));
    double *B = (double *)malloc(elems * sizeof(double));
    /* rank 0's C receives every block: in node-shared page-locked memory the gather runs over all PCIe links at once */
    double *C = (rank == 0 && size > 1) ? (double *)phpc_host_malloc_shared(elems * sizeof(double)) : NULL;
    const int c_shared = C != NULL;
    if (!C) C = (double *)malloc(elems * sizeof(double));
    MPI_ASSERT(A != NULL);
    MPI_ASSERT(B != NULL);
    MPI_ASSERT(C != NULL);
    phpc_fill_host(A, N, N, N, 0, 0, N, fill, 1234);
    phpc_fill_host(B, N, N, N, 0, 0, N, fill, 5678);

    memset(C, 0, elems * sizeof(double));
    MPI_Barrier(MPI_COMM_WORLD);
    start_time = get_cur_time();
    phpc_gemm_summa_cuda(grid_comm, A, B, C, N, gpu_count, grid_width, grid_height, tile_width, &cuda_gpu_time);
    cuda_time = get_cur_time() - start_time;
    if (verify && fill == PHPC_FILL_INDEX && rank == 0) {
      worst_err = 0.0;
      for (size_t t = 0; t < 4096; ++t) {
        const size_t i = (t * 2654435761u) % (size_t)N, j = (t * 40503u + 17) % (size_t)N;
        const double want = index_fill_product((long long)i, (long long)j, N);
        const double err = fabs(C[i * N + j] - want) / (fabs(want) > 0 ? fabs(want) : 1.0);
        if (err > worst_err) worst_err = err;
      }
    }

    memset(C, 0, elems * sizeof(double));
    MPI_Barrier(MPI_COMM_WORLD);
    start_time = get_cur_time();
    phpc_gemm_summa_cublas(grid_comm, A, B, C, N, gpu_count, &cublas_gpu_time);
    cublas_time = get_cur_time() - start_time;
    free(A);
    free(B);
    if (c_shared)
      phpc_host_free_shared(C);
    else
      free(C);
  }

  /* mean kernel time over the ranks (reference src/main.c:97-98) */
  MPI_Reduce(rank == 0 ? MPI_IN_PLACE : &cuda_gpu_time, &cuda_gpu_time, 1, MPI_FLOAT, MPI_SUM, 0, MPI_COMM_WORLD);
  cuda_gpu_time /= size;
  MPI_Reduce(rank == 0 ? MPI_IN_PLACE : &worst_err, &worst_err, 1, MPI_DOUBLE, MPI_MAX, 0, MPI_COMM_WORLD);

  if (rank == 0) {
    log_to_csv(csv_file, N, size, gpu_count, grid_width * grid_height, tile_width * tile_width, cuda_time, cuda_gpu_time, cublas_time);
    fclose(csv_file);
    char sidecar[300];
    snprintf(sidecar, sizeof sidecar, "%s.json", filename);
    FILE *js = fopen(sidecar, "w");
    if (js) {
      const double flops = 2.0 * N * (double)N * N;
      fprintf(js,
              "{\"N\": %d, \"ranks\": %d, \"grid\": \"%dx%d\", \"mode\": \"%s\", \"fill\": \"%s\", \"cuda_time_s\": %.6f, "
              "\"cuda_gpu_time_s\": %.6f, \"cublas_time_s\": %.6f, \"tflops_wall\": %.3f, \"tflops_cublas_wall\": %.3f, "
              "\"summa_total_ms\": %.3f, \"exposed_ms\": %.3f, \"max_rel_err\": %.3e}\n",
              N, size, dims[0], dims[1], device_mode ? "device" : "host", fill == PHPC_FILL_INDEX ? "index" : "seeded", cuda_time,
              cuda_gpu_time, cublas_time, flops / cuda_time / 1e12, flops / cublas_time / 1e12, total_ms, exposed_ms, worst_err);
      fclose(js);
    }
    if (verify && worst_err > 1e-12) {
      fprintf(stderr, "Error: verification failed, max relative error %.3e\n", worst_err);
      MPI_Abort(MPI_COMM_WORLD, EXIT_FAILURE);
    }
  }

  phpc_summa_release_cache();
  phpc_b200_finalize();
  MPI_Finalize();
  return 0;
}
