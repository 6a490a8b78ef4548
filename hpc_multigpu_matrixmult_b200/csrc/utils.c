/*
 * utils.c — wall clock and the one-line CSV record of the harness (see include/utils.h).
 * Replaces reference src/utils.c: same two entry points, same output bytes.
 */
#include "../../include/utils.h"

#include <time.h>

/* Seconds since the epoch as a double.  CLOCK_REALTIME is the clock gettimeofday() reads
 * (reference src/utils.c:13), here with nanosecond resolution. */
double get_cur_time(void) {
  struct timespec ts;
  if (clock_gettime(CLOCK_REALTIME, &ts) != 0) return 0.0;
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/*
 * One record per run, nine comma-separated fields, fixed notation for the three times:
 *   N, ranks, GPUs per rank, CUDA blocks, threads per block, total threads,
 *   wall seconds of the CUDA pass, device seconds of its kernels, wall seconds of the cuBLAS pass
 * total threads = GPUs * blocks * threads per block (reference src/utils.c:26-27).
 * scripts/run_tests_csv.py and the reference's scripts/tests.sh both parse exactly this line.
 */
void log_to_csv(FILE *csv_file, int N, int size, int gpu_count, int num_blocks, int threads_per_block, double cuda_time,
                float cuda_gpu_time, double cublas_time) {
  if (csv_file == NULL) return;
  const int launched_threads = gpu_count * num_blocks * threads_per_block;
  fprintf(csv_file, "%d,%d,%d,", N, size, gpu_count);
  fprintf(csv_file, "%d,%d,%d,", num_blocks, threads_per_block, launched_threads);
  fprintf(csv_file, "%f,%f,%f\n", cuda_time, (double)cuda_gpu_time, cublas_time);
}
