/* utils.c — see include/utils.h (replaces reference src/utils.c). */
#include "../../include/utils.h"

#include <sys/time.h>

double get_cur_time(void) {
  struct timeval now;
  gettimeofday(&now, NULL);
  return (double)now.tv_sec + (double)now.tv_usec * 1e-6;
}

void log_to_csv(FILE *csv_file, int N, int size, int gpu_count, int num_blocks, int threads_per_block, double cuda_time,
                float cuda_gpu_time, double cublas_time) {
  if (!csv_file) return;
  fprintf(csv_file, "%d,%d,%d,%d,%d,%d,%f,%f,%f\n", N, size, gpu_count, num_blocks, threads_per_block,
          gpu_count * num_blocks * threads_per_block, cuda_time, cuda_gpu_time, cublas_time);
}
