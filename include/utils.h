/* utils.h — timing and CSV helpers of the harness (drop-in for reference src/utils.h:6-7). */
#ifndef _PHPC_UTILS_H
#define _PHPC_UTILS_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Wall clock in seconds (reference src/utils.c:9-18, gettimeofday resolution). */
double get_cur_time(void);

/* Appends the reference's one-line record (src/utils.c:26-27), byte for byte:
 *   N,size,gpu_count,num_blocks,threads_per_block,total_threads,cuda_time,cuda_gpu_time,cublas_time
 * with total_threads = gpu_count*num_blocks*threads_per_block and "%f" for the times. */
void log_to_csv(FILE *csv_file, int N, int size, int gpu_count, int num_blocks, int threads_per_block, double cuda_time,
                float cuda_gpu_time, double cublas_time);

#ifdef __cplusplus
}
#endif

#endif
