/*
 * phpc_summa.h — C-ABI of the SUMMA outer loop (drop-in for the reference's
 * src/phpc_summa.h:6-8; same names, argument order and meaning), plus the
 * device-resident additions the B200 build needs.
 *
 * Semantics kept from reference src/phpc_summa.c:24-122:
 *   - grid_comm is a 2-D Cartesian communicator (r x c); N % r == N % c == 0;
 *   - A, B, C are FULL N x N row-major HOST matrices on every rank; rank (i,j)
 *     owns rows i*N/r.. of A and C, columns j*N/c.. of B and C; the K dimension is
 *     cut into lcm(r,c) panels, A panel k is broadcast along the process row by
 *     column k%c, B panel k along the process column by row k%r (:64-89);
 *   - C += A*B (the local GEMM accumulates, :93); after return rank 0 holds the
 *     full C, every other rank its own block at its global offset (:97-110);
 *   - *compute_time = device seconds of the local GEMMs summed over the steps (:94).
 * What changed underneath: the panels never touch host memory.  Each rank uploads its owned blocks chunk by chunk under
 * the GEMMs of earlier chunks, the k-loop runs on device: the K chunks a rank does not own are pulled straight out of the
 * owner's HBM by the copy engines over NVLink (CUDA IPC; PHPC_PANEL=nccl uses one ncclGroup of ncclBroadcast on the
 * per-row / per-column NCCL communicators instead), multi-buffered on communication streams so the transfer of chunk
 * q+1.. overlaps the local GEMM of chunk q (the tcgen05 kernel of csrc/ozaki_gemm.cuh; PHPC_GEMM=dmma = native FP64),
 * and the C block comes back once at the end.  MPI (real or the single-node shim in mpi_shim/) is only the control plane:
 * NCCL id / IPC handle exchange, barriers, the optional MPI gather.
 */
#ifndef _PHPC_SUMMA_H
#define _PHPC_SUMMA_H

#include <mpi.h>

#ifdef __cplusplus
extern "C" {
#endif

/* replaces reference src/phpc_summa.c:124-126 */
void phpc_gemm_summa_cuda(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int gpu_count, int grid_width,
                          int grid_height, int block_width, float *compute_time);

/* replaces reference src/phpc_summa.c:128-130 (same loop, cuBLAS Dgemm as the local GEMM) */
void phpc_gemm_summa_cublas(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int gpu_count, float *compute_time);

/* ======================= additions (not in the reference) ======================= */

#define PHPC_BACKEND_DMMA 0
#define PHPC_BACKEND_CUBLAS 1
#define PHPC_BACKEND_OZAKI 2 /* FP64 rebuilt from int8 tcgen05 MMAs, see phpc_gemm_device_ozaki */

/* One k-step of the schedule as rank (pi, pj) sees it. */
typedef struct phpc_summa_step {
  int panel;        /* logical SUMMA panel k of reference :63 (0 .. lcm-1) */
  int a_root;       /* process column that broadcasts the A chunk (= panel % c, :64) */
  int b_root;       /* process row that broadcasts the B chunk    (= panel % r, :65) */
  long long k0;     /* global K offset of the chunk */
  int width;        /* K extent of the chunk (<= kc) */
  int own_a, own_b; /* this rank is the root of the A / B broadcast */
  long long a_off;  /* element offset of the chunk inside this rank's A store (own_a) */
  long long b_off;  /* element offset of the chunk inside this rank's B store (own_b) */
} phpc_summa_step;

/* Pure host arithmetic (no GPU, no MPI): the schedule of an N x N SUMMA on an
 * r x c grid for rank (pi, pj) with K chunks of at most kc columns.  Writes up to
 * max_steps entries, returns the number of steps (or -1 if N is not divisible by r
 * and c).  m/n/lda_pad/ldb_pad describe the rank's blocks (may be NULL). */
int phpc_summa_schedule(int N, int r, int c, int pi, int pj, int kc, phpc_summa_step *steps, int max_steps, int *m_out, int *n_out);
/* The same for C[M x N] += A[M x K] * B[K x N]: blocks M/r x N/c, K panels of K/lcm(r,c) (-1 unless M % r == N % c == K % lcm == 0). */
int phpc_summa_schedule_mkn(int M, int K, int N, int r, int c, int pi, int pj, int kc, phpc_summa_step *steps, int max_steps, int *m_out,
                            int *n_out);

/* The schedule multi-rank objects actually run: as above, but the very first chunk (of panel 0) is only kc_first wide
 * (0 = like the others).  Nothing can overlap the transfer of the first chunk, so it is kept short (phpc_summa_chunks). */
int phpc_summa_schedule_first(int M, int K, int N, int r, int c, int pi, int pj, int kc, int kc_first, phpc_summa_step *steps, int max_steps,
                              int *m_out, int *n_out);

/* One operation of the band-pipelined host-sourced run on a 1 x 1 grid (phpc_summa_run_host): the rank's
 * C block is cut into row bands (1/2, 1/4, ... of the block); band b is multiplied over every K chunk into a zeroed
 * block, the caller's C rows (uploaded in the meantime into a side buffer) are added, and the band is downloaded while
 * band b+1 computes, so the C traffic the reference pays around every kernel (src/phpc_gemm.cu:113,121)
 * hides under the GEMMs together with the A/B uploads. */
#define PHPC_HOP_UPLOAD_C 0   /* host C band   -> side buffer band  (copy-in stream)  */
#define PHPC_HOP_UPLOAD_A 1   /* host A window -> A store (band, step)                */
#define PHPC_HOP_UPLOAD_B 2   /* host B window -> B store (step)                      */
#define PHPC_HOP_GEMM 3       /* dC band += A(band, step) * B(step) (compute stream)  */
#define PHPC_HOP_DOWNLOAD_C 4 /* dC band -> host C band             (copy-out stream) */
#define PHPC_HOP_ZERO_C 5     /* dC band = 0                        (compute stream)  */
#define PHPC_HOP_ADD_C 6      /* dC band += side buffer band        (compute stream)  */
typedef struct phpc_host_op {
  int kind;   /* PHPC_HOP_* */
  int stream; /* 0 = copy-in, 1 = compute, 2 = copy-out; operations of one stream run in list order */
  int band;   /* row band (-1 for PHPC_HOP_UPLOAD_B) */
  int step;   /* K chunk = index into the schedule (-1 for the C operations) */
  int row0, rows; /* the band, in rows of the rank's block */
  int ndeps;
  int deps[3]; /* earlier operations on OTHER streams that must have completed */
} phpc_host_op;

/* Pure host arithmetic: the operation list for a block of m rows, nsteps K chunks and at most `bands` row bands
 * (half of what is left each, the last one takes the rest) whose height is rounded up to a multiple of `align` rows.  Writes up to max_ops entries in issue order
 * (dependencies always point backwards) and returns the number of operations. */
int phpc_host_plan(int m, int nsteps, int bands, int align, phpc_host_op *ops, int max_ops);

typedef struct phpc_summa_stats {
  float total_ms;   /* first broadcast enqueued -> last GEMM complete (device events) */
  float gemm_ms;    /* sum of the local GEMM kernel durations */
  float exposed_ms; /* total_ms - gemm_ms: broadcast (and launch) time NOT hidden */
  int steps;
  int launches;      /* GEMM kernels launched */
  int broadcasts;    /* ncclBroadcast calls issued */
  long long bytes_received; /* NVLink bytes this rank received */
} phpc_summa_stats;

typedef struct phpc_summa phpc_summa; /* opaque: blocks in HBM + NCCL row/col communicators */

/* Collective over grid_comm.  Binds the rank to a GPU (PHPC_DEVICE, else
 * LOCAL_RANK, else rank % device_count), builds/caches the NCCL communicators and
 * allocates the rank's A, B, C blocks and the receive ring in HBM.  kc <= 0 picks
 * the default (whole panel on a 1x1 grid, 8192 otherwise; env PHPC_KC overrides).
 * Panel transport (env PHPC_PANEL): "nccl" = ncclBroadcast on the row / column
 * communicators; "pull" (default) = each rank copies the chunks it does not own
 * straight out of the owner's HBM with the copy engines over NVLink (CUDA IPC peer
 * mappings): same data movement as the broadcast, but no SMs and no rendezvous.
 * create / destroy / upload / fill / run_host are collective over grid_comm. */
phpc_summa *phpc_summa_create(MPI_Comm grid_comm, int n, int kc);
/* The same object for a general C[M x N] += A[M x K] * B[K x N] (the reference is square only, src/phpc_summa.c:36-39):
 * M % r == 0, N % c == 0, K % lcm(r,c) == 0.  Rank (i,j) owns rows i*M/r.. of A and C, columns j*N/c.. of B and C, and the K
 * panels of width K/lcm as in the square case.  Host matrices passed to upload / run_host / download_c are the FULL
 * A (M x K, ld K), B (K x N, ld N), C (M x N, ld N).  phpc_summa_create(comm, n, kc) == phpc_summa_create_mkn(comm, n, n, n, kc). */
phpc_summa *phpc_summa_create_mkn(MPI_Comm grid_comm, int m, int k, int n, int kc);
/* K chunking of the object: chunk width, width of the very first chunk (multi-rank grids start with a short chunk, 2048 by
 * default / PHPC_KC_FIRST, because nothing can overlap the transfer of the first chunk; 0 = like the others), number of steps. */
void phpc_summa_chunks(const phpc_summa *s, int *kc, int *kc_first, int *steps);
/* Global problem size {M, K, N} of the object. */
void phpc_summa_global(const phpc_summa *s, int mkn[3]);
void phpc_summa_destroy(phpc_summa *s);
/* Upload the rank's owned blocks from FULL host matrices (C may be NULL = zero). */
void phpc_summa_upload(phpc_summa *s, const double *A, const double *B, const double *C);
/* Generate the owned blocks in HBM (PHPC_FILL_INDEX / PHPC_FILL_SEEDED of
 * phpc_b200.h, seeds for A and B) and zero C: no host matrices at all. */
void phpc_summa_fill(phpc_summa *s, int kind, unsigned long long seed_a, unsigned long long seed_b);
void phpc_summa_zero_c(phpc_summa *s);
/* The device-resident SUMMA k-loop.  `stream` (cudaStream_t or NULL) is the caller's
 * stream: the loop starts after work already enqueued on it and the stream waits for
 * the last GEMM, so events recorded on it bracket the whole step.  Not synchronised
 * unless stats != NULL (stats need the events to complete). */
void phpc_summa_run(phpc_summa *s, int backend, int ctas, void *stream, phpc_summa_stats *stats);
/* Per-step GEMM start offsets (ms after the run began) and durations of the LAST run; returns
 * the number of entries written.  Diagnostics for the exposed-broadcast measurement. */
int phpc_summa_timeline(phpc_summa *s, float *start_ms, float *dur_ms, int max_steps);
/* Host-sourced run (what phpc_gemm_summa_cuda does): C += A*B on FULL N x N host matrices,
 * owned chunks uploaded on a copy stream while earlier chunks compute, C block downloaded
 * (and gathered to rank 0 when gather != 0) at the end.  Synchronous.
 * On a 1 x 1 grid with a block of >= 8192 rows (or PHPC_HOST_BANDS > 1) the C block travels in row
 * bands under the GEMMs instead (phpc_host_plan above); results are bit-identical to the chunk loop.
 * Page-locked host matrices (phpc_host_register) are what lets the transfers overlap; pageable memory
 * works and gives the same results, but the CUDA runtime serialises such copies with the host thread. */
void phpc_summa_run_host(phpc_summa *s, int backend, int ctas, const double *A, const double *B, double *C, int gather,
                         phpc_summa_stats *stats);
/* Drop the device blocks cached by phpc_gemm_summa_cuda / phpc_gemm_summa_cublas. */
void phpc_summa_release_cache(void);
/* D2H of the rank's C block into a FULL N x N host matrix at its global offset, then
 * (gather != 0) the reference's gather to rank 0 (src/phpc_summa.c:97-110). */
void phpc_summa_download_c(phpc_summa *s, double *C, int gather);
/* Copy a rows x cols window of the rank's C block (block-local coordinates) to host. */
void phpc_summa_read_c_block(phpc_summa *s, double *dst, long long ld, int row0, int col0, int rows, int cols);
/* Geometry of the rank: dims {r,c}, coords {pi,pj}, block {m,n}. */
void phpc_summa_geometry(const phpc_summa *s, int dims[2], int coords[2], int block[2]);

#ifdef __cplusplus
}
#endif

#endif
