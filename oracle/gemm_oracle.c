/*
 * gemm_oracle.c — CPU restatement of the reference's SUMMA GEMM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hpc_multigpu_matrixmult_b200/ may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity pin: the reference ships no golden vectors or result checks
 * (SURVEY.md section 4).  This restatement is pinned instead against outputs of
 * the reference itself run in this container: oracle/_ref/iterative_dump.out
 * (reference src/iterative.c compiled unchanged, main renamed, C dumped) and
 * oracle/_ref/ref_summa_cpu.out (reference src/phpc_summa.c compiled unchanged
 * over the MPI shim with a CPU gemm_t plugin); see oracle/Makefile and
 * tests/test_oracle_pin.py, and the committed fixtures in tests/golden/.
 *
 * All citations are to /root/reference/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* src/iterative.c:8-13 — loops i (row) -> j (contraction) -> k (column), the
 * statement C[i*n+k] += A[i*n+j]*B[j*n+k]; every C element is therefore summed
 * in ascending contraction order starting from its initial value. */
void oracle_gemm_iterative(const double *A, const double *B, double *C, int n) {
  const size_t N = (size_t)n;
  for (size_t i = 0; i < N; ++i)
    for (size_t j = 0; j < N; ++j) {
      const double a = A[i * N + j];
      const double *brow = B + j * N;
      double *crow = C + i * N;
      for (size_t k = 0; k < N; ++k) crow[k] += a * brow[k];
    }
}

/* src/phpc_gemm.cu:6-57 (gemm_kernel) seen through phpc_gemm_cuda's staging
 * (:111-121): c[m x n, ldc] += a[m x k, lda] * b[k x n, ldb].  The kernel sums
 * each element into a zero-initialised local in ascending k (zero-padded phases
 * add exact zeros, :38-51) and only then does C += c_value (:54-55), so the
 * rounding differs from oracle_gemm_iterative when C starts non-zero. */
void oracle_gemm_block(const double *a, long lda, const double *b, long ldb, double *c, long ldc, int m, int k, int n) {
  double *acc = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (long i = 0; i < m; ++i) {
    memset(acc, 0, sizeof(double) * (size_t)n);
    for (long p = 0; p < k; ++p) {
      const double av = a[i * lda + p];
      const double *brow = b + p * ldb;
      for (long j = 0; j < n; ++j) acc[j] += av * brow[j];
    }
    for (long j = 0; j < n; ++j) c[i * ldc + j] += acc[j];
  }
  free(acc);
}

/* src/phpc_summa.c:9-22 */
int oracle_find_lcm(int a, int b) {
  int x = a, y = b;
  while (y != 0) {
    const int r = x % y;
    x = y;
    y = r;
  }
  return a * b / x;
}

/*
 * src/phpc_summa.c:24-122 for an r x c process grid, all ranks simulated one
 * after the other in this process on the FULL N x N host matrices (every rank
 * of the reference holds full A and B, src/main.c:64-86).  For rank (pr, pc):
 *   lcm = lcm(r, c); local_A_rows = N/r; panel_K = N/lcm; local_B_cols = N/c    (:36-39)
 *   step s uses global K panel s: A[pr*local_A_rows .., s*panel_K ..] and
 *   B[s*panel_K .., pc*local_B_cols ..]  — the owner column s%c / owner row s%r
 *   (:64-65) walk their pointers so that the broadcast block IS global panel s
 *   (:42-43, :74, :84), which is what every receiver multiplies (:93).
 *   C block of the rank += panel product, once per step (:93), with the
 *   gemm_kernel rounding (sum, then +=).
 * The gather to rank 0 (:97-110) places each block at its global offset: the
 * result is simply the full C.
 */
void oracle_summa(const double *A, const double *B, double *C, int N, int r, int c) {
  const int lcm = oracle_find_lcm(r, c);
  const int rows = N / r, pk = N / lcm, cols = N / c;
  for (int pr = 0; pr < r; ++pr)
    for (int pc = 0; pc < c; ++pc)
      for (int s = 0; s < lcm; ++s) {
        const double *a = A + (size_t)pr * rows * N + (size_t)s * pk;
        const double *b = B + (size_t)s * pk * N + (size_t)pc * cols;
        double *cc = C + (size_t)pr * rows * N + (size_t)pc * cols;
        oracle_gemm_block(a, N, b, N, cc, N, rows, pk, cols);
      }
}

/* Ownership tables of the schedule above, for tests of the NCCL plan:
 * owner_col[s] = s % c broadcasts A panel s along its process row, owner_row[s]
 * = s % r broadcasts B panel s along its process column (src/phpc_summa.c:64-65). */
void oracle_summa_owners(int r, int c, int *owner_col, int *owner_row) {
  const int lcm = oracle_find_lcm(r, c);
  for (int s = 0; s < lcm; ++s) {
    owner_col[s] = s % c;
    owner_row[s] = s % r;
  }
}

/*
 * Exact product for the reference's fill A[i] = B[i] = (double)i (src/main.c:85-86,
 * src/iterative.c:30-31): C[i][j] = sum_p (i*N+p)*(p*N+j)
 *   = i*N^2*S1 + i*j*N^2 + N*S2 + j*S1,  S1 = N(N-1)/2, S2 = (N-1)N(2N-1)/6,
 * evaluated in 128-bit integers and rounded once to double (round-to-nearest-even
 * by the int128 -> double conversion).  For N < 1552 every partial sum is an
 * integer below 2^53, so any summation order gives exactly this value.
 */
double oracle_index_fill_exact(long long i, long long j, long long N) {
  const __int128 n = N, S1 = n * (n - 1) / 2, S2 = (n - 1) * n * (2 * n - 1) / 6;
  const __int128 v = (__int128)i * n * n * S1 + (__int128)i * j * n * n + n * S2 + (__int128)j * S1;
  return (double)v;
}

void oracle_index_fill_exact_block(double *out, long ld, long long row0, long long col0, long rows, long cols, long long N) {
  for (long r = 0; r < rows; ++r)
    for (long c = 0; c < cols; ++c) out[r * ld + c] = oracle_index_fill_exact(row0 + r, col0 + c, N);
}

/* Same generators as the product's phpc_fill_host/phpc_fill_device (restated, not
 * shared): index fill and splitmix64-seeded uniform(-1,1). */
double oracle_seeded_value(unsigned long long seed, unsigned long long flat) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (flat + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

void oracle_fill(double *h, long ld, long rows, long cols, long long row0, long long col0, long long N, int kind, unsigned long long seed) {
  for (long r = 0; r < rows; ++r)
    for (long c = 0; c < cols; ++c) {
      const unsigned long long flat = (unsigned long long)((row0 + r) * N + (col0 + c));
      h[r * ld + c] = kind == 0 ? (double)flat : oracle_seeded_value(seed, flat);
    }
}

/* One exact dot product sum_p a[p*sa]*b[p*sb] in double-double (error-free
 * TwoProduct via fma + TwoSum), rounded to double: the sampled-element oracle for
 * sizes where a full CPU GEMM is infeasible. */
#include <math.h>
double oracle_dot_dd(const double *a, long sa, const double *b, long sb, long k) {
  double hi = 0.0, lo = 0.0;
  for (long p = 0; p < k; ++p) {
    const double x = a[p * sa], y = b[p * sb];
    const double prod = x * y;
    const double perr = fma(x, y, -prod);
    const double s = hi + prod;
    const double bb = s - hi;
    const double serr = (hi - (s - bb)) + (prod - bb);
    hi = s;
    lo += serr + perr;
  }
  return hi + lo;
}
