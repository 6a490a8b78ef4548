/* oz_host_probe.cu — test-only: runs the __host__ __device__ bodies of the experimental Ozaki split kernels
 * (hpc_multigpu_matrixmult_b200/csrc/ozaki_split.cuh) on the CPU, so tests/test_ozaki_split_host.py can check
 * the very lines the GPU executes (digits, store layout) against oracle/ozaki_model.py without a GPU.
 * Built by the test with nvcc as host code; never linked into the product library. */
#include "../../hpc_multigpu_matrixmult_b200/csrc/ozaki_split.cuh"

using namespace phpc::oz;

extern "C" {
int oz_probe_zero_exp(void) { return ZERO_EXP; }
int oz_probe_tile_offset(int r, int kb) { return tile_offset(r, kb); }
long long oz_probe_store_offset(int row, int kbyte, int t, int S, int ksteps, int halves) {
  return (long long)store_offset(row, kbyte, t, S, ksteps, halves);
}
/* out[i][t]: digits of x[i] under exponent e; bal = 0: 8 truncated 7-bit digits, bal = 1: 7 balanced base-256 digits */
void oz_probe_digits(int bal, const double *x, int n, int e, int8_t *out) {
  for (int i = 0; i < n; ++i) {
    if (bal)
      digits_of<true, 7>(x[i], e, out + (size_t)i * 7);
    else
      digits_of<false, 8>(x[i], e, out + (size_t)i * 8);
  }
}
void oz_probe_split_a(int bal, const double *A, long long lda, int m, int m_pad, int k, int kp, const int *eA, int8_t *TA) {
  const long long items = (long long)m_pad * (kp / 16);
  for (long long idx = 0; idx < items + 3; ++idx) { /* + 3: the out-of-range guard of the body */
    if (bal)
      split_a_body<true, 7>(idx, A, lda, m, m_pad, k, kp, eA, TA);
    else
      split_a_body<false, 8>(idx, A, lda, m, m_pad, k, kp, eA, TA);
  }
}
void oz_probe_split_b(int bal, const double *B, long long ldb, int k, int n, int n_pad, int kp, const int *eB, int8_t *TB, int halves) {
  for (int ks = 0; ks < kp / 32; ++ks)
    for (int col = 0; col < n_pad + 3; ++col) {
      if (bal)
        split_b_body<true, 7>(col, ks, B, ldb, k, n, n_pad, kp, eB, TB, halves);
      else
        split_b_body<false, 8>(col, ks, B, ldb, k, n, n_pad, kp, eB, TB, halves);
    }
}
}
