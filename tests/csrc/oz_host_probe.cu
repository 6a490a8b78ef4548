/* oz_host_probe.cu — test-only: runs the __host__ __device__ bodies of the split kernels of the tcgen05 path
 * (hpc_multigpu_matrixmult_b200/csrc/ozaki_split.cuh) on the CPU, so tests/test_ozaki_split_host.py can check
 * the very lines the GPU executes (digits, store layout) against oracle/ozaki_model.py without a GPU.
 * Built by the test with nvcc as host code; never linked into the product library. */
#include "../../hpc_multigpu_matrixmult_b200/csrc/ozaki_split.cuh"

using namespace phpc::oz;

extern "C" {
int oz_probe_zero_exp(void) { return ZERO_EXP; }
int oz_probe_digits_per_operand(void) { return S; }
int oz_probe_tile_offset(int r, int kb) { return tile_offset(r, kb); }
long long oz_probe_store_offset(int row, int kbyte, int t, int ksteps) { return (long long)store_offset(row, kbyte, t, ksteps); }
/* out[i][t]: the 7 balanced base-256 digits of x[i] under exponent e, most significant first */
void oz_probe_digits(const double *x, int n, int e, int8_t *out) {
  for (int i = 0; i < n; ++i) balanced_digits(x[i], e, out + (size_t)i * S);
}
void oz_probe_split_a(const double *A, long long lda, int m, int m_pad, int k, int kp, const int *eA, int8_t *TA) {
  const long long items = (long long)m_pad * (kp / 16);
  for (long long idx = 0; idx < items + 3; ++idx) /* + 3: the out-of-range guard of the body */
    split_a_body(idx, A, lda, m, m_pad, k, kp, eA, TA);
}
void oz_probe_split_b(const double *B, long long ldb, int k, int n, int n_pad, int kp, const int *eB, int8_t *TB) {
  for (int ks = 0; ks < kp / 32; ++ks)
    for (int col = 0; col < n_pad + 3; ++col) split_b_body(col, ks, B, ldb, k, n, n_pad, kp, eB, TB);
}
}
