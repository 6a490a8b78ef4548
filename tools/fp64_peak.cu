/*
 * fp64_peak.cu — measures the FP64 roofline denominators BASELINE.md says the
 * builder must measure on this pool's B200 (MEASURED_PEAKS.json only has bf16):
 *   dmma   register-resident mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) issue rate
 *   dfma   register-resident DFMA issue rate (CUDA-core FP64)
 *   dgemm  cublasDgemm N^3, burst (best of R) and sustained (back to back for S seconds)
 * Prints one JSON object; profiles/fp64_peak_*.json keeps the runs quoted in DESIGN.md.
 */
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e = (x);                                                                  \
    if (e != cudaSuccess) {                                                               \
      fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                            \
    }                                                                                     \
  } while (0)

template <int ACCS>
__global__ void __launch_bounds__(1024) dmma_rate(double *out, int iters, double a0, double b0) {
  double c[ACCS][2];
#pragma unroll
  for (int i = 0; i < ACCS; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACCS; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACCS; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int ACCS>
__global__ void __launch_bounds__(1024) dfma_rate(double *out, int iters, double a0, double b0) {
  double c[ACCS];
#pragma unroll
  for (int i = 0; i < ACCS; ++i) c[i] = i;
  const double a = a0, b = b0 + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACCS; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACCS; ++i) s += c[i];
  if (s == 12345.678) out[0] = s;
}

static float time_launch(void (*launch)(void *), void *arg, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch(arg);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch(arg);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

struct RateArgs {
  double *out;
  int sms, warps, iters;
};
static void launch_dmma(void *p) {
  RateArgs *a = (RateArgs *)p;
  dmma_rate<16><<<a->sms, a->warps * 32>>>(a->out, a->iters, 1.0, 1.0);
}
static void launch_dfma(void *p) {
  RateArgs *a = (RateArgs *)p;
  dfma_rate<16><<<a->sms, a->warps * 32>>>(a->out, a->iters, 1.0000001, 1e-9);
}

int main(int argc, char **argv) {
  int n_big = argc > 1 ? atoi(argv[1]) : 16384;
  double sustain_s = argc > 2 ? atof(argv[2]) : 4.0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double *out;
  CK(cudaMalloc(&out, 64));

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_max\": %d", prop.name, sms, prop.clockRate / 1000);

  /* --- DMMA / DFMA issue rates for 4, 8, 16 warps per SM --- */
  const int warp_cfg[3] = {4, 8, 16};
  printf(", \"dmma_tflops\": {");
  for (int w = 0; w < 3; ++w) {
    RateArgs a = {out, sms, warp_cfg[w], 4096};
    const float ms = time_launch(launch_dmma, &a, 5);
    const double flops = 2.0 * 256 * 16 * (double)a.iters * a.warps * sms; /* 8x8x4 = 256 FMA per warp-instruction */
    printf("%s\"%dw\": %.2f", w ? ", " : "", warp_cfg[w], flops / ms / 1e9);
  }
  printf("}, \"dfma_tflops\": {");
  for (int w = 0; w < 3; ++w) {
    RateArgs a = {out, sms, warp_cfg[w], 4096};
    const float ms = time_launch(launch_dfma, &a, 5);
    const double flops = 2.0 * 32 * 16 * (double)a.iters * a.warps * sms;
    printf("%s\"%dw\": %.2f", w ? ", " : "", warp_cfg[w], flops / ms / 1e9);
  }
  printf("}");
  fflush(stdout);

  /* --- cuBLAS Dgemm --- */
  cublasHandle_t h;
  cublasCreate(&h);
  const int sizes[2] = {8192, n_big};
  printf(", \"dgemm\": [");
  for (int si = 0; si < 2; ++si) {
    const size_t n = sizes[si];
    double *A, *B, *C;
    CK(cudaMalloc(&A, n * n * 8));
    CK(cudaMalloc(&B, n * n * 8));
    CK(cudaMalloc(&C, n * n * 8));
    CK(cudaMemset(A, 0, n * n * 8));
    CK(cudaMemset(B, 0, n * n * 8));
    CK(cudaMemset(C, 0, n * n * 8));
    const double one = 1.0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, B, n, A, n, &one, C, n);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
      CK(cudaEventRecord(e0));
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, B, n, A, n, &one, C, n);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    const double flops = 2.0 * n * n * n;
    /* sustained: back to back for sustain_s seconds */
    const int reps = (int)(sustain_s * 1000.0 / best) + 1;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, B, n, A, n, &one, C, n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float total;
    CK(cudaEventElapsedTime(&total, e0, e1));
    printf("%s{\"n\": %zu, \"burst_tflops\": %.2f, \"sustained_tflops\": %.2f, \"sustained_reps\": %d}", si ? ", " : "", n,
           flops / best / 1e9, flops * reps / total / 1e9, reps);
    fflush(stdout);
    cudaFree(A);
    cudaFree(B);
    cudaFree(C);
  }
  printf("]}\n");
  return 0;
}
