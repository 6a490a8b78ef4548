/*
 * host_band_exec.h — the band-pipelined host-sourced run of the single-GPU SUMMA entry point, written against an
 * abstract stream backend so that the SAME lines run on the GPU (CUDA streams, events, copies, the tensor-core GEMM;
 * phpc_summa.cu) and on the CPU under test (tests/csrc/band_exec_test.cpp: deferred per-stream queues executed in
 * random order, host buffers standing in for HBM, a reference GEMM).  No CUDA types in here on purpose.
 *
 * What it replaces: the reference uploads A, B and C before every kernel and downloads C after it
 * (src/phpc_gemm.cu:111-113,121).  Here the rank's C block is cut into row bands; band b is uploaded, multiplied over
 * every K chunk in ascending K (the reference's summation order per element, src/phpc_gemm.cu:33-52) and downloaded
 * while band b+1 computes.  phpc_host_plan() emits the operation list (pure arithmetic), phpc_band_execute() walks it:
 * one backend stream per plan stream, one event per operation that another stream depends on.
 */
#pragma once
#include <stddef.h>

#include <vector>

#include "../../include/phpc_summa.h"

namespace phpc {

inline long long band_pad_ld(long long cols) { return (cols + 15) / 16 * 16; } /* = phpc_pad_ld of phpc_internal.h */

/* Issue order: per band  upload C band, upload its A windows (band 0 also brings every B chunk, interleaved so the first
 * GEMM can start after one chunk), then the GEMMs of the band over all K chunks, then the download. */
inline int host_plan(int m, int nsteps, int bands, int align, phpc_host_op *ops, int max_ops) {
  if (m <= 0 || nsteps <= 0) return 0;
  if (bands < 1) bands = 1;
  if (align < 1) align = 1;
  int rows_per_band = (m + bands - 1) / bands;
  rows_per_band = (rows_per_band + align - 1) / align * align;
  int count = 0;
  auto emit = [&](int kind, int stream, int band, int step, int row0, int rows, int d0, int d1, int d2) {
    if (ops && count < max_ops) {
      phpc_host_op *o = &ops[count];
      o->kind = kind;
      o->stream = stream;
      o->band = band;
      o->step = step;
      o->row0 = row0;
      o->rows = rows;
      o->ndeps = 0;
      const int d[3] = {d0, d1, d2};
      for (int i = 0; i < 3; ++i)
        if (d[i] >= 0) o->deps[o->ndeps++] = d[i];
      for (int i = o->ndeps; i < 3; ++i) o->deps[i] = -1;
    }
    return count++;
  };
  std::vector<int> up_b(nsteps, -1), up_a(nsteps, -1);
  for (int band = 0, row0 = 0; row0 < m; ++band, row0 += rows_per_band) {
    const int rows = (m - row0 < rows_per_band) ? m - row0 : rows_per_band;
    const int up_c = emit(PHPC_HOP_UPLOAD_C, 0, band, -1, row0, rows, -1, -1, -1);
    for (int q = 0; q < nsteps; ++q) {
      up_a[q] = emit(PHPC_HOP_UPLOAD_A, 0, band, q, row0, rows, -1, -1, -1);
      if (band == 0) up_b[q] = emit(PHPC_HOP_UPLOAD_B, 0, -1, q, 0, 0, -1, -1, -1);
    }
    int last = -1;
    for (int q = 0; q < nsteps; ++q) last = emit(PHPC_HOP_GEMM, 1, band, q, row0, rows, q == 0 ? up_c : -1, up_a[q], band == 0 ? up_b[q] : -1);
    emit(PHPC_HOP_DOWNLOAD_C, 2, band, -1, row0, rows, last, -1, -1);
  }
  return count;
}

/* The rank's blocks as the executor needs them (a 1 x 1 grid has pi = pj = 0, m = n = N; the fields are kept general). */
struct BandGeom {
  int N;           /* leading dimension of the FULL host matrices B and C (global columns) */
  int lda_host = 0; /* leading dimension of the FULL host matrix A (global K); 0 = N (square problem) */
  int m, n;        /* the rank's C block */
  int pi, pj;      /* grid coordinates: the block starts at host row pi*m, column pj*n */
  long long ldn;   /* leading dimension of the B store and of the C block in HBM */
  const phpc_summa_step *steps; /* K chunks: k0, width, a_off, b_off */
  int nsteps;
  double *dA, *dB, *dC; /* A store (chunk q: [m][pad(width)] at a_off), B store (chunk q at b_off, ld ldn), C block */
};

/* Streams are 0 = copy-in, 1 = compute, 2 = copy-out.  All calls only ENQUEUE (the backend may run them later, in stream
 * order); an event handle returned by record() stands for "everything enqueued on that stream so far has run". */
struct BandBackend {
  void *self;
  void (*copy2d)(void *self, int stream, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes, size_t rows,
                 int host_to_device);
  void *(*record)(void *self, int stream);
  void (*wait)(void *self, int stream, void *event);
  /* c[rows x n, ldc] += a[rows x width, lda] * b[width x n, ldb]; returns the number of kernels launched */
  int (*gemm)(void *self, int stream, const double *a, long long lda, const double *b, long long ldb, double *c, long long ldc, int rows,
              int width, int n);
};

/* Walks the operation list; returns the number of GEMM kernels launched.  Does not synchronise. */
inline int band_execute(const BandGeom &g, const phpc_host_op *ops, int nops, const double *hA, const double *hB, double *hC,
                        const BandBackend &be) {
  const size_t N = (size_t)g.N, KA = (size_t)(g.lda_host > 0 ? g.lda_host : g.N);
  std::vector<void *> done(nops, nullptr);
  std::vector<char> needed(nops, 0);
  for (int i = 0; i < nops; ++i)
    for (int d = 0; d < ops[i].ndeps; ++d) needed[ops[i].deps[d]] = 1;
  int launches = 0;
  for (int i = 0; i < nops; ++i) {
    const phpc_host_op &o = ops[i];
    for (int d = 0; d < o.ndeps; ++d) be.wait(be.self, o.stream, done[o.deps[d]]);
    const size_t host_row = (size_t)g.pi * g.m + o.row0; /* first row of the band in the full host matrices */
    switch (o.kind) {
      case PHPC_HOP_UPLOAD_C:
        be.copy2d(be.self, o.stream, g.dC + (size_t)o.row0 * g.ldn, (size_t)g.ldn * sizeof(double), hC + host_row * N + (size_t)g.pj * g.n,
                  N * sizeof(double), (size_t)g.n * sizeof(double), (size_t)o.rows, 1);
        break;
      case PHPC_HOP_UPLOAD_A: {
        const phpc_summa_step &q = g.steps[o.step];
        const size_t ld = (size_t)band_pad_ld(q.width);
        be.copy2d(be.self, o.stream, g.dA + q.a_off + (size_t)o.row0 * ld, ld * sizeof(double), hA + host_row * KA + (size_t)q.k0,
                  KA * sizeof(double), (size_t)q.width * sizeof(double), (size_t)o.rows, 1);
        break;
      }
      case PHPC_HOP_UPLOAD_B: {
        const phpc_summa_step &q = g.steps[o.step];
        be.copy2d(be.self, o.stream, g.dB + q.b_off, (size_t)g.ldn * sizeof(double), hB + (size_t)q.k0 * N + (size_t)g.pj * g.n,
                  N * sizeof(double), (size_t)g.n * sizeof(double), (size_t)q.width, 1);
        break;
      }
      case PHPC_HOP_GEMM: {
        const phpc_summa_step &q = g.steps[o.step];
        const long long ld = band_pad_ld(q.width);
        launches += be.gemm(be.self, o.stream, g.dA + q.a_off + (size_t)o.row0 * ld, ld, g.dB + q.b_off, g.ldn,
                            g.dC + (size_t)o.row0 * g.ldn, g.ldn, o.rows, q.width, g.n);
        break;
      }
      case PHPC_HOP_DOWNLOAD_C:
        be.copy2d(be.self, o.stream, hC + host_row * N + (size_t)g.pj * g.n, N * sizeof(double), g.dC + (size_t)o.row0 * g.ldn,
                  (size_t)g.ldn * sizeof(double), (size_t)g.n * sizeof(double), (size_t)o.rows, 0);
        break;
      default:
        return -1;
    }
    if (needed[i]) done[i] = be.record(be.self, o.stream);
  }
  return launches;
}

}  // namespace phpc
