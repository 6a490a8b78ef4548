/*
 * host_band_exec.h — the band-pipelined host-sourced run of the single-GPU SUMMA entry point, written against an
 * abstract stream backend so that the SAME lines run on the GPU (CUDA streams, events, copies, the tensor-core GEMM;
 * phpc_summa.cu) and on the CPU under test (tests/csrc/band_exec_test.cpp: deferred per-stream queues executed in
 * random order, host buffers standing in for HBM, a reference GEMM).  No CUDA types in here on purpose.
 *
 * What it replaces: the reference uploads A, B and C before every kernel and downloads C after it
 * (src/phpc_gemm.cu:111-113,121).  Here the rank's C block is cut into row bands; band b is multiplied over every K chunk
 * in ascending K (the reference's summation order per element, src/phpc_gemm.cu:33-52) into a zeroed block, the caller's C
 * rows (uploaded in the meantime) are added, and the band is downloaded while band b+1 computes.  phpc_host_plan() emits the operation list (pure arithmetic), phpc_band_execute() walks it:
 * one backend stream per plan stream, one event per operation that another stream depends on.
 */
#pragma once
#include <stddef.h>

#include <vector>

#include "../../include/phpc_summa.h"

namespace phpc {

inline long long band_pad_ld(long long cols) { return (cols + 15) / 16 * 16; } /* = phpc_pad_ld of phpc_internal.h */

/* Row bands: the first band is half the block, every further band half of the previous one, the last two equal
 * (1/2, 1/4, 1/8, 1/8 for four bands), each rounded up to a multiple of `align`.  Why not equal bands: the first band has to
 * bring in every B chunk on top of its own A rows, so it must be tall enough for its GEMMs to cover those uploads (and, late
 * in the band, the upload of its own C rows); the last band's download is the exposed tail, so it should be short. */
inline int band_rows(int m, int bands, int align, int band, int *row0_out) {
  if (bands < 1) bands = 1;
  if (align < 1) align = 1;
  int row0 = 0;
  for (int b = 0;; ++b) {
    const int left = m - row0;
    if (left <= 0) return 0; /* fewer bands than asked for: the block is exhausted */
    int rows = left;
    if (b < bands - 1) {
      rows = (left + 1) / 2;
      rows = (rows + align - 1) / align * align;
      if (rows > left) rows = left;
    }
    if (b == band) {
      if (row0_out) *row0_out = row0;
      return rows;
    }
    row0 += rows;
  }
}

/* Issue order: per band  upload its A windows (band 0 also brings every B chunk, interleaved so the first GEMM can start
 * after one chunk), then its C rows into the side buffer; on the compute stream zero the band, the GEMMs of the band over all
 * K chunks in ascending K, add the caller's C rows; then the download.  The caller's C rows are NOT needed before the first
 * GEMM (they are as large as everything the first steps need): C_new = C_old + (P_0 + P_1 + ...). */
inline int host_plan(int m, int nsteps, int bands, int align, phpc_host_op *ops, int max_ops) {
  if (m <= 0 || nsteps <= 0) return 0;
  if (bands < 1) bands = 1;
  if (align < 1) align = 1;
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
  int prev_download = -1; /* the side buffer and the C block are per band: nothing to wait for across bands */
  (void)prev_download;
  for (int band = 0;; ++band) {
    int row0 = 0;
    const int rows = band_rows(m, bands, align, band, &row0);
    if (rows <= 0) break;
    for (int q = 0; q < nsteps; ++q) {
      up_a[q] = emit(PHPC_HOP_UPLOAD_A, 0, band, q, row0, rows, -1, -1, -1);
      if (band == 0) up_b[q] = emit(PHPC_HOP_UPLOAD_B, 0, -1, q, 0, 0, -1, -1, -1);
    }
    const int up_c = emit(PHPC_HOP_UPLOAD_C, 0, band, -1, row0, rows, -1, -1, -1);
    emit(PHPC_HOP_ZERO_C, 1, band, -1, row0, rows, -1, -1, -1);
    for (int q = 0; q < nsteps; ++q) emit(PHPC_HOP_GEMM, 1, band, q, row0, rows, up_a[q], band == 0 ? up_b[q] : -1, -1);
    const int add = emit(PHPC_HOP_ADD_C, 1, band, -1, row0, rows, up_c, -1, -1);
    emit(PHPC_HOP_DOWNLOAD_C, 2, band, -1, row0, rows, add, -1, -1);
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
  double *dC0 = nullptr; /* side buffer of the size of the C block: the caller's C rows, added at the end of each band */
};

/* Streams are 0 = copy-in, 1 = compute, 2 = copy-out.  All calls only ENQUEUE (the backend may run them later, in stream
 * order); an event handle returned by record() stands for "everything enqueued on that stream so far has run". */
struct BandBackend {
  void *self;
  void (*copy2d)(void *self, int stream, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes, size_t rows,
                 int host_to_device);
  void *(*record)(void *self, int stream);
  void (*wait)(void *self, int stream, void *event);
  /* c[rows x n, ldc] += a[rows x width, lda] * b[width x n, ldb]; returns the number of kernels launched.  `step` = K chunk:
   * every band multiplies by the same B chunk, so a backend may keep what it derives from B (the tcgen05 path: its digits) */
  int (*gemm)(void *self, int stream, const double *a, long long lda, const double *b, long long ldb, double *c, long long ldc, int rows,
              int width, int n, int step);
  /* count doubles at dst = 0  /  dst[i] += src[i] for i < count (device memory, contiguous) */
  void (*zero)(void *self, int stream, double *dst, size_t count);
  void (*add)(void *self, int stream, double *dst, const double *src, size_t count);
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
      case PHPC_HOP_UPLOAD_C: /* into the side buffer; its padding columns are zeroed so that ADD_C can run over whole rows */
        if (g.ldn != g.n) be.zero(be.self, o.stream, g.dC0 + (size_t)o.row0 * g.ldn, (size_t)o.rows * g.ldn);
        be.copy2d(be.self, o.stream, g.dC0 + (size_t)o.row0 * g.ldn, (size_t)g.ldn * sizeof(double), hC + host_row * N + (size_t)g.pj * g.n,
                  N * sizeof(double), (size_t)g.n * sizeof(double), (size_t)o.rows, 1);
        break;
      case PHPC_HOP_ZERO_C:
        be.zero(be.self, o.stream, g.dC + (size_t)o.row0 * g.ldn, (size_t)o.rows * g.ldn);
        break;
      case PHPC_HOP_ADD_C:
        be.add(be.self, o.stream, g.dC + (size_t)o.row0 * g.ldn, g.dC0 + (size_t)o.row0 * g.ldn, (size_t)o.rows * g.ldn);
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
                            g.dC + (size_t)o.row0 * g.ldn, g.ldn, o.rows, q.width, g.n, o.step);
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
