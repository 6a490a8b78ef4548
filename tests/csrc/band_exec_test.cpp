/* band_exec_test.cpp — test-only: runs phpc::band_execute (hpc_multigpu_matrixmult_b200/csrc/host_band_exec.h), the very
 * walk over the operation list that the CUDA build executes, against a deferred stream backend on the CPU: every call
 * is only queued on its stream; afterwards the queues are drained in a random interleaving that respects stream order
 * and event waits, exactly the freedom a GPU has.  Host buffers stand in for HBM and start as NaN, so anything used
 * before it arrived, or any wrong window, poisons the result.  Built and driven by tests/test_band_executor.py. */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

#include "../../hpc_multigpu_matrixmult_b200/csrc/host_band_exec.h"

namespace {
struct Event {
  int stream;
  size_t position; /* complete once `position` operations of `stream` have run */
};
struct Op {
  std::function<void()> run;
  const Event *wait; /* non-null: a cudaStreamWaitEvent; runnable once the event is complete */
};
struct Sim {
  std::vector<Op> queue[3];
  size_t head[3] = {0, 0, 0};
  std::vector<Event *> events;
  int gemms = 0;
};
void sim_copy2d(void *self, int stream, void *dst, size_t dp, const void *src, size_t sp, size_t w, size_t rows, int) {
  Sim *s = (Sim *)self;
  s->queue[stream].push_back({[=]() {
                                for (size_t r = 0; r < rows; ++r) memcpy((char *)dst + r * dp, (const char *)src + r * sp, w);
                              },
                              nullptr});
}
void *sim_record(void *self, int stream) {
  Sim *s = (Sim *)self;
  Event *e = new Event{stream, s->queue[stream].size()};
  s->events.push_back(e);
  return e;
}
void sim_wait(void *self, int stream, void *event) {
  Sim *s = (Sim *)self;
  s->queue[stream].push_back({[]() {}, (const Event *)event});
}
int sim_gemm(void *self, int stream, const double *a, long long lda, const double *b, long long ldb, double *c, long long ldc, int rows,
             int width, int n, int /* step */) {
  Sim *s = (Sim *)self;
  s->queue[stream].push_back({[=]() {
                                for (int i = 0; i < rows; ++i)
                                  for (int j = 0; j < n; ++j) {
                                    double acc = 0.0; /* sum over the chunk first, then C += (reference src/phpc_gemm.cu:33-55) */
                                    for (int q = 0; q < width; ++q) acc += a[(long long)i * lda + q] * b[(long long)q * ldb + j];
                                    c[(long long)i * ldc + j] += acc;
                                  }
                              },
                              nullptr});
  ++s->gemms;
  return 1;
}
void sim_zero(void *self, int stream, double *dst, size_t count) {
  Sim *s = (Sim *)self;
  s->queue[stream].push_back({[=]() {
                                for (size_t i = 0; i < count; ++i) dst[i] = 0.0;
                              },
                              nullptr});
}
void sim_add(void *self, int stream, double *dst, const double *src, size_t count) {
  Sim *s = (Sim *)self;
  s->queue[stream].push_back({[=]() {
                                for (size_t i = 0; i < count; ++i) dst[i] += src[i];
                              },
                              nullptr});
}
}  // namespace

/* C += A * B for FULL N x N host matrices through the band pipeline; order 0 = every stream runs eagerly in issue order,
 * 1 = random interleaving (seeded).  Returns the number of GEMMs, -1 on a deadlock, -2 on an unknown operation. */
extern "C" int band_exec_sim(int N, int kc, int bands, int align, const double *hA, const double *hB, double *hC, unsigned seed, int order) {
  const int m = N, n = N;
  const long long ldn = phpc::band_pad_ld(n);
  std::vector<phpc_summa_step> steps;
  long long a_off = 0;
  for (int k0 = 0; k0 < N; k0 += kc) { /* the 1 x 1 schedule of phpc_summa_schedule: every chunk is owned */
    phpc_summa_step st;
    memset(&st, 0, sizeof st);
    st.k0 = k0;
    st.width = (N - k0 < kc) ? N - k0 : kc;
    st.own_a = st.own_b = 1;
    st.a_off = a_off;
    st.b_off = (long long)k0 * ldn;
    a_off += (long long)m * phpc::band_pad_ld(st.width);
    steps.push_back(st);
  }
  std::vector<double> dA((size_t)a_off, NAN), dB((size_t)N * ldn, NAN), dC((size_t)m * ldn, NAN), dC0((size_t)m * ldn, NAN);
  const int nops = phpc::host_plan(m, (int)steps.size(), bands, align, nullptr, 0);
  std::vector<phpc_host_op> ops(nops);
  phpc::host_plan(m, (int)steps.size(), bands, align, ops.data(), nops);
  Sim sim;
  phpc::BandBackend be = {&sim, sim_copy2d, sim_record, sim_wait, sim_gemm, sim_zero, sim_add};
  phpc::BandGeom g;
  g.N = N;
  g.m = m;
  g.n = n;
  g.pi = g.pj = 0;
  g.ldn = ldn;
  g.steps = steps.data();
  g.nsteps = (int)steps.size();
  g.dA = dA.data();
  g.dB = dB.data();
  g.dC = dC.data();
  g.dC0 = dC0.data();
  if (phpc::band_execute(g, ops.data(), nops, hA, hB, hC, be) < 0) return -2;
  std::mt19937 rng(seed);
  for (;;) {
    int runnable[3], nr = 0;
    bool pending = false;
    for (int s = 0; s < 3; ++s) {
      if (sim.head[s] >= sim.queue[s].size()) continue;
      pending = true;
      const Op &op = sim.queue[s][sim.head[s]];
      if (op.wait && sim.head[op.wait->stream] < op.wait->position) continue;
      runnable[nr++] = s;
    }
    if (!pending) break;
    if (nr == 0) return -1;
    const int s = order == 0 ? runnable[0] : runnable[rng() % nr];
    sim.queue[s][sim.head[s]].run();
    ++sim.head[s];
  }
  for (Event *e : sim.events) delete e;
  return sim.gemms;
}
