/*
 * phpc_summa.cu — the SUMMA outer loop, device resident, NCCL over NVLink.
 *
 * Follows the distribution and schedule of reference src/phpc_summa.c:24-122
 * (lcm(r,c) K panels; A panel k broadcast along the process row by column k%c,
 * B panel k along the process column by row k%r; C block += panel product; gather
 * to rank 0) but is a different program: blocks live in HBM, panels are cut into
 * K chunks that are stored contiguously on their owner (so a chunk IS an NCCL
 * send buffer, the MPI_Type_vector packing of reference :53-59 disappears), the
 * two MPI_Bcast of :75-88 become one ncclGroup of two ncclBroadcast on a
 * high-priority communication stream, and a ring of receive buffers lets the
 * broadcasts of the next chunks run under the DMMA GEMM of the current one.
 */
#include <mpi.h>
#include <nccl.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "../../include/phpc_b200.h"
#include "../../include/phpc_gemm.cuh"
#include "../../include/phpc_summa.h"
#include "host_band_exec.h"
#include "phpc_internal.h"

#define NCCL_CHECK(call)                                                            \
  do {                                                                              \
    ncclResult_t r__ = (call);                                                      \
    if (r__ != ncclSuccess) phpc_die(#call, ncclGetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

#define PHPC_MAX_HOST_BANDS 4

static int gcd_int(int a, int b) {
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

#include <sys/time.h>
static double now_s() {
  struct timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec + tv.tv_usec * 1e-6;
}
#define PHPC_TRACE(rank, what, t0)                                                         \
  do {                                                                                     \
    if (getenv("PHPC_DEBUG")) fprintf(stderr, "[phpc %d] %-28s %.3f s\n", rank, what, now_s() - (t0)); \
  } while (0)

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

/* ------------------------------------------------------------------------- */
/* schedule (pure host arithmetic)                                            */
/* ------------------------------------------------------------------------- */
/* kc_first > 0: the very first chunk (of panel 0) is only kc_first wide.  Nothing can overlap the transfer of the first chunk
 * (every rank waits for it, and up to c - 1 peers pull it from one owner at once), so it is kept small; its short GEMM then
 * covers the transfer of the next chunk.  With grow_pct > 100 the following chunks of panel 0 grow by that factor
 * (rounded up to 128) until they reach kc: the single-GPU host-sourced run, where the upload of chunk q+1 (PCIe) is only ~25 %
 * faster than the GEMM of chunk q, so a jump from a short chunk to a full one would expose most of the full chunk's upload. */
static int summa_schedule(int M, int K, int N, int r, int c, int pi, int pj, int kc, int kc_first, int grow_pct, phpc_summa_step *steps,
                          int max_steps, int *m_out, int *n_out) {
  if (M <= 0 || K <= 0 || N <= 0 || r <= 0 || c <= 0 || M % r || N % c) return -1;
  if (M <= 0 || K <= 0 || N <= 0 || r <= 0 || c <= 0 || M % r || N % c) return -1;
  const int lcm = r / gcd_int(r, c) * c;
  if (K % lcm) return -1;
  const int m = M / r, n = N / c, pk = K / lcm; /* reference :36-39 (square there: M = K = N) */
  if (kc <= 0 || kc > pk) kc = pk;
  if (m_out) *m_out = m;
  if (n_out) *n_out = n;
  const long long ldn = phpc_pad_ld(n);
  int count = 0;
  long long a_off = 0;
  for (int k = 0; k < lcm; ++k) {
    const int a_root = k % c, b_root = k % r; /* reference :64-65 */
    const int own_a = (a_root == pj), own_b = (b_root == pi);
    const int b_local_panel = k / r; /* how many panels this row owned before k (own_b) */
    int ramp = (k == 0) ? kc_first : 0;
    for (int k_in = 0, width = 0; k_in < pk; k_in += width) {
      width = (pk - k_in < kc) ? pk - k_in : kc;
      if (ramp > 0 && ramp < width) {
        width = ramp;
        ramp = grow_pct > 100 ? (int)(((long long)ramp * grow_pct / 100 + 127) / 128 * 128) : 0;
      } else {
        ramp = 0;
      }
      if (steps && count < max_steps) {
        phpc_summa_step *s = &steps[count];
        s->panel = k;
        s->a_root = a_root;
        s->b_root = b_root;
        s->k0 = (long long)k * pk + k_in;
        s->width = width;
        s->own_a = own_a;
        s->own_b = own_b;
        s->a_off = own_a ? a_off : -1;
        s->b_off = own_b ? ((long long)b_local_panel * pk + k_in) * ldn : -1;
      }
      if (own_a) a_off += (long long)m * phpc_pad_ld(width);
      ++count;
    }
  }
  return count;
}

extern "C" int phpc_summa_schedule_mkn(int M, int K, int N, int r, int c, int pi, int pj, int kc, phpc_summa_step *steps, int max_steps,
                                       int *m_out, int *n_out) {
  return summa_schedule(M, K, N, r, c, pi, pj, kc, 0, 0, steps, max_steps, m_out, n_out);
}

extern "C" int phpc_summa_schedule_first(int M, int K, int N, int r, int c, int pi, int pj, int kc, int kc_first, phpc_summa_step *steps,
                                         int max_steps, int *m_out, int *n_out) {
  return summa_schedule(M, K, N, r, c, pi, pj, kc, kc_first, 0, steps, max_steps, m_out, n_out);
}

extern "C" int phpc_summa_schedule(int N, int r, int c, int pi, int pj, int kc, phpc_summa_step *steps, int max_steps, int *m_out,
                                   int *n_out) {
  if (N <= 0 || r <= 0 || c <= 0 || N % r || N % c) return -1; /* N % lcm == 0 follows */
  return phpc_summa_schedule_mkn(N, N, N, r, c, pi, pj, kc, steps, max_steps, m_out, n_out);
}

/* Operation list of the band-pipelined host-sourced run on one GPU (phpc_host_op in phpc_summa.h): pure arithmetic,
 * shared with the CPU test of the executor through host_band_exec.h. */
extern "C" int phpc_host_plan(int m, int nsteps, int bands, int align, phpc_host_op *ops, int max_ops) {
  return phpc::host_plan(m, nsteps, bands, align, ops, max_ops);
}

/* ------------------------------------------------------------------------- */
/* NCCL communicators, cached per grid shape                                  */
/* ------------------------------------------------------------------------- */
struct NcclGrid {
  bool ready = false;
  int size = 0, r = 0, c = 0, rank = 0;
  unsigned generation = 0; /* bumped whenever the communicators are rebuilt (registrations die with them) */
  ncclComm_t world = nullptr, row = nullptr, col = nullptr; /* row: same pi, rank = pj; col: same pj, rank = pi */
};
static NcclGrid g_nccl;

static void nccl_grid_release() {
  if (!g_nccl.ready) return;
  if (g_nccl.row) ncclCommDestroy(g_nccl.row);
  if (g_nccl.col) ncclCommDestroy(g_nccl.col);
  if (g_nccl.world) ncclCommDestroy(g_nccl.world);
  const unsigned gen = g_nccl.generation + 1;
  g_nccl = NcclGrid();
  g_nccl.generation = gen;
}

static void nccl_grid_get(MPI_Comm grid_comm, int size, int rank, int r, int c, int pi, int pj) {
  if (g_nccl.ready && g_nccl.size == size && g_nccl.r == r && g_nccl.c == c && g_nccl.rank == rank) return;
  nccl_grid_release();
  g_nccl.size = size;
  g_nccl.r = r;
  g_nccl.c = c;
  g_nccl.rank = rank;
  if (size > 1) {
    /* one node, NVLink only: do not let NCCL probe InfiniBand / every NIC or set up NVLS
     * multicast for communicators that only broadcast (user settings win) */
    setenv("NCCL_IB_DISABLE", "1", 0);
    setenv("NCCL_SOCKET_IFNAME", "lo", 0);
    setenv("NCCL_NVLS_ENABLE", "0", 0);
    ncclUniqueId id;
    memset(&id, 0, sizeof id);
    if (rank == 0) NCCL_CHECK(ncclGetUniqueId(&id));
    MPI_Bcast(&id, (int)sizeof id, MPI_BYTE, 0, grid_comm);
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    /* the GEMM is persistent and wants every SM: keep the broadcasts on a handful of CTAs */
    cfg.minCTAs = 1;
    cfg.maxCTAs = env_int("PHPC_NCCL_MAX_CTAS", 4);
    NCCL_CHECK(ncclCommInitRankConfig(&g_nccl.world, size, id, rank, &cfg));
    /* MPI_Cart_sub(remain {0,1}) / {1,0} of reference :27-34 */
    ncclConfig_t cfg_row = cfg, cfg_col = cfg;
    NCCL_CHECK(ncclCommSplit(g_nccl.world, c > 1 ? pi : NCCL_SPLIT_NOCOLOR, pj, &g_nccl.row, &cfg_row));
    NCCL_CHECK(ncclCommSplit(g_nccl.world, r > 1 ? pj : NCCL_SPLIT_NOCOLOR, pi, &g_nccl.col, &cfg_col));
  }
  g_nccl.ready = true;
}

/* ------------------------------------------------------------------------- */
/* the SUMMA object                                                           */
/* ------------------------------------------------------------------------- */
struct phpc_summa {
  MPI_Comm grid_comm;
  int rank, size;
  int kc_first = 0;                       /* width of the very first K chunk (0 = like the others) */
  int grow_pct = 0;                       /* > 100: the chunks after it grow by this factor until they reach kc (one-rank host runs) */
  int N, r, c, pi, pj, lcm, m, n, pk, kc; /* N = global columns of B and C (= leading dimension of host B, C) */
  int gM, gK;                             /* global rows of A and C, global K (= leading dimension of host A); square: all N */
  long long ldn;   /* padded leading dimension of B chunks and of C */
  long long lda_k; /* padded leading dimension of a full-width A chunk */
  std::vector<phpc_summa_step> steps;
  DeviceCtx *ctx;
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  size_t a_elems = 0, b_elems = 0, c_elems = 0;
  int nbuf = 0;
  double *ringA = nullptr, *ringB = nullptr; /* nbuf receive buffers each */
  /* PHPC_NCCL_REGISTER=1: stores and rings registered with the row / column communicator (ncclCommRegister) */
  struct NcclReg {
    ncclComm_t comm;
    void *handle;
  };
  std::vector<NcclReg> nccl_regs;
  unsigned nccl_generation = 0;
  double *gather_stage = nullptr;            /* rank 0: two C-block landing buffers for the gather */
  double *dC0 = nullptr;                     /* multi-rank host-sourced runs: the caller's C block, uploaded under the loop and added at the end */
  /* panel transport: 0 = ncclBroadcast on row/column communicators, 1 = copy-engine pull from the
   * owner's store through CUDA IPC peer mappings (no SMs, no rendezvous) */
  int transport = 0;
  std::vector<double *> peerA, peerB;        /* peerA[pj'] = dA of rank (pi, pj'); peerB[pi'] = dB of rank (pi', pj) */
  std::vector<void *> peerA_base, peerB_base; /* what cudaIpcOpenMemHandle returned (allocation bases) */
  std::vector<void *> peerC;                  /* rank 0 only: dC of every other rank */
  std::vector<long long> root_a_off, root_b_off; /* per step: chunk offset inside its ROOT's store */
  std::vector<cudaEvent_t> ev_bcast2;        /* per ring slot: B pull done (pull transport) */
  size_t ringA_elems = 0, ringB_elems = 0;
  std::vector<cudaEvent_t> ev_bcast, ev_free; /* per ring slot */
  std::vector<cudaEvent_t> ev_g0, ev_g1;      /* per step: GEMM start / stop */
  std::vector<cudaEvent_t> ev_up;             /* per step: owned chunks uploaded (host-sourced runs); interprocess events on a multi-rank pull grid */
  std::vector<cudaEvent_t> peer_up_a, peer_up_b; /* per step: the A / B root's ev_up of that step, opened through CUDA IPC (null when own) */
  /* multi-rank pull grids carry PHPC_MAX_HOST_BANDS x nsteps upload events (index band * nsteps + step): the banded
   * host-sourced run uploads the A rows of one row band at a time */
  std::vector<cudaEvent_t> ev_c0, ev_band;     /* per band: caller's C band uploaded / band final in HBM */
  std::vector<cudaEvent_t> ev_bg0, ev_bg1;     /* per (band, step): GEMM start / stop of the banded run */
  cudaStream_t d2h[2] = {nullptr, nullptr};    /* downloads of finished bands: own C, rank 0's shared C */
  /* banded host-sourced runs multiply every row band by the same B chunks: exponents and digits of B chunk q are computed by
   * the first band and kept for the others (tcgen05 path; 7 bytes per element of the rank's B columns over all of K) */
  std::vector<OzBCache> bcache;
  signed char *bcache_tb = nullptr;
  int *bcache_e = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_user = nullptr, ev_cup = nullptr;
};

static int pick_device(int rank) {
  const int count = phpc_b200_device_count();
  PHPC_REQUIRE(count > 0, "no CUDA device visible (this library has no CPU fallback)");
  const char *d = getenv("PHPC_DEVICE");
  if (d && *d) return atoi(d);
  const char *lr = getenv("LOCAL_RANK");
  if (lr && *lr) return atoi(lr) % count;
  return rank % count;
}

static phpc_summa *summa_create(MPI_Comm grid_comm, int gm, int gk, int n, int kc, int kc_first_single, int grow_pct_single);

extern "C" phpc_summa *phpc_summa_create(MPI_Comm grid_comm, int n, int kc) { return summa_create(grid_comm, n, n, n, kc, 0, 0); }

extern "C" phpc_summa *phpc_summa_create_mkn(MPI_Comm grid_comm, int gm, int gk, int n, int kc) {
  return summa_create(grid_comm, gm, gk, n, kc, 0, 0);
}

/* kc_first_single / grow_pct_single: short first chunk and growth of the following ones on a ONE-rank grid (the host-sourced
 * entry points ask for 1024 / 133 %); multi-rank grids always start with one short chunk (PHPC_KC_FIRST, 2048) */
static phpc_summa *summa_create(MPI_Comm grid_comm, int gm, int gk, int n, int kc, int kc_first_single, int grow_pct_single) {
  phpc_summa *s = new phpc_summa();
  s->grid_comm = grid_comm;
  int dims[2], periods[2], coords[2];
  MPI_Comm_rank(grid_comm, &s->rank);
  MPI_Comm_size(grid_comm, &s->size);
  MPI_Cart_get(grid_comm, 2, dims, periods, coords);
  s->N = n;
  s->gM = gm;
  s->gK = gk;
  s->r = dims[0];
  s->c = dims[1];
  s->pi = coords[0];
  s->pj = coords[1];
  s->lcm = s->r / gcd_int(s->r, s->c) * s->c;
  PHPC_REQUIRE(n > 0 && gm > 0 && gk > 0 && gm % s->r == 0 && n % s->c == 0 && gk % s->lcm == 0,
               "matrix size must be divisible by the process grid dimensions");
  s->m = gm / s->r;
  s->n = n / s->c;
  s->pk = gk / s->lcm;
  /* multi-rank default 8192 = the K chunk of the tcgen05 launcher: one set of exponent / split kernels and two C passes per
   * transferred chunk (4096 doubled both per flop); the ring buffers stay below 2 GiB per slot up to N = 65536 on 2 x 4 */
  if (kc <= 0) kc = env_int("PHPC_KC", s->size == 1 ? s->pk : 8192);
  if (kc > s->pk) kc = s->pk;
  s->kc = kc;
  s->ldn = phpc_pad_ld(s->n);
  s->lda_k = phpc_pad_ld(kc);

  if (s->size > 1 && kc >= 4096) s->kc_first = env_int("PHPC_KC_FIRST", 2048) / 128 * 128;
  if (s->size == 1 && kc_first_single > 0 && kc >= 2 * kc_first_single) {
    s->kc_first = kc_first_single / 128 * 128;
    s->grow_pct = grow_pct_single;
  }
  const int kf = s->kc_first, gp = s->grow_pct;
  const int nsteps = summa_schedule(gm, gk, n, s->r, s->c, s->pi, s->pj, kc, kf, gp, nullptr, 0, nullptr, nullptr);
  PHPC_REQUIRE(nsteps > 0, "empty SUMMA schedule");
  s->steps.resize(nsteps);
  summa_schedule(gm, gk, n, s->r, s->c, s->pi, s->pj, kc, kf, gp, s->steps.data(), nsteps, nullptr, nullptr);

  double t0 = now_s();
  phpc_b200_set_device(pick_device(s->rank));
  s->ctx = phpc_cur_ctx();
  PHPC_TRACE(s->rank, "create: device context", t0);
  {
    const char *t = getenv("PHPC_PANEL");
    s->transport = (s->size > 1 && !(t && !strcmp(t, "nccl"))) ? 1 : 0;
  }
  t0 = now_s();
  if (s->size > 1 && s->transport == 0) nccl_grid_get(grid_comm, s->size, s->rank, s->r, s->c, s->pi, s->pj);
  PHPC_TRACE(s->rank, "create: nccl communicators", t0);
  t0 = now_s();

  /* owned blocks: A chunks back to back ([m][pad(width)] each), B panels [pk][ldn] back to back, C [m][ldn] */
  for (const phpc_summa_step &st : s->steps)
    if (st.own_a) s->a_elems += (size_t)s->m * phpc_pad_ld(st.width);
  s->b_elems = (size_t)(s->lcm / s->r) * s->pk * s->ldn;
  s->c_elems = (size_t)s->m * s->ldn;
  /* The A and B stores are exported through CUDA IPC.  An IPC handle maps the whole backing
   * allocation, and cudaMalloc packs requests under 1 MiB into shared 2 MiB blocks, so the
   * stores are rounded up to whole 2 MiB granules: each is then its own, granule-aligned
   * allocation and the imported pointer is the store itself. */
  const size_t granule = (size_t)2 << 20;
  const size_t a_tag_off = ((s->a_elems ? s->a_elems : 2) * sizeof(double) + 255) / 256 * 256; /* 8-byte mapping tag after the data */
  const size_t b_tag_off = ((s->b_elems ? s->b_elems : 2) * sizeof(double) + 255) / 256 * 256;
  const size_t a_bytes = (a_tag_off + 256 + granule - 1) / granule * granule;
  const size_t b_bytes = (b_tag_off + 256 + granule - 1) / granule * granule;
  CUDA_CHECK(cudaMalloc(&s->dA, a_bytes));
  CUDA_CHECK(cudaMalloc(&s->dB, b_bytes));
  PHPC_REQUIRE(((uintptr_t)s->dA % granule) == 0 && ((uintptr_t)s->dB % granule) == 0, "IPC-exported stores must be 2 MiB aligned");
  const size_t c_tag_off = (s->c_elems * sizeof(double) + 255) / 256 * 256;
  const size_t c_bytes = (c_tag_off + 256 + granule - 1) / granule * granule;
  CUDA_CHECK(cudaMalloc(&s->dC, c_bytes));
  PHPC_REQUIRE(((uintptr_t)s->dC % granule) == 0, "IPC-exported stores must be 2 MiB aligned");
  CUDA_CHECK(cudaMemset(s->dC, 0, s->c_elems * sizeof(double)));

  /* where every chunk lives on the rank that owns it (needed to pull it) */
  s->root_a_off.assign(nsteps, -1);
  s->root_b_off.assign(nsteps, -1);
  {
    std::vector<phpc_summa_step> tmp(nsteps);
    for (int pj2 = 0; pj2 < s->c; ++pj2) {
      summa_schedule(gm, gk, n, s->r, s->c, s->pi, pj2, kc, kf, gp, tmp.data(), nsteps, nullptr, nullptr);
      for (int q = 0; q < nsteps; ++q)
        if (tmp[q].own_a) s->root_a_off[q] = tmp[q].a_off;
    }
    for (int pi2 = 0; pi2 < s->r; ++pi2) {
      summa_schedule(gm, gk, n, s->r, s->c, pi2, s->pj, kc, kf, gp, tmp.data(), nsteps, nullptr, nullptr);
      for (int q = 0; q < nsteps; ++q)
        if (tmp[q].own_b) s->root_b_off[q] = tmp[q].b_off;
    }
  }
  if (s->size > 1 && s->transport == 1) {
    /* exchange CUDA IPC handles of the A and B stores over the control plane */
    /* Opening the handle of a cudaMalloc base pointer yields the mapping of that pointer (the
     * stores are whole 2 MiB granules, see above).  Every mapping is verified before use: the
     * exporter plants a random tag behind its data, the importer must read the same tag through
     * its mapping, otherwise the process aborts instead of multiplying the wrong memory. */
    struct Handles {
      cudaIpcMemHandle_t a, b, c;
      unsigned long long a_tag, b_tag, c_tag;
      unsigned long long a_tag_off, b_tag_off, c_tag_off;
    };
    std::vector<Handles> all(s->size);
    memset(all.data(), 0, sizeof(Handles) * all.size());
    CUDA_CHECK(cudaIpcGetMemHandle(&all[s->rank].a, s->dA));
    CUDA_CHECK(cudaIpcGetMemHandle(&all[s->rank].b, s->dB));
    CUDA_CHECK(cudaIpcGetMemHandle(&all[s->rank].c, s->dC));
    {
      const unsigned long long salt = (unsigned long long)(now_s() * 1e6) ^ ((unsigned long long)getpid() << 32);
      all[s->rank].a_tag = salt * 0x9E3779B97F4A7C15ull + (unsigned long long)(uintptr_t)s->dA;
      all[s->rank].b_tag = salt * 0xBF58476D1CE4E5B9ull + (unsigned long long)(uintptr_t)s->dB;
      all[s->rank].c_tag = salt * 0x94D049BB133111EBull + (unsigned long long)(uintptr_t)s->dC;
      all[s->rank].a_tag_off = a_tag_off;
      all[s->rank].b_tag_off = b_tag_off;
      all[s->rank].c_tag_off = c_tag_off;
      CUDA_CHECK(cudaMemcpy((char *)s->dC + c_tag_off, &all[s->rank].c_tag, 8, cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMemcpy((char *)s->dA + a_tag_off, &all[s->rank].a_tag, 8, cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMemcpy((char *)s->dB + b_tag_off, &all[s->rank].b_tag, 8, cudaMemcpyHostToDevice));
    }
    for (int root = 0; root < s->size; ++root) MPI_Bcast(&all[root], (int)sizeof(Handles), MPI_BYTE, root, grid_comm);
    s->peerA.assign(s->c, nullptr);
    s->peerB.assign(s->r, nullptr);
    s->peerA_base.assign(s->c, nullptr);
    s->peerB_base.assign(s->r, nullptr);
    for (int pj2 = 0; pj2 < s->c; ++pj2)
      if (pj2 != s->pj) {
        const Handles &h = all[s->pi * s->c + pj2];
        CUDA_CHECK(cudaIpcOpenMemHandle(&s->peerA_base[pj2], h.a, cudaIpcMemLazyEnablePeerAccess));
        s->peerA[pj2] = (double *)s->peerA_base[pj2];
        unsigned long long seen = 0;
        CUDA_CHECK(cudaMemcpy(&seen, (char *)s->peerA_base[pj2] + h.a_tag_off, 8, cudaMemcpyDeviceToHost));
        PHPC_REQUIRE(seen == h.a_tag, "CUDA IPC mapping of a peer's A store does not show the peer's tag");
      }
    for (int pi2 = 0; pi2 < s->r; ++pi2)
      if (pi2 != s->pi) {
        const Handles &h = all[pi2 * s->c + s->pj];
        CUDA_CHECK(cudaIpcOpenMemHandle(&s->peerB_base[pi2], h.b, cudaIpcMemLazyEnablePeerAccess));
        s->peerB[pi2] = (double *)s->peerB_base[pi2];
        unsigned long long seen = 0;
        CUDA_CHECK(cudaMemcpy(&seen, (char *)s->peerB_base[pi2] + h.b_tag_off, 8, cudaMemcpyDeviceToHost));
        PHPC_REQUIRE(seen == h.b_tag, "CUDA IPC mapping of a peer's B store does not show the peer's tag");
      }
    if (s->rank == 0) { /* the gather root reads every C block */
      s->peerC.assign(s->size, nullptr);
      for (int i = 1; i < s->size; ++i) {
        CUDA_CHECK(cudaIpcOpenMemHandle(&s->peerC[i], all[i].c, cudaIpcMemLazyEnablePeerAccess));
        unsigned long long seen = 0;
        CUDA_CHECK(cudaMemcpy(&seen, (char *)s->peerC[i] + all[i].c_tag_off, 8, cudaMemcpyDeviceToHost));
        PHPC_REQUIRE(seen == all[i].c_tag, "CUDA IPC mapping of a peer's C block does not show the peer's tag");
      }
    }
  }

  PHPC_TRACE(s->rank, "create: blocks + ipc", t0);
  s->nbuf = env_int("PHPC_NBUF", 3);
  if (s->nbuf < 2) s->nbuf = 2;
  {
    /* Alternative schedule, PHPC_SCHEDULE=prefetch-all: stationary C with an all-gather prefetch.  The ring holds
     * EVERY chunk of the run, so all panel transfers are issued up front (an all-gather of the A panels along the process row
     * and of the B panels along the process column, expressed as copy-engine pulls) and the GEMMs consume the chunks as they
     * land; no slot is ever reused, so no transfer ever waits for a GEMM.  Costs (K/c + K/r) x block bytes of HBM instead of 3
     * chunks.  Default: the 3-slot ring (transfer of chunk q+1, q+2 under GEMM q). */
    const char *sch = getenv("PHPC_SCHEDULE");
    if (sch && !strcmp(sch, "prefetch-all") && s->nbuf < nsteps) s->nbuf = nsteps;
  }
  if (s->c > 1) {
    s->ringA_elems = (size_t)s->m * s->lda_k;
    CUDA_CHECK(cudaMalloc(&s->ringA, s->ringA_elems * s->nbuf * sizeof(double)));
  }
  if (s->r > 1) {
    s->ringB_elems = (size_t)kc * s->ldn;
    CUDA_CHECK(cudaMalloc(&s->ringB, s->ringB_elems * s->nbuf * sizeof(double)));
  }
  if (s->size > 1 && s->transport == 0 && env_int("PHPC_NCCL_REGISTER", 0)) {
    /* User-buffer registration of everything a broadcast reads or writes: with registered buffers on every rank NCCL can
     * move the panel straight between the user buffers over NVLink instead of through its own staging FIFOs.  Opt-in. */
    auto reg = [&](ncclComm_t comm, double *buf, size_t elems) {
      if (!comm || !buf || !elems) return;
      void *h = nullptr;
      NCCL_CHECK(ncclCommRegister(comm, buf, elems * sizeof(double), &h));
      s->nccl_regs.push_back({comm, h});
    };
    s->nccl_generation = g_nccl.generation;
    if (s->c > 1) {
      reg(g_nccl.row, s->dA, s->a_elems);
      reg(g_nccl.row, s->ringA, s->ringA_elems * s->nbuf);
    }
    if (s->r > 1) {
      reg(g_nccl.col, s->dB, s->b_elems);
      reg(g_nccl.col, s->ringB, s->ringB_elems * s->nbuf);
    }
  }
  s->ev_bcast.resize(s->nbuf);
  s->ev_bcast2.resize(s->nbuf);
  s->ev_free.resize(s->nbuf);
  for (int b = 0; b < s->nbuf; ++b) {
    CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_bcast[b], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_bcast2[b], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_free[b], cudaEventDisableTiming));
  }
  s->ev_g0.resize(nsteps);
  s->ev_g1.resize(nsteps);
  const bool ipc_events = s->size > 1 && s->transport == 1;
  const int nup = ipc_events ? PHPC_MAX_HOST_BANDS * nsteps : nsteps;
  s->ev_up.resize(nup);
  for (int q = 0; q < nsteps; ++q) {
    CUDA_CHECK(cudaEventCreate(&s->ev_g0[q]));
    CUDA_CHECK(cudaEventCreate(&s->ev_g1[q]));
  }
  for (int q = 0; q < nup; ++q) CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_up[q], cudaEventDisableTiming | (ipc_events ? cudaEventInterprocess : 0)));
  s->peer_up_a.assign(nup, nullptr);
  s->peer_up_b.assign(nsteps, nullptr);
  if (s->size > 1 && s->transport == 1) {
    /* "owned chunks of step q are in my store" as an event the peers that pull the chunk can wait for on THEIR streams:
     * host-sourced runs upload chunk by chunk under the GEMMs instead of "upload everything, synchronise, barrier" */
    std::vector<cudaIpcEventHandle_t> mine(nup), theirs(nup);
    for (int q = 0; q < nup; ++q) CUDA_CHECK(cudaIpcGetEventHandle(&mine[q], s->ev_up[q]));
    for (int root = 0; root < s->size; ++root) {
      if (root == s->rank) theirs = mine;
      MPI_Bcast(theirs.data(), (int)(sizeof(cudaIpcEventHandle_t) * nup), MPI_BYTE, root, grid_comm);
      if (root == s->rank) continue;
      const int ri = root / s->c, rj = root % s->c;
      for (int g = 0; g < nup; ++g) {
        const int q = g % nsteps;
        const phpc_summa_step &st = s->steps[q];
        if (ri == s->pi && rj == st.a_root && !st.own_a) CUDA_CHECK(cudaIpcOpenEventHandle(&s->peer_up_a[g], theirs[g]));
        if (g < nsteps && rj == s->pj && ri == st.b_root && !st.own_b) CUDA_CHECK(cudaIpcOpenEventHandle(&s->peer_up_b[q], theirs[q]));
      }
    }
  }
  CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_cup, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreate(&s->ev_begin));
  CUDA_CHECK(cudaEventCreate(&s->ev_end));
  CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_user, cudaEventDisableTiming));
  return s;
}

extern "C" void phpc_summa_destroy(phpc_summa *s) {
  if (!s) return;
  const double t0 = now_s();
  const int rank_dbg = s->rank;
  CUDA_CHECK(cudaSetDevice(s->ctx->device));
  CUDA_CHECK(cudaDeviceSynchronize());
  if (s->size > 1 && s->transport == 1) {
    MPI_Barrier(s->grid_comm); /* nobody is still pulling from the stores freed below */
    for (void *p : s->peerA_base)
      if (p) cudaIpcCloseMemHandle(p);
    for (void *p : s->peerB_base)
      if (p) cudaIpcCloseMemHandle(p);
    for (void *p : s->peerC)
      if (p) cudaIpcCloseMemHandle(p);
    MPI_Barrier(s->grid_comm); /* every importer has unmapped before the exporter frees */
  }
  for (cudaEvent_t e : s->ev_bcast2) cudaEventDestroy(e);
  if (g_nccl.ready && s->nccl_generation == g_nccl.generation) /* else the communicators (and the registrations) are gone */
    for (const phpc_summa::NcclReg &g : s->nccl_regs) ncclCommDeregister(g.comm, g.handle);
  cudaFree(s->dA);
  cudaFree(s->dB);
  cudaFree(s->dC);
  if (s->ringA) cudaFree(s->ringA);
  if (s->ringB) cudaFree(s->ringB);
  if (s->gather_stage) cudaFree(s->gather_stage);
  if (s->dC0) cudaFree(s->dC0);
  if (s->bcache_tb) cudaFree(s->bcache_tb);
  if (s->bcache_e) cudaFree(s->bcache_e);
  for (cudaEvent_t e : s->ev_bcast) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_free) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_g0) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_g1) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_up) cudaEventDestroy(e);
  for (cudaEvent_t e : s->peer_up_a)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : s->peer_up_b)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_c0) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_band) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_bg0) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_bg1) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i)
    if (s->d2h[i]) cudaStreamDestroy(s->d2h[i]);
  cudaEventDestroy(s->ev_cup);
  cudaEventDestroy(s->ev_begin);
  cudaEventDestroy(s->ev_end);
  cudaEventDestroy(s->ev_user);
  delete s;
  PHPC_TRACE(rank_dbg, "destroy", t0);
}

extern "C" void phpc_summa_chunks(const phpc_summa *s, int *kc, int *kc_first, int *steps) {
  if (kc) *kc = s->kc;
  if (kc_first) *kc_first = s->kc_first;
  if (steps) *steps = (int)s->steps.size();
}

extern "C" void phpc_summa_global(const phpc_summa *s, int mkn[3]) {
  mkn[0] = s->gM;
  mkn[1] = s->gK;
  mkn[2] = s->N;
}

extern "C" void phpc_summa_geometry(const phpc_summa *s, int dims[2], int coords[2], int block[2]) {
  dims[0] = s->r;
  dims[1] = s->c;
  coords[0] = s->pi;
  coords[1] = s->pj;
  block[0] = s->m;
  block[1] = s->n;
}

extern "C" void phpc_summa_zero_c(phpc_summa *s) {
  CUDA_CHECK(cudaMemsetAsync(s->dC, 0, s->c_elems * sizeof(double), s->ctx->compute));
  CUDA_CHECK(cudaStreamSynchronize(s->ctx->compute));
}

/* owned chunks of step q out of FULL N x N host matrices: the windows reference :42-44 point into */
static void upload_step(phpc_summa *s, int qi, const double *A, const double *B, cudaStream_t st) {
  const phpc_summa_step &q = s->steps[qi];
  const size_t N = (size_t)s->N, K = (size_t)s->gK;
  if (q.own_a) {
    const double *src = A + (size_t)s->pi * s->m * K + (size_t)q.k0;
    const size_t ld = phpc_pad_ld(q.width);
    CUDA_CHECK(cudaMemcpy2DAsync(s->dA + q.a_off, ld * sizeof(double), src, K * sizeof(double), (size_t)q.width * sizeof(double), s->m,
                                 cudaMemcpyHostToDevice, st));
  }
  if (q.own_b) {
    const double *src = B + (size_t)q.k0 * N + (size_t)s->pj * s->n;
    CUDA_CHECK(cudaMemcpy2DAsync(s->dB + q.b_off, s->ldn * sizeof(double), src, N * sizeof(double), (size_t)s->n * sizeof(double), q.width,
                                 cudaMemcpyHostToDevice, st));
  }
}

static void upload_c(phpc_summa *s, const double *C, cudaStream_t st) {
  const size_t N = (size_t)s->N;
  if (C) {
    const double *src = C + (size_t)s->pi * s->m * N + (size_t)s->pj * s->n;
    CUDA_CHECK(cudaMemcpy2DAsync(s->dC, s->ldn * sizeof(double), src, N * sizeof(double), (size_t)s->n * sizeof(double), s->m,
                                 cudaMemcpyHostToDevice, st));
  } else {
    CUDA_CHECK(cudaMemsetAsync(s->dC, 0, s->c_elems * sizeof(double), st));
  }
}

extern "C" void phpc_summa_upload(phpc_summa *s, const double *A, const double *B, const double *C) {
  cudaStream_t st = s->ctx->copy;
  for (int q = 0; q < (int)s->steps.size(); ++q) upload_step(s, q, A, B, st);
  upload_c(s, C, st);
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (s->size > 1) MPI_Barrier(s->grid_comm); /* every store is valid before anyone pulls from it */
}

extern "C" void phpc_summa_fill(phpc_summa *s, int kind, unsigned long long seed_a, unsigned long long seed_b) {
  cudaStream_t st = s->ctx->compute;
  for (const phpc_summa_step &q : s->steps) {
    if (q.own_a)
      phpc_fill_device(s->dA + q.a_off, phpc_pad_ld(q.width), s->m, q.width, (long long)s->pi * s->m, q.k0, s->gK, kind, seed_a, st);
    if (q.own_b) phpc_fill_device(s->dB + q.b_off, s->ldn, q.width, s->n, q.k0, (long long)s->pj * s->n, s->N, kind, seed_b, st);
  }
  CUDA_CHECK(cudaMemsetAsync(s->dC, 0, s->c_elems * sizeof(double), st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (s->size > 1) MPI_Barrier(s->grid_comm); /* every store is valid before anyone pulls from it */
}

__global__ void add_inplace_kernel(double *__restrict__ c, const double *__restrict__ c0, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c[i] += c0[i];
}

/* ------------------------------------------------------------------------- */
/* the k-loop                                                                 */
/* ------------------------------------------------------------------------- */
/*
 * hA/hB/hC != NULL: host-sourced run (the reference-facing entry points).  The owned
 * chunks of step q are uploaded on the copy stream while earlier steps compute, so
 * the H2D traffic the reference pays in full before every kernel (src/phpc_gemm.cu:
 * 111-113) hides under the GEMMs; hC (may be NULL = zeros) is uploaded first.
 */
static void summa_run(phpc_summa *s, int backend, int ctas, void *user_stream, phpc_summa_stats *stats, const double *hA,
                      const double *hB, const double *hC, bool host_src) {
  DeviceCtx *ctx = s->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t comm = ctx->comm, comp = ctx->compute, copy = ctx->copy;
  const int nsteps = (int)s->steps.size();
  const int comm_sms = env_int("PHPC_COMM_SMS", 0); /* SMs left free for NCCL while broadcasts are in flight */
  int launches = 0, broadcasts = 0;
  long long bytes_rx = 0;

  if (user_stream) { /* start after the caller's earlier work */
    CUDA_CHECK(cudaEventRecord(s->ev_user, (cudaStream_t)user_stream));
    CUDA_CHECK(cudaStreamWaitEvent(comm, s->ev_user, 0));
    CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_user, 0));
    CUDA_CHECK(cudaStreamWaitEvent(copy, s->ev_user, 0));
  }
  CUDA_CHECK(cudaEventRecord(s->ev_begin, comp));
  CUDA_CHECK(cudaStreamWaitEvent(comm, s->ev_begin, 0));
  const bool any_comm = (s->r > 1 || s->c > 1);
  const bool pull = any_comm && s->transport == 1;
  bool late_c = false;
  cudaStream_t comm2 = ctx->comm2;
  if (pull) CUDA_CHECK(cudaStreamWaitEvent(comm2, s->ev_begin, 0));
  if (host_src) {
    CUDA_CHECK(cudaStreamWaitEvent(copy, s->ev_begin, 0));
    if (pull) {
      /* Peers pull straight from this rank's store.  All uploads are enqueued now, in step order, each followed by an
       * interprocess event; the host barrier only orders the ENQUEUE of those
       * records before the peers enqueue their waits - the copies themselves run under the GEMMs of earlier steps. */
      for (int q = 0; q < nsteps; ++q) {
        upload_step(s, q, hA, hB, copy);
        CUDA_CHECK(cudaEventRecord(s->ev_up[q], copy));
      }
      /* The caller's C block must not hold up the first GEMM (it is as large as everything step 0 needs): the loop runs on a
       * zeroed block, the caller's block follows the operands into a side buffer and is added once at the end
       * (C0 + sum of the chunk products instead of ((C0 + P0) + P1) + ...: one rounding of difference at most). */
      CUDA_CHECK(cudaMemsetAsync(s->dC, 0, s->c_elems * sizeof(double), comp));
      if (hC) {
        if (!s->dC0) CUDA_CHECK(cudaMalloc(&s->dC0, s->c_elems * sizeof(double)));
        const size_t Nn = (size_t)s->N;
        CUDA_CHECK(cudaMemcpy2DAsync(s->dC0, s->ldn * sizeof(double), hC + (size_t)s->pi * s->m * Nn + (size_t)s->pj * s->n, Nn * sizeof(double),
                                     (size_t)s->n * sizeof(double), s->m, cudaMemcpyHostToDevice, copy));
        if (s->ldn != s->n) /* the padding columns of the side buffer are added too: keep them defined */
          CUDA_CHECK(cudaMemset2DAsync(s->dC0 + s->n, s->ldn * sizeof(double), 0, (size_t)(s->ldn - s->n) * sizeof(double), s->m, copy));
      }
      CUDA_CHECK(cudaEventRecord(s->ev_cup, copy));
      late_c = hC != nullptr;
      MPI_Barrier(s->grid_comm);
    } else {
      upload_c(s, hC, copy);
      CUDA_CHECK(cudaEventRecord(s->ev_cup, copy));
    }
    if (!late_c && !(pull)) CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_cup, 0));
  }
  /* stage-in of step q = upload of the owned chunks (host-sourced) + the panel transfers */
  auto stage_in = [&](int q) {
    const phpc_summa_step &st = s->steps[q];
    const int slot = q % s->nbuf;
    if (pull) {
      /* copy-engine pull of the chunks this rank does not own, from the owner's HBM over NVLink */
      if (s->c > 1 && !st.own_a) {
        const size_t count = (size_t)s->m * phpc_pad_ld(st.width);
        if (q >= s->nbuf) CUDA_CHECK(cudaStreamWaitEvent(comm, s->ev_free[slot], 0));
        if (host_src) CUDA_CHECK(cudaStreamWaitEvent(comm, s->peer_up_a[q], 0)); /* the owner's upload of this chunk has landed */
        CUDA_CHECK(cudaMemcpyAsync(s->ringA + (size_t)slot * s->ringA_elems, s->peerA[st.a_root] + s->root_a_off[q], count * 8,
                                   cudaMemcpyDeviceToDevice, comm));
        ++broadcasts;
        bytes_rx += (long long)count * 8;
      }
      CUDA_CHECK(cudaEventRecord(s->ev_bcast[slot], comm));
      if (s->r > 1 && !st.own_b) {
        const size_t count = (size_t)st.width * s->ldn;
        if (q >= s->nbuf) CUDA_CHECK(cudaStreamWaitEvent(comm2, s->ev_free[slot], 0));
        if (host_src) CUDA_CHECK(cudaStreamWaitEvent(comm2, s->peer_up_b[q], 0));
        CUDA_CHECK(cudaMemcpyAsync(s->ringB + (size_t)slot * s->ringB_elems, s->peerB[st.b_root] + s->root_b_off[q], count * 8,
                                   cudaMemcpyDeviceToDevice, comm2));
        ++broadcasts;
        bytes_rx += (long long)count * 8;
      }
      CUDA_CHECK(cudaEventRecord(s->ev_bcast2[slot], comm2));
      return;
    }
    const bool uploaded = host_src && (st.own_a || st.own_b);
    if (uploaded) {
      upload_step(s, q, hA, hB, copy);
      CUDA_CHECK(cudaEventRecord(s->ev_up[q], copy));
    }
    if (!any_comm) return;
    if (uploaded) CUDA_CHECK(cudaStreamWaitEvent(comm, s->ev_up[q], 0));
    if (q >= s->nbuf) CUDA_CHECK(cudaStreamWaitEvent(comm, s->ev_free[slot], 0)); /* GEMM q-nbuf released the slot */
    NCCL_CHECK(ncclGroupStart());
    if (s->c > 1) {
      const size_t count = (size_t)s->m * phpc_pad_ld(st.width);
      double *buf = st.own_a ? s->dA + st.a_off : s->ringA + (size_t)slot * s->ringA_elems;
      NCCL_CHECK(ncclBroadcast(buf, buf, count, ncclDouble, st.a_root, g_nccl.row, comm));
      ++broadcasts;
      if (!st.own_a) bytes_rx += (long long)count * 8;
    }
    if (s->r > 1) {
      const size_t count = (size_t)st.width * s->ldn;
      double *buf = st.own_b ? s->dB + st.b_off : s->ringB + (size_t)slot * s->ringB_elems;
      NCCL_CHECK(ncclBroadcast(buf, buf, count, ncclDouble, st.b_root, g_nccl.col, comm));
      ++broadcasts;
      if (!st.own_b) bytes_rx += (long long)count * 8;
    }
    NCCL_CHECK(ncclGroupEnd());
    CUDA_CHECK(cudaEventRecord(s->ev_bcast[slot], comm));
  };

  int issued = 0;
  stage_in(issued++);
  for (int q = 0; q < nsteps; ++q) {
    const phpc_summa_step &st = s->steps[q];
    const int slot = q % s->nbuf;
    if (any_comm) {
      CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_bcast[slot], 0));
      if (pull) CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_bcast2[slot], 0));
      if (pull && host_src && (st.own_a || st.own_b)) CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_up[q], 0));
    } else if (host_src) {
      CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_up[q], 0));
    }
    const double *a = st.own_a ? s->dA + st.a_off : s->ringA + (size_t)slot * s->ringA_elems;
    const double *b = st.own_b ? s->dB + st.b_off : s->ringB + (size_t)slot * s->ringB_elems;
    const long long lda = phpc_pad_ld(st.width);
    CUDA_CHECK(cudaEventRecord(s->ev_g0[q], comp));
    if (backend == PHPC_BACKEND_CUBLAS) {
      phpc_launch_cublas(ctx, a, lda, b, s->ldn, s->dC, s->ldn, s->m, st.width, s->n, comp);
      ++launches;
    } else if (backend == PHPC_BACKEND_OZAKI) {
      /* NCCL transport: the persistent tcgen05 grid waits for ALL its CTAs to be resident, so a broadcast kernel that cannot
       * get an SM until the GEMM ends (and whose peers then wait for it) serialises transfer and compute completely
       * (measured: 73 % exposed on 8 GPUs).  While broadcasts are in flight the GEMM leaves SMs free for them. */
      int use = ctas;
      if (any_comm && !pull && q + 1 < nsteps) {
        const int base = (ctas <= 1 || ctas > ctx->sm_count) ? ctx->sm_count : ctas;
        const int reserve = env_int("PHPC_COMM_SMS", 8);
        use = base - reserve > 2 ? base - reserve : 2;
      }
      launches += phpc_launch_ozaki(ctx, a, lda, b, s->ldn, s->dC, s->ldn, s->m, st.width, s->n, use, comp);
    } else {
      int use = ctas;
      if (any_comm && !pull && comm_sms > 0 && q + 1 < nsteps) {
        const int base = (ctas <= 1 || ctas > ctx->sm_count) ? ctx->sm_count : ctas;
        use = base - comm_sms > 1 ? base - comm_sms : 2;
      }
      launches += phpc_launch_dmma(ctx, a, lda, b, s->ldn, s->dC, s->ldn, s->m, st.width, s->n, use, comp);
    }
    CUDA_CHECK(cudaEventRecord(s->ev_g1[q], comp));
    if (any_comm) CUDA_CHECK(cudaEventRecord(s->ev_free[slot], comp));
    /* prefetch: the stage-in of the next nbuf-1 steps runs under this GEMM */
    while (issued < nsteps && issued < q + s->nbuf) stage_in(issued++);
  }
  if (late_c) {
    CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_cup, 0));
    add_inplace_kernel<<<ctx->sm_count * 8, 256, 0, comp>>>(s->dC, s->dC0, s->c_elems);
    CUDA_CHECK(cudaGetLastError());
  }
  CUDA_CHECK(cudaEventRecord(s->ev_end, comp));
  if (user_stream) CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)user_stream, s->ev_end, 0));

  if (stats) {
    CUDA_CHECK(cudaEventSynchronize(s->ev_end));
    CUDA_CHECK(cudaStreamSynchronize(comm));
    CUDA_CHECK(cudaStreamSynchronize(comm2));
    CUDA_CHECK(cudaStreamSynchronize(copy));
    float total = 0.f, gemm = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&total, s->ev_begin, s->ev_end));
    for (int q = 0; q < nsteps; ++q) {
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, s->ev_g0[q], s->ev_g1[q]));
      gemm += ms;
    }
    stats->total_ms = total;
    stats->gemm_ms = gemm;
    stats->exposed_ms = total - gemm;
    stats->steps = nsteps;
    stats->launches = launches;
    stats->broadcasts = broadcasts;
    stats->bytes_received = bytes_rx;
  }
}

/* per-step GEMM start offsets (from the run's begin event) and durations of the LAST run, ms */
extern "C" int phpc_summa_timeline(phpc_summa *s, float *start_ms, float *dur_ms, int max_steps) {
  CUDA_CHECK(cudaSetDevice(s->ctx->device));
  CUDA_CHECK(cudaEventSynchronize(s->ev_end));
  const int n = (int)s->steps.size() < max_steps ? (int)s->steps.size() : max_steps;
  for (int q = 0; q < n; ++q) {
    CUDA_CHECK(cudaEventElapsedTime(&start_ms[q], s->ev_begin, s->ev_g0[q]));
    CUDA_CHECK(cudaEventElapsedTime(&dur_ms[q], s->ev_g0[q], s->ev_g1[q]));
  }
  return n;
}

extern "C" void phpc_summa_run(phpc_summa *s, int backend, int ctas, void *user_stream, phpc_summa_stats *stats) {
  summa_run(s, backend, ctas, user_stream, stats, nullptr, nullptr, nullptr, false);
}

/* ------------------------------------------------------------------------- */
/* band-pipelined host-sourced run, one GPU                                   */
/* ------------------------------------------------------------------------- */
/* start of a banded run: B is about to be uploaded again, nothing derived from it is valid */
static void bcache_arm(phpc_summa *s, int backend) {
  if (backend != PHPC_BACKEND_OZAKI) {
    s->bcache.clear();
    return;
  }
  const int nsteps = (int)s->steps.size();
  if (!s->bcache_tb) {
    size_t bytes = 0, ints = 0;
    for (const phpc_summa_step &st : s->steps) {
      size_t e = 0;
      bytes += phpc_ozaki_bcache_bytes(st.width, s->n, &e);
      ints += e;
    }
    CUDA_CHECK(cudaMalloc(&s->bcache_tb, bytes));
    CUDA_CHECK(cudaMalloc(&s->bcache_e, ints * sizeof(int)));
  }
  s->bcache.assign(nsteps, OzBCache());
  size_t off = 0, eoff = 0;
  for (int q = 0; q < nsteps; ++q) {
    size_t e = 0;
    s->bcache[q].TB = s->bcache_tb + off;
    s->bcache[q].eB = s->bcache_e + eoff;
    s->bcache[q].ready = false;
    off += phpc_ozaki_bcache_bytes(s->steps[q].width, s->n, &e);
    eoff += e;
  }
}

static int launch_local_gemm(phpc_summa *s, int backend, int ctas, const double *a, long long lda, const double *b, double *c, int rows,
                             int width, cudaStream_t st, int step = -1) {
  DeviceCtx *ctx = s->ctx;
  if (backend == PHPC_BACKEND_CUBLAS) {
    phpc_launch_cublas(ctx, a, lda, b, s->ldn, c, s->ldn, rows, width, s->n, st);
    return 1;
  }
  if (backend == PHPC_BACKEND_OZAKI)
    return phpc_launch_ozaki(ctx, a, lda, b, s->ldn, c, s->ldn, rows, width, s->n, ctas, st,
                             (step >= 0 && step < (int)s->bcache.size()) ? &s->bcache[step] : nullptr);
  return phpc_launch_dmma(ctx, a, lda, b, s->ldn, c, s->ldn, rows, width, s->n, ctas, st);
}

/* CUDA side of phpc::band_execute (host_band_exec.h): plan stream -> CUDA stream, one event per dependency, timing
 * events around every local GEMM.  The walk over the operation list, with all its pointer arithmetic, is the shared
 * header and runs on the CPU in tests/test_band_executor.py (deferred streams in random order, host buffers as HBM). */
struct CudaBandBackend {
  phpc_summa *s;
  int backend, ctas;
  cudaStream_t streams[3];
  std::vector<cudaEvent_t> deps, g0, g1;
};
static void cb_copy2d(void *self, int stream, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes,
                      size_t rows, int host_to_device) {
  CudaBandBackend *b = (CudaBandBackend *)self;
  CUDA_CHECK(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows,
                               host_to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, b->streams[stream]));
}
static void *cb_record(void *self, int stream) {
  CudaBandBackend *b = (CudaBandBackend *)self;
  cudaEvent_t e;
  CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(e, b->streams[stream]));
  b->deps.push_back(e);
  return (void *)e;
}
static void cb_wait(void *self, int stream, void *event) {
  CudaBandBackend *b = (CudaBandBackend *)self;
  PHPC_REQUIRE(event != nullptr, "host plan dependency does not point at an earlier, recorded operation");
  CUDA_CHECK(cudaStreamWaitEvent(b->streams[stream], (cudaEvent_t)event, 0));
}
static int cb_gemm(void *self, int stream, const double *a, long long lda, const double *bm, long long ldb, double *c, long long ldc,
                   int rows, int width, int n, int step) {
  CudaBandBackend *b = (CudaBandBackend *)self;
  (void)ldb;
  (void)ldc;
  (void)n; /* the B store and the C block of the object carry their own leading dimension and width */
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventRecord(e0, b->streams[stream]));
  const int launches = launch_local_gemm(b->s, b->backend, b->ctas, a, lda, bm, c, rows, width, b->streams[stream], step);
  CUDA_CHECK(cudaEventRecord(e1, b->streams[stream]));
  b->g0.push_back(e0);
  b->g1.push_back(e1);
  return launches;
}

static void cb_zero(void *self, int stream, double *dst, size_t count) {
  CudaBandBackend *b = (CudaBandBackend *)self;
  CUDA_CHECK(cudaMemsetAsync(dst, 0, count * sizeof(double), b->streams[stream]));
}
static void cb_add(void *self, int stream, double *dst, const double *src, size_t count) {
  CudaBandBackend *b = (CudaBandBackend *)self;
  add_inplace_kernel<<<b->s->ctx->sm_count * 4, 256, 0, b->streams[stream]>>>(dst, src, count);
  CUDA_CHECK(cudaGetLastError());
}

static void summa_run_host_banded(phpc_summa *s, int backend, int ctas, const double *hA, const double *hB, double *hC, int bands,
                                  phpc_summa_stats *stats) {
  PHPC_REQUIRE(s->size == 1, "the band pipeline is the single-rank host path");
  DeviceCtx *ctx = s->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  const int nsteps = (int)s->steps.size();
  const int nops = phpc::host_plan(s->m, nsteps, bands, 128, nullptr, 0);
  std::vector<phpc_host_op> ops(nops);
  phpc::host_plan(s->m, nsteps, bands, 128, ops.data(), nops);

  bcache_arm(s, backend);
  CudaBandBackend cb;
  cb.s = s;
  cb.backend = backend;
  cb.ctas = ctas;
  cb.streams[0] = ctx->copy;
  cb.streams[1] = ctx->compute;
  cb.streams[2] = ctx->comm; /* idle on one GPU: it carries the downloads */
  phpc::BandBackend be = {&cb, cb_copy2d, cb_record, cb_wait, cb_gemm, cb_zero, cb_add};
  if (!s->dC0) CUDA_CHECK(cudaMalloc(&s->dC0, s->c_elems * sizeof(double)));
  phpc::BandGeom g;
  g.N = s->N;
  g.lda_host = s->gK;
  g.m = s->m;
  g.n = s->n;
  g.pi = s->pi;
  g.pj = s->pj;
  g.ldn = s->ldn;
  g.steps = s->steps.data();
  g.nsteps = nsteps;
  g.dA = s->dA;
  g.dB = s->dB;
  g.dC = s->dC;
  g.dC0 = s->dC0;

  CUDA_CHECK(cudaEventRecord(s->ev_begin, ctx->compute));
  CUDA_CHECK(cudaStreamWaitEvent(cb.streams[0], s->ev_begin, 0));
  CUDA_CHECK(cudaStreamWaitEvent(cb.streams[2], s->ev_begin, 0));
  const int launches = phpc::band_execute(g, ops.data(), nops, hA, hB, hC, be);
  PHPC_REQUIRE(launches >= 0, "unknown host plan operation");
  CUDA_CHECK(cudaEventRecord(s->ev_end, ctx->compute));
  for (int i = 0; i < 3; ++i) CUDA_CHECK(cudaStreamSynchronize(cb.streams[i]));
  if (stats) {
    float total = 0.f, gemm = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&total, s->ev_begin, s->ev_end));
    for (size_t i = 0; i < cb.g0.size(); ++i) {
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, cb.g0[i], cb.g1[i]));
      gemm += ms;
    }
    stats->total_ms = total;
    stats->gemm_ms = gemm;
    stats->exposed_ms = total - gemm;
    stats->steps = nsteps;
    stats->launches = launches;
    stats->broadcasts = 0;
    stats->bytes_received = 0;
  }
  for (cudaEvent_t e : cb.deps) CUDA_CHECK(cudaEventDestroy(e));
  for (cudaEvent_t e : cb.g0) CUDA_CHECK(cudaEventDestroy(e));
  for (cudaEvent_t e : cb.g1) CUDA_CHECK(cudaEventDestroy(e));
}

/* ------------------------------------------------------------------------- */
/* band-pipelined host-sourced run, several ranks                             */
/* ------------------------------------------------------------------------- */
struct ShareInfo { /* where rank 0's result matrix lives when every rank can map it (phpc_host_malloc_shared) */
  int shared;
  char name[64];
  unsigned long long off, bytes;
};

/*
 * The host side bounds the multi-rank call (8 GPUs, N = 32768: 3.2 GB up and 2.1 GB down per rank against 80 ms of GEMMs),
 * and with the K-outer loop the first byte of C can only leave when the last chunk has been multiplied, so uploads and
 * downloads ran one after the other.  Here the rank's C block is cut into row bands and the SUMMA k-loop runs once per band:
 * band b needs the A rows of that band only (uploaded and pulled per band; a band of a stored chunk is contiguous) and every
 * B chunk (uploaded once, pulled again per band: NVLink has the headroom), and it is final, and on its way to the host over
 * the D2H direction of the link, while the uploads and GEMMs of band b+1 proceed.  Per element the K chunks are still added
 * in ascending order.  Used when rank 0's C is node-shared memory (every rank then writes its bands straight into it).
 */
static void summa_run_host_multi_banded(phpc_summa *s, int backend, int ctas, const double *hA, const double *hB, double *hC, const ShareInfo &sh,
                                        int bands, phpc_summa_stats *stats) {
  DeviceCtx *ctx = s->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  cudaStream_t comm = ctx->comm, comm2 = ctx->comm2, comp = ctx->compute, copy = ctx->copy;
  const int nsteps = (int)s->steps.size();
  int rows_per_band = (s->m + bands - 1) / bands;
  rows_per_band = (rows_per_band + 127) / 128 * 128;
  const int nb = (s->m + rows_per_band - 1) / rows_per_band;
  PHPC_REQUIRE(nb <= PHPC_MAX_HOST_BANDS, "too many host row bands");
  const int total = nb * nsteps;
  const size_t N = (size_t)s->N, K = (size_t)s->gK;
  for (int i = 0; i < 2; ++i)
    if (!s->d2h[i]) CUDA_CHECK(cudaStreamCreateWithFlags(&s->d2h[i], cudaStreamNonBlocking));
  while ((int)s->ev_c0.size() < nb) {
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    s->ev_c0.push_back(a);
    s->ev_band.push_back(b);
  }
  while ((int)s->ev_bg0.size() < total) {
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a));
    CUDA_CHECK(cudaEventCreate(&b));
    s->ev_bg0.push_back(a);
    s->ev_bg1.push_back(b);
  }
  if (hC && !s->dC0) CUDA_CHECK(cudaMalloc(&s->dC0, s->c_elems * sizeof(double)));
  char *shared_base = nullptr;
  if (s->rank != 0) {
    const unsigned long long blk_off = sh.off + ((unsigned long long)s->pi * s->m * N + (unsigned long long)s->pj * s->n) * sizeof(double);
    const unsigned long long blk_len = ((unsigned long long)(s->m - 1) * N + s->n) * sizeof(double);
    shared_base = (char *)phpc_host_shared_map(sh.name, sh.bytes, blk_off, blk_len) + blk_off;
  }

  bcache_arm(s, backend);
  CUDA_CHECK(cudaEventRecord(s->ev_begin, comp));
  for (cudaStream_t st : {comm, comm2, copy, s->d2h[0], s->d2h[1]}) CUDA_CHECK(cudaStreamWaitEvent(st, s->ev_begin, 0));
  CUDA_CHECK(cudaMemsetAsync(s->dC, 0, s->c_elems * sizeof(double), comp));

  /* 1. every upload of the call, band after band, each followed by an interprocess event */
  for (int b = 0; b < nb; ++b) {
    const int row0 = b * rows_per_band, rows = (s->m - row0 < rows_per_band) ? s->m - row0 : rows_per_band;
    for (int q = 0; q < nsteps; ++q) {
      const phpc_summa_step &st = s->steps[q];
      if (st.own_a) {
        const size_t ld = phpc_pad_ld(st.width);
        CUDA_CHECK(cudaMemcpy2DAsync(s->dA + st.a_off + (size_t)row0 * ld, ld * sizeof(double), hA + ((size_t)s->pi * s->m + row0) * K + (size_t)st.k0,
                                     K * sizeof(double), (size_t)st.width * sizeof(double), rows, cudaMemcpyHostToDevice, copy));
      }
      if (b == 0 && st.own_b)
        CUDA_CHECK(cudaMemcpy2DAsync(s->dB + st.b_off, s->ldn * sizeof(double), hB + (size_t)st.k0 * N + (size_t)s->pj * s->n, N * sizeof(double),
                                     (size_t)s->n * sizeof(double), st.width, cudaMemcpyHostToDevice, copy));
      CUDA_CHECK(cudaEventRecord(s->ev_up[b * nsteps + q], copy));
    }
    if (hC) {
      CUDA_CHECK(cudaMemcpy2DAsync(s->dC0 + (size_t)row0 * s->ldn, s->ldn * sizeof(double), hC + ((size_t)s->pi * s->m + row0) * N + (size_t)s->pj * s->n,
                                   N * sizeof(double), (size_t)s->n * sizeof(double), rows, cudaMemcpyHostToDevice, copy));
      if (s->ldn != s->n)
        CUDA_CHECK(cudaMemset2DAsync(s->dC0 + (size_t)row0 * s->ldn + s->n, s->ldn * sizeof(double), 0, (size_t)(s->ldn - s->n) * sizeof(double), rows, copy));
    }
    CUDA_CHECK(cudaEventRecord(s->ev_c0[b], copy));
  }
  MPI_Barrier(s->grid_comm); /* orders the ENQUEUE of the records above before the peers enqueue their waits */

  /* 2. the k-loop, once per band */
  int launches = 0, broadcasts = 0;
  long long bytes_rx = 0;
  auto stage_in = [&](int g) {
    const int b = g / nsteps, q = g % nsteps, slot = g % s->nbuf;
    const phpc_summa_step &st = s->steps[q];
    const int row0 = b * rows_per_band, rows = (s->m - row0 < rows_per_band) ? s->m - row0 : rows_per_band;
    if (s->c > 1 && !st.own_a) {
      const size_t ld = phpc_pad_ld(st.width), count = (size_t)rows * ld;
      if (g >= s->nbuf) CUDA_CHECK(cudaStreamWaitEvent(comm, s->ev_free[slot], 0));
      CUDA_CHECK(cudaStreamWaitEvent(comm, s->peer_up_a[b * nsteps + q], 0));
      CUDA_CHECK(cudaMemcpyAsync(s->ringA + (size_t)slot * s->ringA_elems, s->peerA[st.a_root] + s->root_a_off[q] + (size_t)row0 * ld, count * 8,
                                 cudaMemcpyDeviceToDevice, comm));
      ++broadcasts;
      bytes_rx += (long long)count * 8;
    }
    CUDA_CHECK(cudaEventRecord(s->ev_bcast[slot], comm));
    if (s->r > 1 && !st.own_b) {
      const size_t count = (size_t)st.width * s->ldn;
      if (g >= s->nbuf) CUDA_CHECK(cudaStreamWaitEvent(comm2, s->ev_free[slot], 0));
      CUDA_CHECK(cudaStreamWaitEvent(comm2, s->peer_up_b[q], 0));
      CUDA_CHECK(cudaMemcpyAsync(s->ringB + (size_t)slot * s->ringB_elems, s->peerB[st.b_root] + s->root_b_off[q], count * 8, cudaMemcpyDeviceToDevice,
                                 comm2));
      ++broadcasts;
      bytes_rx += (long long)count * 8;
    }
    CUDA_CHECK(cudaEventRecord(s->ev_bcast2[slot], comm2));
  };
  int issued = 0;
  stage_in(issued++);
  for (int g = 0; g < total; ++g) {
    const int b = g / nsteps, q = g % nsteps, slot = g % s->nbuf;
    const phpc_summa_step &st = s->steps[q];
    const int row0 = b * rows_per_band, rows = (s->m - row0 < rows_per_band) ? s->m - row0 : rows_per_band;
    CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_bcast[slot], 0));
    CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_bcast2[slot], 0));
    if (st.own_a) CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_up[b * nsteps + q], 0));
    if (st.own_b) CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_up[q], 0));
    const long long lda = phpc_pad_ld(st.width);
    const double *a = st.own_a ? s->dA + st.a_off + (size_t)row0 * lda : s->ringA + (size_t)slot * s->ringA_elems;
    const double *bm = st.own_b ? s->dB + st.b_off : s->ringB + (size_t)slot * s->ringB_elems;
    CUDA_CHECK(cudaEventRecord(s->ev_bg0[g], comp));
    launches += launch_local_gemm(s, backend, ctas, a, lda, bm, s->dC + (size_t)row0 * s->ldn, rows, st.width, comp, q);
    CUDA_CHECK(cudaEventRecord(s->ev_bg1[g], comp));
    CUDA_CHECK(cudaEventRecord(s->ev_free[slot], comp));
    while (issued < total && issued < g + s->nbuf) stage_in(issued++);
    if (q == nsteps - 1) { /* the band is complete: add the caller's C band, send it home */
      if (hC) {
        CUDA_CHECK(cudaStreamWaitEvent(comp, s->ev_c0[b], 0));
        add_inplace_kernel<<<ctx->sm_count * 4, 256, 0, comp>>>(s->dC + (size_t)row0 * s->ldn, s->dC0 + (size_t)row0 * s->ldn, (size_t)rows * s->ldn);
        CUDA_CHECK(cudaGetLastError());
      }
      CUDA_CHECK(cudaEventRecord(s->ev_band[b], comp));
      const size_t row_bytes = (size_t)s->n * sizeof(double);
      if (s->rank != 0) { /* first into rank 0's result: that is the copy everybody waits for */
        CUDA_CHECK(cudaStreamWaitEvent(s->d2h[1], s->ev_band[b], 0));
        CUDA_CHECK(cudaMemcpy2DAsync(shared_base + (size_t)row0 * N * sizeof(double), N * sizeof(double), s->dC + (size_t)row0 * s->ldn,
                                     s->ldn * sizeof(double), row_bytes, rows, cudaMemcpyDeviceToHost, s->d2h[1]));
      }
      CUDA_CHECK(cudaStreamWaitEvent(s->d2h[0], s->ev_band[b], 0));
      CUDA_CHECK(cudaMemcpy2DAsync(hC + ((size_t)s->pi * s->m + row0) * N + (size_t)s->pj * s->n, N * sizeof(double), s->dC + (size_t)row0 * s->ldn,
                                   s->ldn * sizeof(double), row_bytes, rows, cudaMemcpyDeviceToHost, s->d2h[0]));
    }
  }
  CUDA_CHECK(cudaEventRecord(s->ev_end, comp));
  for (cudaStream_t st : {comp, comm, comm2, copy, s->d2h[0], s->d2h[1]}) CUDA_CHECK(cudaStreamSynchronize(st));
  MPI_Barrier(s->grid_comm); /* every band has landed in rank 0's C; nobody is still pulling from a store */
  if (stats) {
    float tot = 0.f, gemm = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&tot, s->ev_begin, s->ev_end));
    for (int g = 0; g < total; ++g) {
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, s->ev_bg0[g], s->ev_bg1[g]));
      gemm += ms;
    }
    stats->total_ms = tot;
    stats->gemm_ms = gemm;
    stats->exposed_ms = tot - gemm;
    stats->steps = total;
    stats->launches = launches;
    stats->broadcasts = broadcasts;
    stats->bytes_received = bytes_rx;
  }
}

/* Row bands of the single-GPU host-sourced run: PHPC_HOST_BANDS, else 4 once the block is big enough for the C transfers to
 * matter (>= 8192 rows), else 1.  Bands are 1/2, 1/4, 1/8, 1/8 of the block (host_band_exec.h).  Why, at N = 32768 (PCIe
 * ~55 GB/s, 8 K chunks of 4096, GEMM 0.64 s): per chunk the first band uploads its A rows + one B chunk (0.54 + 1.07 GB =
 * 29 ms) under 40 ms of GEMM, and the slack it gains over the 8 chunks pays for the upload of its own C rows (4.3 GB) before
 * they are added; equal quarters were upload bound in the first band (and waited for the C rows up front: 86 TFLOP/s).  The
 * exposed tail is the last band's download (1.07 GB, 20 ms).  Every band repeats the split of the B chunks (+0.4 % each). */
static int host_bands(const phpc_summa *s) {
  if (s->size != 1) return 1;
  const int e = env_int("PHPC_HOST_BANDS", 0);
  if (e > 0) return e;
  return s->m >= 8192 ? 4 : 1;
}

extern "C" void phpc_summa_run_host(phpc_summa *s, int backend, int ctas, const double *A, const double *B, double *C, int gather,
                                    phpc_summa_stats *stats) {
  phpc_summa_stats local;
  const int bands = host_bands(s);
  if (s->size == 1 && A && B && C) { /* any band count, 1 included: the same arithmetic per element (zeroed block, chunks, + caller's C) */
    summa_run_host_banded(s, backend, ctas, A, B, C, bands, stats ? stats : &local);
    return;
  }
  const double t0 = now_s();
  if (s->size > 1 && s->transport == 1 && gather && A && B && C) {
    /* rank 0's C in node-shared memory: the band-pipelined run, every rank delivers its bands itself */
    ShareInfo sh;
    memset(&sh, 0, sizeof sh);
    if (s->rank == 0) sh.shared = phpc_host_shared_lookup(C, sh.name, &sh.off, &sh.bytes);
    MPI_Bcast(&sh, (int)sizeof sh, MPI_BYTE, 0, s->grid_comm);
    int mb = env_int("PHPC_HOST_BANDS", 0);
    if (mb <= 0) mb = s->m >= 4096 ? 4 : (s->m >= 512 ? 2 : 1);
    if (mb > PHPC_MAX_HOST_BANDS) mb = PHPC_MAX_HOST_BANDS;
    if (sh.shared) {
      summa_run_host_multi_banded(s, backend, ctas, A, B, C, sh, mb, stats ? stats : &local);
      if (getenv("PHPC_DEBUG")) {
        const phpc_summa_stats *st = stats ? stats : &local;
        fprintf(stderr, "[phpc %d] run_host (banded, %d bands): %.1f ms wall (device: loop %.1f ms, GEMMs %.1f ms)\n", s->rank, mb, (now_s() - t0) * 1e3,
                st->total_ms, st->gemm_ms);
      }
      return;
    }
  }
  summa_run(s, backend, ctas, nullptr, stats ? stats : &local, A, B, C, true);
  const double t1 = now_s();
  phpc_summa_download_c(s, C, gather);
  const double t2 = now_s();
  if (s->size > 1 && s->transport == 1) MPI_Barrier(s->grid_comm); /* peers are done pulling before the next upload */
  if (getenv("PHPC_DEBUG")) {
    const phpc_summa_stats *st = stats ? stats : &local;
    fprintf(stderr, "[phpc %d] run_host: uploads+loop %.1f ms (device: loop %.1f ms, GEMMs %.1f ms), download+gather %.1f ms, final barrier %.1f ms\n",
            s->rank, (t1 - t0) * 1e3, st->total_ms, st->gemm_ms, (t2 - t1) * 1e3, (now_s() - t2) * 1e3);
  }
}

/* ------------------------------------------------------------------------- */
/* results                                                                    */
/* ------------------------------------------------------------------------- */
extern "C" void phpc_summa_read_c_block(phpc_summa *s, double *dst, long long ld, int row0, int col0, int rows, int cols) {
  CUDA_CHECK(cudaSetDevice(s->ctx->device));
  CUDA_CHECK(cudaStreamSynchronize(s->ctx->compute));
  CUDA_CHECK(cudaMemcpy2D(dst, (size_t)ld * sizeof(double), s->dC + (size_t)row0 * s->ldn + col0, s->ldn * sizeof(double),
                          (size_t)cols * sizeof(double), rows, cudaMemcpyDeviceToHost));
}

/*
 * The reference gathers the C blocks with one strided MPI_Send per rank into rank 0's
 * host matrix (src/phpc_summa.c:97-110).  Here every block first crosses NVLink into a
 * staging buffer on rank 0's GPU (ncclSend/ncclRecv on the world communicator, two
 * buffers so the next receive overlaps the previous D2H) and reaches rank 0's host C
 * through rank 0's own PCIe link; no host-to-host copy at all.  PHPC_GATHER=mpi keeps
 * the reference's MPI path (and is what runs when the grid has a single rank).
 */
extern "C" void phpc_summa_download_c(phpc_summa *s, double *C, int gather) {
  const size_t N = (size_t)s->N;
  DeviceCtx *ctx = s->ctx;
  CUDA_CHECK(cudaSetDevice(ctx->device));
  const char *mode = getenv("PHPC_GATHER");
  const bool use_mpi = mode && !strcmp(mode, "mpi");
  if (!gather || s->size == 1 || use_mpi) {
    phpc_summa_read_c_block(s, C + (size_t)s->pi * s->m * N + (size_t)s->pj * s->n, (long long)N, 0, 0, s->m, s->n);
    if (!gather || s->size == 1) return;
    MPI_Datatype block_c;
    MPI_Type_vector(s->m, s->n, s->N, MPI_DOUBLE, &block_c);
    MPI_Type_commit(&block_c);
    if (s->rank == 0) {
      for (int i = 1; i < s->size; ++i) {
        int co[2];
        MPI_Cart_coords(s->grid_comm, i, 2, co);
        MPI_Recv(C + N * (size_t)co[0] * s->m + (size_t)co[1] * s->n, 1, block_c, i, 0, s->grid_comm, MPI_STATUS_IGNORE);
      }
    } else {
      MPI_Send(C + N * (size_t)s->pi * s->m + (size_t)s->pj * s->n, 1, block_c, 0, 0, s->grid_comm);
    }
    MPI_Type_free(&block_c);
    return;
  }
  cudaStream_t st = ctx->compute;
  const size_t row_bytes = (size_t)s->n * sizeof(double);
  /* own block -> own place (every rank keeps its block at its global offset, reference :44) */
  CUDA_CHECK(cudaMemcpy2DAsync(C + (size_t)s->pi * s->m * N + (size_t)s->pj * s->n, N * sizeof(double), s->dC, s->ldn * sizeof(double),
                               row_bytes, s->m, cudaMemcpyDeviceToHost, st));
  if (s->transport == 1) {
    /* Parallel gather: when rank 0's C lives in memory every rank can map (phpc_host_malloc_shared), every rank writes its
     * block into it over its OWN PCIe link, all links at once; rank 0's link carries one block instead of all of them. */
    struct {
      int shared;
      char name[64];
      unsigned long long off, bytes;
    } info;
    memset(&info, 0, sizeof info);
    if (s->rank == 0) info.shared = phpc_host_shared_lookup(C, info.name, &info.off, &info.bytes);
    MPI_Bcast(&info, (int)sizeof info, MPI_BYTE, 0, s->grid_comm);
    if (info.shared) {
      if (s->rank != 0) {
        const unsigned long long blk_off = info.off + ((unsigned long long)s->pi * s->m * N + (unsigned long long)s->pj * s->n) * sizeof(double);
        const unsigned long long blk_len = ((unsigned long long)(s->m - 1) * N + s->n) * sizeof(double);
        char *base = (char *)phpc_host_shared_map(info.name, info.bytes, blk_off, blk_len);
        CUDA_CHECK(cudaMemcpy2DAsync(base + blk_off, N * sizeof(double), s->dC, s->ldn * sizeof(double), row_bytes, s->m, cudaMemcpyDeviceToHost,
                                     ctx->copy));
        CUDA_CHECK(cudaStreamSynchronize(ctx->copy));
      }
      CUDA_CHECK(cudaStreamSynchronize(st));
      MPI_Barrier(s->grid_comm); /* every block has landed in rank 0's C */
      return;
    }
  }
  if (s->transport == 1) {
    /* pull transport: the root copies each peer's finished C block out of the peer's HBM */
    CUDA_CHECK(cudaStreamSynchronize(st));
    MPI_Barrier(s->grid_comm); /* every block is final */
  } else if (s->rank != 0) {
    NCCL_CHECK(ncclSend(s->dC, s->c_elems, ncclDouble, 0, g_nccl.world, st));
  }
  if (s->rank == 0) {
    if (!s->gather_stage) CUDA_CHECK(cudaMalloc(&s->gather_stage, 2 * s->c_elems * sizeof(double)));
    cudaEvent_t drained[2];
    for (int b = 0; b < 2; ++b) CUDA_CHECK(cudaEventCreateWithFlags(&drained[b], cudaEventDisableTiming));
    for (int i = 1; i < s->size; ++i) {
      int co[2];
      MPI_Cart_coords(s->grid_comm, i, 2, co);
      const int b = i & 1;
      double *stage = s->gather_stage + (size_t)b * s->c_elems;
      if (i > 2) CUDA_CHECK(cudaStreamWaitEvent(st, drained[b], 0));
      if (s->transport == 1)
        CUDA_CHECK(cudaMemcpyAsync(stage, s->peerC[i], s->c_elems * sizeof(double), cudaMemcpyDeviceToDevice, st));
      else
        NCCL_CHECK(ncclRecv(stage, s->c_elems, ncclDouble, i, g_nccl.world, st));
      CUDA_CHECK(cudaEventRecord(s->ev_user, st));
      CUDA_CHECK(cudaStreamWaitEvent(ctx->copy, s->ev_user, 0));
      CUDA_CHECK(cudaMemcpy2DAsync(C + N * (size_t)co[0] * s->m + (size_t)co[1] * s->n, N * sizeof(double), stage,
                                   s->ldn * sizeof(double), row_bytes, s->m, cudaMemcpyDeviceToHost, ctx->copy));
      CUDA_CHECK(cudaEventRecord(drained[b], ctx->copy));
    }
    CUDA_CHECK(cudaStreamSynchronize(ctx->copy));
    for (int b = 0; b < 2; ++b) CUDA_CHECK(cudaEventDestroy(drained[b]));
  }
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (s->transport == 1) MPI_Barrier(s->grid_comm); /* the root is done reading before anyone touches its block */
}

/* ------------------------------------------------------------------------- */
/* the reference's entry points                                               */
/* ------------------------------------------------------------------------- */
/* The device blocks and communicators of the last host-pointer call are kept: the
 * reference's main.c calls the CUDA pass and the cuBLAS pass back to back on the same
 * grid and size (src/main.c:94,106), and a bench loop calls it repeatedly. */
static phpc_summa *g_host_plan = nullptr;

extern "C" void phpc_summa_release_cache(void) {
  if (g_host_plan) phpc_summa_destroy(g_host_plan);
  g_host_plan = nullptr;
  phpc_host_shared_release_imports(); /* mappings of rank 0's shared result matrix */
}

static void summa_host(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int backend, int ctas, float *seconds) {
  int dims[2], periods[2], coords[2], size;
  MPI_Comm_size(grid_comm, &size);
  MPI_Cart_get(grid_comm, 2, dims, periods, coords);
  phpc_summa *s = g_host_plan;
  if (s && !(s->grid_comm == grid_comm && s->N == n && s->size == size && s->r == dims[0] && s->c == dims[1] && s->pi == coords[0] &&
             s->pj == coords[1])) {
    phpc_summa_release_cache();
    s = nullptr;
  }
  if (!s) {
    /* K chunks of 2048 columns even on one GPU: the uploads pipeline under the GEMMs */
    s = g_host_plan = summa_create(grid_comm, n, n, n, env_int("PHPC_KC", phpc_use_ozaki() ? 4096 : 2048), 1024, 133);
  }
  phpc_summa_stats stats;
  phpc_summa_run_host(s, backend, ctas, A, B, C, 1, &stats);
  if (seconds) *seconds = stats.gemm_ms / 1000.f;
}

extern "C" void phpc_gemm_summa_cuda(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int gpu_count, int grid_width,
                                     int grid_height, int block_width, float *compute_time) {
  (void)gpu_count; /* one rank drives one GPU; see INTEGRATION.md */
  (void)block_width;
  const long long ctas = (long long)grid_width * grid_height;
  summa_host(grid_comm, A, B, C, n, phpc_use_ozaki() ? PHPC_BACKEND_OZAKI : PHPC_BACKEND_DMMA, ctas > (1 << 20) ? (1 << 20) : (int)ctas,
             compute_time);
}

extern "C" void phpc_gemm_summa_cublas(MPI_Comm grid_comm, const double *A, const double *B, double *C, int n, int gpu_count,
                                       float *compute_time) {
  (void)gpu_count;
  summa_host(grid_comm, A, B, C, n, PHPC_BACKEND_CUBLAS, 0, compute_time);
}
