/*
 * phpc_core.cu — device contexts, the DMMA kernel launcher, the cuBLAS
 * comparison path and the reference-compatible host-pointer entry points
 * phpc_gemm_cuda / phpc_gemm_cublas (reference src/phpc_gemm.cu:59-174).
 */
#include <cuda.h>
#include <string.h>

#include "../../include/phpc_b200.h"
#include "../../include/phpc_gemm.cuh"
#include "dmma_gemm.cuh"
#include "ozaki_gemm.cuh"
#include "ozaki_split.cuh"
#include "phpc_internal.h"

/* ------------------------------------------------------------------------- */
/* errors                                                                     */
/* ------------------------------------------------------------------------- */
[[noreturn]] void phpc_die(const char *what, const char *detail, const char *file, int line) {
  fprintf(stderr, "phpc: %s failed at %s:%d: %s\n", what, file, line, detail ? detail : "");
  fflush(stderr);
  abort();
}

/* ------------------------------------------------------------------------- */
/* contexts                                                                   */
/* ------------------------------------------------------------------------- */
static DeviceCtx g_ctx[PHPC_MAX_DEVICES];
static int g_bound_device = -1;

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode_tiled = nullptr;
static void load_driver_entry_points() {
  if (g_encode_tiled) return;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  PHPC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "driver has no cuTensorMapEncodeTiled (need CUDA 12+ driver)");
  g_encode_tiled = (encode_tiled_fn)fn;
}

DeviceCtx *phpc_ctx(int device) {
  PHPC_REQUIRE(device >= 0 && device < PHPC_MAX_DEVICES, "device index out of range");
  DeviceCtx *ctx = &g_ctx[device];
  CUDA_CHECK(cudaSetDevice(device));
  if (ctx->ready) return ctx;

  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    char msg[160];
    snprintf(msg, sizeof msg, "device %d is sm_%d%d; this library only carries sm_100a (B200) code and has no fallback", device,
             prop.major, prop.minor);
    phpc_die("phpc_ctx", msg, __FILE__, __LINE__);
  }
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  int lo = 0, hi = 0;
  CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->compute, cudaStreamNonBlocking, lo));
  CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->comm, cudaStreamNonBlocking, hi));
  CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->comm2, cudaStreamNonBlocking, hi));
  CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->copy, cudaStreamNonBlocking, lo));
  CUBLAS_CHECK(cublasCreate(&ctx->blas));
  CUDA_CHECK(cudaMalloc(&ctx->sched, 64));
  CUDA_CHECK(cudaMemset(ctx->sched, 0, 64));
  CUDA_CHECK(cudaEventCreate(&ctx->ev0));
  CUDA_CHECK(cudaEventCreate(&ctx->ev1));
  CUDA_CHECK(cudaFuncSetAttribute(phpc::dmma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, phpc::GEMM_SMEM_BYTES));
  CUDA_CHECK(cudaFuncSetAttribute(phpc::oz::ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, phpc::oz::SMEM_BYTES));
  CUDA_CHECK(cudaEventCreateWithFlags(&ctx->gemm_done, cudaEventDisableTiming));
  load_driver_entry_points();
  ctx->ready = true;
  return ctx;
}

DeviceCtx *phpc_cur_ctx(void) {
  if (g_bound_device < 0) g_bound_device = 0;
  return phpc_ctx(g_bound_device);
}

void *phpc_buf_reserve(DevBuf *b, size_t bytes) {
  if (bytes > b->bytes) {
    if (b->ptr) CUDA_CHECK(cudaFree(b->ptr));
    b->ptr = nullptr;
    b->bytes = 0;
    CUDA_CHECK(cudaMalloc(&b->ptr, bytes));
    b->bytes = bytes;
  }
  return b->ptr;
}

extern "C" int phpc_b200_version(void) { return PHPC_B200_VERSION; }

extern "C" int phpc_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

/* Host matrices are first touched by the rank that uploads them: run the rank on the CPUs of ITS GPU's NUMA node, so that its
 * pages and its PCIe link sit on the same socket (8 ranks moving 25.8 + 16 GB per call through one socket's memory controllers
 * and the inter-socket link is what bounds the host-sourced call on an 8-GPU box).  PHPC_NUMA=0 leaves the affinity alone. */
#include <sched.h>
static void bind_to_gpu_numa_node(int device) {
  const char *e = getenv("PHPC_NUMA");
  if (e && atoi(e) == 0) return;
  char bus[32] = {0}, path[128];
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  for (char *c = bus; *c; ++c)
    if (*c >= 'A' && *c <= 'F') *c += 'a' - 'A';
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE *f = fopen(path, "r");
  int node = -1;
  if (f) {
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
  }
  if (node >= 0) {
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    cpu_set_t want, have, both;
    CPU_ZERO(&want);
    if (f) {
      int a, b;
      char sep;
      while (fscanf(f, "%d", &a) == 1) {
        b = a;
        if (fscanf(f, "%c", &sep) == 1 && sep == '-') {
          if (fscanf(f, "%d", &b) != 1) b = a;
          if (fscanf(f, "%c", &sep) != 1) sep = 0;
        }
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET(c, &want);
      }
      fclose(f);
    }
    if (sched_getaffinity(0, sizeof have, &have) == 0) {
      CPU_AND(&both, &want, &have);
      if (CPU_COUNT(&both) > 0) sched_setaffinity(0, sizeof both, &both);
    }
  }
  if (getenv("PHPC_DEBUG")) fprintf(stderr, "[phpc] device %d (%s): NUMA node %d\n", device, bus, node);
}

extern "C" void phpc_b200_set_device(int device) {
  const bool first = g_bound_device != device;
  g_bound_device = device;
  phpc_ctx(device);
  if (first) bind_to_gpu_numa_node(device);
}

extern "C" int phpc_b200_get_device(void) { return phpc_cur_ctx()->device; }
extern "C" int phpc_b200_sm_count(void) { return phpc_cur_ctx()->sm_count; }

void phpc_host_shared_release_imports(void);

extern "C" void phpc_b200_finalize(void) {
  phpc_host_shared_release_imports();
  for (int d = 0; d < PHPC_MAX_DEVICES; ++d) {
    DeviceCtx *ctx = &g_ctx[d];
    if (!ctx->ready) continue;
    cudaSetDevice(d);
    cudaDeviceSynchronize();
    if (ctx->bufA.ptr) cudaFree(ctx->bufA.ptr);
    if (ctx->bufB.ptr) cudaFree(ctx->bufB.ptr);
    if (ctx->bufC.ptr) cudaFree(ctx->bufC.ptr);
    if (ctx->ozA.ptr) cudaFree(ctx->ozA.ptr);
    if (ctx->ozB.ptr) cudaFree(ctx->ozB.ptr);
    if (ctx->ozE.ptr) cudaFree(ctx->ozE.ptr);
    if (ctx->ozSync.ptr) cudaFree(ctx->ozSync.ptr);
    if (ctx->ozG.ptr) cudaFree(ctx->ozG.ptr);
    cudaEventDestroy(ctx->gemm_done);
    if (ctx->ozT.ptr) cudaFree(ctx->ozT.ptr);
    cudaFree(ctx->sched);
    cublasDestroy(ctx->blas);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->compute);
    cudaStreamDestroy(ctx->comm);
    cudaStreamDestroy(ctx->comm2);
    cudaStreamDestroy(ctx->copy);
    *ctx = DeviceCtx();
  }
  if (g_bound_device >= 0) cudaSetDevice(g_bound_device);
}

/* ------------------------------------------------------------------------- */
/* memory helpers                                                             */
/* ------------------------------------------------------------------------- */
extern "C" void *phpc_device_malloc(size_t bytes) {
  phpc_cur_ctx();
  void *p = nullptr;
  CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 16));
  return p;
}
extern "C" void phpc_device_free(void *p) {
  if (p) CUDA_CHECK(cudaFree(p));
}
extern "C" void *phpc_host_malloc_pinned(size_t bytes) {
  phpc_cur_ctx();
  void *p = nullptr;
  CUDA_CHECK(cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable));
  return p;
}
extern "C" void phpc_host_free_pinned(void *p) {
  if (p) CUDA_CHECK(cudaFreeHost(p));
}
/* ---- host memory that every rank of the node can map: lets all ranks write their C block straight into rank 0's result
 * matrix over their own PCIe link (phpc_summa_download_c) instead of funnelling the gather through rank 0's GPU ---- */
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <string>
#include <vector>
struct SharedHost {
  void *base;
  size_t bytes;
  std::string name;
  bool owner;
};
static std::vector<SharedHost> g_shared;

extern "C" void *phpc_host_malloc_shared(size_t bytes) {
  phpc_cur_ctx();
  static int counter = 0;
  char name[64];
  snprintf(name, sizeof name, "/phpc_host_%d_%d", (int)getpid(), counter++);
  const size_t len = (bytes + 4095) / 4096 * 4096;
  int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return nullptr;
  /* every GPU of the node writes into this allocation: spread its pages over all NUMA nodes (MPOL_INTERLEAVE while the pages
   * are allocated) instead of piling them onto the node of the allocating rank */
  unsigned long all_nodes[16];
  memset(all_nodes, 0xff, sizeof all_nodes);
  const bool interleaved = syscall(SYS_set_mempolicy, 3 /* MPOL_INTERLEAVE */, all_nodes, 64ul) == 0;
  const int rc = posix_fallocate(fd, 0, (off_t)len); /* reserve now: a full /dev/shm must not turn into SIGBUS later */
  if (interleaved) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
  if (rc != 0) {
    close(fd);
    shm_unlink(name);
    return nullptr;
  }
  void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) {
    shm_unlink(name);
    return nullptr;
  }
  CUDA_CHECK(cudaHostRegister(p, len, cudaHostRegisterPortable));
  g_shared.push_back({p, len, name, true});
  return p;
}

extern "C" void phpc_host_free_shared(void *p) {
  for (size_t i = 0; i < g_shared.size(); ++i)
    if (g_shared[i].base == p) {
      cudaHostUnregister(p);
      munmap(p, g_shared[i].bytes);
      if (g_shared[i].owner) shm_unlink(g_shared[i].name.c_str());
      g_shared.erase(g_shared.begin() + i);
      return;
    }
}

/* is [p, p + 1) inside a shared allocation OWNED by this process?  name (>= 64 bytes), offset of p and size of the allocation */
int phpc_host_shared_lookup(const void *p, char *name, unsigned long long *offset, unsigned long long *bytes) {
  for (const SharedHost &h : g_shared)
    if (h.owner && (const char *)p >= (const char *)h.base && (const char *)p < (const char *)h.base + h.bytes) {
      snprintf(name, 64, "%s", h.name.c_str());
      *offset = (unsigned long long)((const char *)p - (const char *)h.base);
      *bytes = h.bytes;
      return 1;
    }
  return 0;
}

/* importer side: map another rank's shared allocation (cached by name); [off, off + len) of it is page-locked for this
 * process' GPU on first use */
struct PinnedRange {
  std::string name;
  size_t lo, hi;
};
static std::vector<PinnedRange> g_pinned;

/* importer side: drop every mapping of other ranks' shared allocations (their memory is only returned to the system once the
 * last mapping is gone); called when the cached SUMMA object is released and at finalize */
void phpc_host_shared_release_imports(void) {
  for (size_t i = 0; i < g_shared.size();) {
    if (g_shared[i].owner) {
      ++i;
      continue;
    }
    for (size_t j = 0; j < g_pinned.size();) {
      if (g_pinned[j].name == g_shared[i].name) {
        cudaHostUnregister((char *)g_shared[i].base + g_pinned[j].lo);
        g_pinned.erase(g_pinned.begin() + j);
      } else {
        ++j;
      }
    }
    munmap(g_shared[i].base, g_shared[i].bytes);
    g_shared.erase(g_shared.begin() + i);
  }
}

void *phpc_host_shared_map(const char *name, unsigned long long bytes, unsigned long long off, unsigned long long len) {
  std::vector<PinnedRange> &pinned = g_pinned;
  void *base = nullptr;
  for (const SharedHost &h : g_shared)
    if (h.name == name) base = h.base;
  if (!base) {
    int fd = shm_open(name, O_RDWR, 0600);
    PHPC_REQUIRE(fd >= 0, "cannot open the shared host allocation of the gather root");
    base = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    PHPC_REQUIRE(base != MAP_FAILED, "cannot map the shared host allocation of the gather root");
    g_shared.push_back({base, (size_t)bytes, name, false});
  }
  size_t lo = off / 4096 * 4096, hi = (off + len + 4095) / 4096 * 4096 < bytes ? (off + len + 4095) / 4096 * 4096 : bytes;
  bool have = false;
  for (const PinnedRange &q : pinned)
    if (q.name == name && q.lo <= lo && q.hi >= hi) have = true;
  if (!have) {
    phpc_cur_ctx();
    /* a range may only be registered once: merge with whatever of this allocation is pinned already and overlaps or touches */
    for (size_t i = 0; i < pinned.size();) {
      if (pinned[i].name == name && pinned[i].lo <= hi && pinned[i].hi >= lo) {
        CUDA_CHECK(cudaHostUnregister((char *)base + pinned[i].lo));
        lo = pinned[i].lo < lo ? pinned[i].lo : lo;
        hi = pinned[i].hi > hi ? pinned[i].hi : hi;
        pinned.erase(pinned.begin() + i);
      } else {
        ++i;
      }
    }
    CUDA_CHECK(cudaHostRegister((char *)base + lo, hi - lo, cudaHostRegisterPortable));
    pinned.push_back({name, lo, hi});
  }
  return base;
}

extern "C" void phpc_host_register(void *p, size_t bytes) {
  phpc_cur_ctx();
  CUDA_CHECK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
}
extern "C" void phpc_host_unregister(void *p) { CUDA_CHECK(cudaHostUnregister(p)); }
extern "C" void phpc_device_memset(void *p, int value, size_t bytes) {
  DeviceCtx *ctx = phpc_cur_ctx();
  CUDA_CHECK(cudaMemsetAsync(p, value, bytes, ctx->compute));
  CUDA_CHECK(cudaStreamSynchronize(ctx->compute));
}
extern "C" void phpc_copy2d_to_host(double *host, long long ld_host, const double *dev, long long ld_dev, long long rows,
                                    long long cols) {
  phpc_cur_ctx();
  if (rows <= 0 || cols <= 0) return;
  CUDA_CHECK(cudaMemcpy2D(host, (size_t)ld_host * 8, dev, (size_t)ld_dev * 8, (size_t)cols * 8, (size_t)rows, cudaMemcpyDeviceToHost));
}
extern "C" void phpc_copy2d_to_device(double *dev, long long ld_dev, const double *host, long long ld_host, long long rows,
                                      long long cols) {
  phpc_cur_ctx();
  if (rows <= 0 || cols <= 0) return;
  CUDA_CHECK(cudaMemcpy2D(dev, (size_t)ld_dev * 8, host, (size_t)ld_host * 8, (size_t)cols * 8, (size_t)rows, cudaMemcpyHostToDevice));
}
extern "C" void phpc_device_synchronize(void) {
  phpc_cur_ctx();
  CUDA_CHECK(cudaDeviceSynchronize());
}

/* ------------------------------------------------------------------------- */
/* DMMA kernel launcher                                                       */
/* ------------------------------------------------------------------------- */
static void encode_map_2d(CUtensorMap *map, const double *base, long long inner, long long outer, long long ld, int box_inner,
                          int box_outer) {
  PHPC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand base must be 16-byte aligned");
  PHPC_REQUIRE((ld & 1) == 0, "TMA operand leading dimension must be even (16-byte global stride)");
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "CUresult %d (inner=%lld outer=%lld ld=%lld box=%dx%d)", (int)r, inner, outer, ld, box_inner, box_outer);
    phpc_die("cuTensorMapEncodeTiled", msg, __FILE__, __LINE__);
  }
}

/* Every GEMM launch of a device is ordered after the previous one, whatever stream it was given: the tile-scheduler words,
 * the digit stores / exponents / guard of the tcgen05 path and its wave counters are per-device scratch, and two persistent
 * grids that each wait for all of their CTAs to be resident must never share the SMs. */
static void gemm_order_begin(DeviceCtx *ctx, cudaStream_t stream) {
  if (ctx->gemm_in_flight && ctx->gemm_last_stream != stream) CUDA_CHECK(cudaStreamWaitEvent(stream, ctx->gemm_done, 0));
}
static void gemm_order_end(DeviceCtx *ctx, cudaStream_t stream) {
  CUDA_CHECK(cudaEventRecord(ctx->gemm_done, stream));
  ctx->gemm_in_flight = true;
  ctx->gemm_last_stream = stream;
}

static int launch_dmma_guarded(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                               int k, int n, int ctas, cudaStream_t stream, const int *guard) {
  if (m <= 0 || n <= 0 || k <= 0) return 0; /* C += 0 */
  PHPC_REQUIRE(lda >= k && ldb >= n && ldc >= n, "leading dimension smaller than the row length");
  CUtensorMap tmA, tmB;
  encode_map_2d(&tmA, dA, k, m, lda, phpc::BK, phpc::BM);         /* A: inner = k, outer = m */
  encode_map_2d(&tmB, dB, n, k, ldb, phpc::B_BOX_COLS, phpc::BK); /* B: inner = n, outer = k */

  phpc::GemmParams p;
  p.C = dC;
  p.ldc = ldc;
  p.M = m;
  p.N = n;
  p.K = k;
  p.tiles_m = (m + phpc::BM - 1) / phpc::BM;
  p.tiles_n = (n + phpc::BN - 1) / phpc::BN;
  p.k_iters = (k + phpc::BK - 1) / phpc::BK;
  p.sched = ctx->sched;
  p.guard = guard;
  const long long tiles = (long long)p.tiles_m * p.tiles_n;
  PHPC_REQUIRE(tiles < (1ll << 30), "too many output tiles for the 32-bit tile counter");
  int grid = (ctas <= 1) ? ctx->sm_count : (ctas < ctx->sm_count ? ctas : ctx->sm_count);
  if ((long long)grid > tiles) grid = (int)tiles;
  phpc::dmma_gemm_kernel<<<grid, phpc::GEMM_THREADS, phpc::GEMM_SMEM_BYTES, stream>>>(tmA, tmB, p);
  CUDA_CHECK(cudaGetLastError());
  return 1;
}

int phpc_launch_dmma(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                     int k, int n, int ctas, cudaStream_t stream) {
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  gemm_order_begin(ctx, stream);
  /* K chunks of 4096: the persistent CTAs of one launch drift apart over thousands of k iterations and the panel re-reads
   * then miss L2 (1.5 TB of DRAM traffic for one N = 32768 launch, profiles/ncu_dmma_n32768_dram_r01.csv); between chunks
   * the grid re-synchronises for free.  The per-element sum still runs in ascending k. */
  int launches = 0;
  const int kc_max = 4096;
  for (int k0 = 0; k0 < k; k0 += kc_max) {
    const int kc = (k - k0 < kc_max) ? k - k0 : kc_max;
    launches += launch_dmma_guarded(ctx, dA + k0, lda, dB + (long long)k0 * ldb, ldb, dC, ldc, m, kc, n, ctas, stream, nullptr);
  }
  gemm_order_end(ctx, stream);
  return launches;
}

/* ------------------------------------------------------------------------- */
/* tcgen05 (Ozaki, int8) launcher                                             */
/* ------------------------------------------------------------------------- */
extern "C" void phpc_ozaki_config(int *digits, int *products, int *k_chunk, int *max_spread) {
  if (digits) *digits = phpc::oz::S;
  if (products) *products = phpc::oz::PRODUCTS;
  if (k_chunk) *k_chunk = phpc::oz::KC_MAX;
  if (max_spread) *max_spread = phpc::oz::MAX_SPREAD;
}

static long long g_oz_tstamp_tiles = 0;
/* diagnostics (PHPC_OZ_TSTAMP=1): per tile of the LAST launch, globaltimer ns at the start of its loads and at the end of its epilogue */
extern "C" long long phpc_oz_tstamp_read(unsigned long long *out, long long max_tiles) {
  DeviceCtx *ctx = phpc_cur_ctx();
  CUDA_CHECK(cudaDeviceSynchronize());
  const long long n = g_oz_tstamp_tiles < max_tiles ? g_oz_tstamp_tiles : max_tiles;
  if (n > 0 && ctx->ozT.ptr) CUDA_CHECK(cudaMemcpy(out, ctx->ozT.ptr, (size_t)n * 16, cudaMemcpyDeviceToHost));
  return n;
}
/* How many K chunks of the tcgen05 path the guard handed to the native-FP64 kernel since the last call (synchronises the device). */
extern "C" long long phpc_ozaki_fallback_chunks(void) {
  DeviceCtx *ctx = phpc_cur_ctx();
  CUDA_CHECK(cudaDeviceSynchronize());
  if (!ctx->ozG.ptr) return 0;
  int words[2] = {0, 0};
  CUDA_CHECK(cudaMemcpy(words, (int *)ctx->ozG.ptr + 1, sizeof words, cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemset((int *)ctx->ozG.ptr + 1, 0, sizeof words));
  return words[0];
}
__global__ void guard_count_kernel(int *g) { /* g[0] = guard of the chunk, g[1] = chunks sent to the native kernel, g[2] = OR of their reasons */
  if (g[0] != 0) {
    g[1] += 1;
    g[2] |= g[0];
  }
}

size_t phpc_ozaki_bcache_bytes(int k, int n, size_t *exp_ints) {
  using namespace phpc::oz;
  const size_t n_pad = (size_t)((n + BN - 1) / BN) * BN, kp = (size_t)((k + 127) / 128) * 128;
  if (exp_ints) *exp_ints = 2 * (size_t)n;
  return (size_t)S * n_pad * kp;
}

int phpc_launch_ozaki(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                      int k, int n, int ctas, cudaStream_t stream, OzBCache *bcache) {
  using namespace phpc::oz;
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  if (bcache && k > KC_MAX) bcache = nullptr; /* the cache describes ONE K chunk */
  PHPC_REQUIRE(lda >= k && ldb >= n && ldc >= n, "leading dimension smaller than the row length");
  gemm_order_begin(ctx, stream);
  int launches = 0;
  const char *fl = getenv("PHPC_OZ_FLAGS"), *ts = getenv("PHPC_OZ_TSTAMP"); /* diagnostics only (tools/ozaki_knobs.py) */
  const int flags = (fl && *fl) ? atoi(fl) : 0;
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  const long long tiles = (long long)tiles_m * tiles_n;
  PHPC_REQUIRE(tiles < (1ll << 30), "too many output tiles");
  const size_t m_pad = (size_t)tiles_m * BM, n_pad = (size_t)tiles_n * BN;
  /* grid_width x grid_height of the reference's CLI = number of persistent CTAs (<= 1: one per SM), as for the DMMA kernel */
  int grid = (ctas <= 1) ? ctx->sm_count : (ctas < ctx->sm_count ? ctas : ctx->sm_count);
  if ((long long)grid > tiles) grid = (int)tiles;
  const size_t waves = (size_t)((tiles + grid - 1) / grid);
  if (!ctx->ozG.ptr) {
    phpc_buf_reserve(&ctx->ozG, 64);
    CUDA_CHECK(cudaMemsetAsync(ctx->ozG.ptr, 0, 64, stream));
  }
  int *guard = (int *)ctx->ozG.ptr;
  for (int k0 = 0; k0 < k; k0 += KC_MAX) {
    const int kc = (k - k0 < KC_MAX) ? k - k0 : KC_MAX;
    const int kp = (kc + 127) / 128 * 128;
    int8_t *TA = (int8_t *)phpc_buf_reserve(&ctx->ozA, (size_t)S * m_pad * kp);
    int8_t *TB = bcache ? bcache->TB : (int8_t *)phpc_buf_reserve(&ctx->ozB, (size_t)S * n_pad * kp);
    int *eA = (int *)phpc_buf_reserve(&ctx->ozE, 2 * ((size_t)m + n) * sizeof(int)); /* A: maxima then minima; B likewise behind them */
    int *eminA = eA + m;
    int *eB = bcache ? bcache->eB : eA + 2 * (size_t)m, *eminB = eB + n;
    const bool b_ready = bcache && bcache->ready; /* exponents and digits of this B chunk were computed by an earlier call */
    unsigned int *wave_sync = (unsigned int *)phpc_buf_reserve(&ctx->ozSync, (waves + 1) * sizeof(unsigned int));
    const double *a = dA + k0;
    const double *b = dB + (long long)k0 * ldb;
    CUDA_CHECK(cudaMemsetAsync(wave_sync, 0, (waves + 1) * sizeof(unsigned int), stream));
    exp_init_kernel<<<(2 * m + 255) / 256, 256, 0, stream>>>(eA, m, guard);
    if (!b_ready) exp_init_kernel<<<(2 * n + 255) / 256, 256, 0, stream>>>(eB, n, guard);
    {
      const int segs = (kc + 1023) / 1024;
      const long long units = (long long)m * segs;
      row_exp_kernel<<<(unsigned)((units + 7) / 8), 256, 0, stream>>>(a, lda, m, kc, eA, eminA, guard);
      dim3 cgrid((n + 255) / 256, (kc + 63) / 64);
      if (!b_ready) col_exp_kernel<<<cgrid, 256, 0, stream>>>(b, ldb, kc, n, eB, eminB, guard);
      guard_kernel<<<1, 1024, 0, stream>>>(eA, eminA, m, eB, eminB, n, guard);
      guard_count_kernel<<<1, 1, 0, stream>>>(guard);
    }
    {
      const long long threads = (long long)m_pad * (kp / 16);
      dim3 bgrid((unsigned)((n_pad + 127) / 128), kp / 32);
      split_a_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(a, lda, m, (int)m_pad, kc, kp, eA, TA, guard);
      /* a cached B chunk is split whatever THIS call's guard says: a later call with other A rows may take the tcgen05 path */
      if (!b_ready) split_b_kernel<<<bgrid, 128, 0, stream>>>(b, ldb, kc, n, (int)n_pad, kp, eB, TB, bcache ? nullptr : guard);
      if (bcache) bcache->ready = true;
    }
    Params p;
    p.C = dC;
    p.ldc = ldc;
    p.M = m;
    p.N = n;
    p.ksteps = kp / BKB;
    p.eA = eA;
    p.eB = eB;
    p.tiles_m = tiles_m;
    p.tiles_n = tiles_n;
    p.TA = TA;
    p.TB = TB;
    p.guard = guard;
    p.wave_sync = wave_sync;
    p.flags = flags;
    p.tstamp = nullptr;
    if (ts && atoi(ts)) {
      p.tstamp = (unsigned long long *)phpc_buf_reserve(&ctx->ozT, (size_t)tiles * 16);
      g_oz_tstamp_tiles = (long long)tiles;
    }
    ozaki_gemm_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(p);
    CUDA_CHECK(cudaGetLastError());
    /* the same chunk on the native-FP64 kernel, which only runs when the guard is set (and the kernel above returned at once) */
    launches += (b_ready ? 6 : 9) + launch_dmma_guarded(ctx, a, lda, b, ldb, dC, ldc, m, kc, n, ctas, stream, guard);
  }
  gemm_order_end(ctx, stream);
  return launches;
}

void phpc_launch_cublas(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                        int k, int n, cudaStream_t stream) {
  if (m <= 0 || n <= 0 || k <= 0) return;
  const double one = 1.0;
  CUBLAS_CHECK(cublasSetStream(ctx->blas, stream));
  /* row-major C = A*B  <=>  column-major C^T = B^T * A^T (reference src/phpc_gemm.cu:169-170) */
  CUBLAS_CHECK(cublasDgemm(ctx->blas, CUBLAS_OP_N, CUBLAS_OP_N, n, m, k, &one, dB, (int)ldb, dA, (int)lda, &one, dC, (int)ldc));
}

extern "C" int phpc_gemm_device(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m, int k,
                                int n, int ctas, void *stream) {
  DeviceCtx *ctx = phpc_cur_ctx();
  return phpc_launch_dmma(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, ctas, stream ? (cudaStream_t)stream : ctx->compute);
}

extern "C" void phpc_gemm_device_cublas(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                                        int k, int n, void *stream) {
  DeviceCtx *ctx = phpc_cur_ctx();
  phpc_launch_cublas(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, stream ? (cudaStream_t)stream : ctx->compute);
}

extern "C" int phpc_gemm_device_ozaki(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                                      int k, int n, void *stream) {
  DeviceCtx *ctx = phpc_cur_ctx();
  return phpc_launch_ozaki(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, 0, stream ? (cudaStream_t)stream : ctx->compute);
}

extern "C" float phpc_gemm_device_timed(const double *dA, long long lda, const double *dB, long long ldb, double *dC, long long ldc, int m,
                                        int k, int n, int ctas, int reps, int use_cublas) {
  DeviceCtx *ctx = phpc_cur_ctx();
  if (reps < 1) reps = 1;
  CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->compute));
  for (int r = 0; r < reps; ++r) {
    if (use_cublas == 1)
      phpc_launch_cublas(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, ctx->compute);
    else if (use_cublas == 2)
      phpc_launch_ozaki(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, ctas, ctx->compute);
    else
      phpc_launch_dmma(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, ctas, ctx->compute);
  }
  CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->compute));
  CUDA_CHECK(cudaEventSynchronize(ctx->ev1));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  return ms / reps;
}

/* ------------------------------------------------------------------------- */
/* synthetic fills                                                            */
/* ------------------------------------------------------------------------- */
__host__ __device__ static inline double phpc_seeded_value(unsigned long long seed, unsigned long long flat) {
  /* splitmix64 of (seed, flat index) -> 53 random bits -> uniform in (-1, 1) */
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (flat + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

__global__ void fill_kernel(double *d, long long ld, long long rows, long long cols, long long row0, long long col0, long long N, int kind,
                            unsigned long long seed) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    const unsigned long long flat = (unsigned long long)((row0 + r) * N + (col0 + c));
    d[r * ld + c] = (kind == PHPC_FILL_INDEX) ? (double)flat : phpc_seeded_value(seed, flat);
  }
}

extern "C" void phpc_fill_device(double *d, long long ld, long long rows, long long cols, long long row0, long long col0, long long N,
                                 int kind, unsigned long long seed, void *stream) {
  DeviceCtx *ctx = phpc_cur_ctx();
  if (rows <= 0 || cols <= 0) return;
  fill_kernel<<<ctx->sm_count * 8, 256, 0, stream ? (cudaStream_t)stream : ctx->compute>>>(d, ld, rows, cols, row0, col0, N, kind, seed);
  CUDA_CHECK(cudaGetLastError());
}

extern "C" void phpc_fill_host(double *h, long long ld, long long rows, long long cols, long long row0, long long col0, long long N,
                               int kind, unsigned long long seed) {
  for (long long r = 0; r < rows; ++r)
    for (long long c = 0; c < cols; ++c) {
      const unsigned long long flat = (unsigned long long)((row0 + r) * N + (col0 + c));
      h[r * ld + c] = (kind == PHPC_FILL_INDEX) ? (double)flat : phpc_seeded_value(seed, flat);
    }
}

/* ------------------------------------------------------------------------- */
/* host-pointer entry points (reference src/phpc_gemm.cu:59-174)              */
/* ------------------------------------------------------------------------- */
typedef void (*launch_fn)(DeviceCtx *, const double *, long long, const double *, long long, double *, long long, int, int, int, int,
                          cudaStream_t);

/* Which kernel the reference-named entry points (phpc_gemm_cuda, phpc_gemm_summa_cuda) run.
 * Default: the tcgen05/TMEM kernel (FP64 rebuilt from int8 MMAs, 8 digits); PHPC_GEMM=dmma selects
 * the native-FP64 DMMA kernel.  Both are sm_100a code; there is no other path. */
bool phpc_use_ozaki(void) {
  const char *g = getenv("PHPC_GEMM");
  return !(g && !strcmp(g, "dmma"));
}
/* the same choice for C callers of the object API (main.out in device mode): PHPC_BACKEND_OZAKI (2) or PHPC_BACKEND_DMMA (0) */
extern "C" int phpc_default_backend(void) { return phpc_use_ozaki() ? 2 : 0; }

static void launch_dmma_adapter(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC,
                                long long ldc, int m, int k, int n, int ctas, cudaStream_t s) {
  if (phpc_use_ozaki())
    phpc_launch_ozaki(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, ctas, s);
  else
    phpc_launch_dmma(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, ctas, s);
}
static void launch_cublas_adapter(DeviceCtx *ctx, const double *dA, long long lda, const double *dB, long long ldb, double *dC,
                                  long long ldc, int m, int k, int n, int, cudaStream_t s) {
  phpc_launch_cublas(ctx, dA, lda, dB, ldb, dC, ldc, m, k, n, s);
}

/*
 * Column split over the local GPUs exactly as reference :97-129 (A replicated,
 * B and C sliced by columns, dev_n = n/g + (gpu < n%g)); device buffers are
 * cached per device and padded to a 128-byte leading dimension for TMA, the
 * host ranges are never pinned (the reference's cudaHostRegister of m*lda
 * elements from an interior pointer overruns the allocation).
 */
static float host_gemm(launch_fn launch, const double *a, int lda, const double *b, int ldb, double *c, int ldc, int m, int k, int n,
                       int gpu_count, int ctas) {
  if (m <= 0 || n <= 0) return 0.f;
  const int visible = phpc_b200_device_count();
  PHPC_REQUIRE(visible > 0, "no CUDA device visible (this library has no CPU fallback)");
  if (gpu_count < 1) gpu_count = 1;
  if (gpu_count > visible) gpu_count = visible;
  if (gpu_count > n) gpu_count = n;
  const int first = (gpu_count == 1 && g_bound_device >= 0) ? g_bound_device : 0;

  int col = 0;
  for (int gi = 0; gi < gpu_count; ++gi) {
    const int dev_n = n / gpu_count + (gi < n % gpu_count);
    DeviceCtx *ctx = phpc_ctx(first + gi);
    const long long pa = phpc_pad_ld(k), pb = phpc_pad_ld(dev_n), pc = phpc_pad_ld(dev_n);
    double *dA = (double *)phpc_buf_reserve(&ctx->bufA, (size_t)m * pa * sizeof(double));
    double *dB = (double *)phpc_buf_reserve(&ctx->bufB, (size_t)(k > 0 ? k : 1) * pb * sizeof(double));
    double *dC = (double *)phpc_buf_reserve(&ctx->bufC, (size_t)m * pc * sizeof(double));
    cudaStream_t s = ctx->compute;
    if (k > 0) {
      CUDA_CHECK(cudaMemcpy2DAsync(dA, pa * sizeof(double), a, (size_t)lda * sizeof(double), (size_t)k * sizeof(double), m,
                                   cudaMemcpyHostToDevice, s));
      CUDA_CHECK(cudaMemcpy2DAsync(dB, pb * sizeof(double), b + col, (size_t)ldb * sizeof(double), (size_t)dev_n * sizeof(double), k,
                                   cudaMemcpyHostToDevice, s));
    }
    CUDA_CHECK(cudaMemcpy2DAsync(dC, pc * sizeof(double), c + col, (size_t)ldc * sizeof(double), (size_t)dev_n * sizeof(double), m,
                                 cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaEventRecord(ctx->ev0, s));
    launch(ctx, dA, pa, dB, pb, dC, pc, m, k, dev_n, ctas, s);
    CUDA_CHECK(cudaEventRecord(ctx->ev1, s));
    col += dev_n;
  }
  /* downloads in a second loop: a D2H into pageable memory returns only when it has finished, which would keep GPU g+1 from
   * even starting its upload until GPU g is completely done (the kernels of all GPUs are in flight by now) */
  col = 0;
  for (int gi = 0; gi < gpu_count; ++gi) {
    const int dev_n = n / gpu_count + (gi < n % gpu_count);
    DeviceCtx *ctx = phpc_ctx(first + gi);
    const long long pc = phpc_pad_ld(dev_n);
    CUDA_CHECK(cudaMemcpy2DAsync(c + col, (size_t)ldc * sizeof(double), ctx->bufC.ptr, pc * sizeof(double), (size_t)dev_n * sizeof(double), m,
                                 cudaMemcpyDeviceToHost, ctx->compute));
    col += dev_n;
  }
  float total_ms = 0.f;
  for (int gi = 0; gi < gpu_count; ++gi) {
    DeviceCtx *ctx = phpc_ctx(first + gi);
    CUDA_CHECK(cudaStreamSynchronize(ctx->compute));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    total_ms += ms;
  }
  if (g_bound_device >= 0) CUDA_CHECK(cudaSetDevice(g_bound_device));
  return total_ms / (gpu_count * 1000.f);
}

extern "C" void phpc_gemm_cuda(const double *a, int lda, const double *b, int ldb, double *c, int ldc, int m, int k, int n, int gpu_count,
                               int grid_width, int grid_height, int block_width, float *compute_time) {
  (void)block_width;
  const long long ctas = (long long)grid_width * grid_height;
  const float secs = host_gemm(launch_dmma_adapter, a, lda, b, ldb, c, ldc, m, k, n, gpu_count, ctas > 1 << 20 ? 1 << 20 : (int)ctas);
  if (compute_time) *compute_time = secs;
}

extern "C" void phpc_gemm_cublas(const double *a, int lda, const double *b, int ldb, double *c, int ldc, int m, int k, int n,
                                 int gpu_count, int grid_width, int grid_height, int block_width, float *gpu_time) {
  (void)grid_width;
  (void)grid_height;
  (void)block_width;
  host_gemm(launch_cublas_adapter, a, lda, b, ldb, c, ldc, m, k, n, gpu_count, 0);
  if (gpu_time) *gpu_time = 0; /* reference src/phpc_gemm.cu:173 */
}
