/*
 * mpi_shim.c — the 19-function single-node MPI subset behind mpi.h, over one
 * shared-memory segment (see mpi.h for why it exists).
 *
 * Segment layout:  header | MAX_SLOTS communicator slots | MAX_RANKS mailboxes
 *   slot     = sense-reversing barrier words + a bounce buffer; every
 *              communicator owns one slot, collectives stream through it in
 *              SLOT_BYTES chunks (pack on the root, barrier, unpack, barrier).
 *   mailbox  = one outgoing single-message buffer per sender for MPI_Send/Recv.
 * Datatypes are basic types or one level of MPI_Type_vector over a basic type;
 * a message is its packed byte stream, so a root may send "1 x vector" while the
 * receivers post "count x MPI_DOUBLE" (reference src/phpc_summa.c:75 vs :78).
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#define SHIM_MAGIC 0x50485043534d5049ull /* "PHPCSMPI" */
#define MAX_RANKS 64
#define MAX_SLOTS 128
#define SLOT_BYTES (1u << 20)
#define MAIL_BYTES (1u << 20)
#define MAX_COMMS 256
#define MAX_TYPES 256

typedef struct {
  _Atomic unsigned count;
  _Atomic unsigned sense;
  char pad[56];
  unsigned char data[SLOT_BYTES];
} slot_t;

typedef struct {
  _Atomic int full; /* 0 = empty, 1 = holds a chunk for `dst` */
  int dst;
  int tag;
  size_t nbytes; /* bytes in this chunk */
  size_t total;  /* bytes of the whole message */
  char pad[32];
  unsigned char data[MAIL_BYTES];
} mail_t;

typedef struct {
  _Atomic uint64_t magic;
  int nranks;
  _Atomic int attached;
  _Atomic int abort_code; /* 0 = running */
  _Atomic int lock;       /* spinlock for the slot allocator */
  int next_slot;
  int free_top;
  int free_stack[MAX_SLOTS];
  char pad[64];
  slot_t slots[MAX_SLOTS];
  mail_t mail[MAX_RANKS];
} seg_t;

typedef struct {
  int used;
  int slot;
  int n;
  int rank;
  unsigned sense; /* local barrier sense */
  int world[MAX_RANKS];
  int ndims; /* 0 = not Cartesian */
  int dims[2];
  int periods[2];
} comm_t;

typedef struct {
  int used;
  int elem; /* bytes of the basic element */
  int base; /* basic datatype handle */
  int count, blocklen, stride; /* vector; basic: 1,1,1 */
} type_t;

static seg_t *g_seg = NULL;
static seg_t g_single_hdr_only; /* never used for data */
static int g_world_rank = 0, g_world_size = 1;
static int g_initialized = 0, g_finalized = 0;
static comm_t g_comms[MAX_COMMS];
static type_t g_types[MAX_TYPES];
static unsigned char *g_single_slot = NULL; /* singleton world: private bounce buffer */

/* ------------------------------------------------------------------------- */
static void shim_fatal(const char *msg) {
  fprintf(stderr, "phpc-mpi[%d]: %s\n", g_world_rank, msg);
  fflush(stderr);
  if (g_seg) atomic_store(&g_seg->abort_code, 1);
  _exit(1);
}

static void check_abort(void) {
  if (g_seg) {
    int code = atomic_load_explicit(&g_seg->abort_code, memory_order_relaxed);
    if (code) _exit(code);
  }
}

static void backoff(unsigned *spins) {
  if (++*spins < 2000) {
    __builtin_ia32_pause();
  } else {
    sched_yield();
    if ((*spins & 0x3ff) == 0) check_abort();
  }
}

static int basic_size(MPI_Datatype t) {
  switch (t) {
    case MPI_BYTE:
    case MPI_CHAR:
      return 1;
    case MPI_INT:
    case MPI_FLOAT:
      return 4;
    case MPI_DOUBLE:
    case MPI_LONG_LONG:
    case MPI_UNSIGNED_LONG_LONG:
      return 8;
    default:
      return 0;
  }
}

static void init_tables(void) {
  memset(g_comms, 0, sizeof g_comms);
  memset(g_types, 0, sizeof g_types);
  for (int t = MPI_BYTE; t <= MPI_UNSIGNED_LONG_LONG; ++t) {
    g_types[t].used = 1;
    g_types[t].elem = basic_size(t);
    g_types[t].base = t;
    g_types[t].count = g_types[t].blocklen = g_types[t].stride = 1;
  }
  comm_t *w = &g_comms[0];
  w->used = 1;
  w->slot = 0;
  w->n = g_world_size;
  w->rank = g_world_rank;
  w->sense = 0;
  for (int i = 0; i < g_world_size; ++i) w->world[i] = i;
}

static comm_t *get_comm(MPI_Comm c) {
  if (!g_initialized) shim_fatal("MPI call before MPI_Init");
  if (c < 0 || c >= MAX_COMMS || !g_comms[c].used) shim_fatal("invalid communicator");
  return &g_comms[c];
}

static type_t *get_type(MPI_Datatype t) {
  if (t <= 0 || t >= MAX_TYPES || !g_types[t].used) shim_fatal("invalid datatype");
  return &g_types[t];
}

static size_t type_bytes(const type_t *t) { return (size_t)t->count * t->blocklen * t->elem; }
static size_t type_extent(const type_t *t) { return ((size_t)(t->count - 1) * t->stride + t->blocklen) * t->elem; }

/* copy `len` bytes of the packed stream of (buf, count x type) starting at stream
 * offset `off`; to_stream != 0 packs into `chunk`, otherwise unpacks from it */
static void stream_copy(void *buf, const type_t *t, size_t off, size_t len, unsigned char *chunk, int to_stream) {
  unsigned char *base = (unsigned char *)buf;
  if (t->count == 1 || t->stride == t->blocklen) { /* contiguous */
    if (to_stream)
      memcpy(chunk, base + off, len);
    else
      memcpy(base + off, chunk, len);
    return;
  }
  const size_t block = (size_t)t->blocklen * t->elem, per_item = type_bytes(t), extent = type_extent(t);
  const size_t stride_b = (size_t)t->stride * t->elem;
  size_t done = 0;
  while (done < len) {
    const size_t p = off + done;
    const size_t item = p / per_item, in_item = p % per_item;
    const size_t b = in_item / block, w = in_item % block;
    size_t run = block - w;
    if (run > len - done) run = len - done;
    unsigned char *addr = base + item * extent + b * stride_b + w;
    if (to_stream)
      memcpy(chunk + done, addr, run);
    else
      memcpy(addr, chunk + done, run);
    done += run;
  }
}

/* ------------------------------------------------------------------------- */
/* segment                                                                    */
/* ------------------------------------------------------------------------- */
int phpc_mpi_segment_create(const char *path, int nranks) {
  if (nranks < 1 || nranks > MAX_RANKS) return -1;
  unlink(path);
  int fd = open(path, O_RDWR | O_CREAT | O_EXCL, 0600);
  if (fd < 0) return -1;
  if (ftruncate(fd, (off_t)sizeof(seg_t)) != 0) {
    close(fd);
    return -1;
  }
  seg_t *s = (seg_t *)mmap(NULL, sizeof(seg_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (s == MAP_FAILED) return -1;
  /* fresh file pages are zero: only the header needs values */
  s->nranks = nranks;
  s->next_slot = 1; /* slot 0 = MPI_COMM_WORLD */
  s->free_top = 0;
  atomic_store(&s->magic, SHIM_MAGIC);
  munmap(s, sizeof(seg_t));
  return 0;
}

int phpc_mpi_segment_unlink(const char *path) { return unlink(path); }

static void attach(const char *path, int rank, int nranks) {
  g_world_rank = rank;
  g_world_size = nranks;
  int fd = -1;
  for (int tries = 0; tries < 20000; ++tries) { /* up to ~20 s for the creator */
    fd = open(path, O_RDWR);
    if (fd >= 0) {
      struct stat st;
      if (fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(seg_t)) break;
      close(fd);
      fd = -1;
    }
    usleep(1000);
  }
  if (fd < 0) shim_fatal("cannot open the shared segment (PHPC_MPI_SHM)");
  g_seg = (seg_t *)mmap(NULL, sizeof(seg_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (g_seg == MAP_FAILED) {
    g_seg = NULL;
    shim_fatal("mmap of the shared segment failed");
  }
  unsigned spins = 0;
  while (atomic_load(&g_seg->magic) != SHIM_MAGIC) backoff(&spins);
  if (g_seg->nranks != nranks) shim_fatal("segment was created for a different number of ranks");
  atomic_fetch_add(&g_seg->attached, 1);
}

int phpc_mpi_init_explicit(const char *path, int rank, int nranks) {
  if (g_initialized) return MPI_SUCCESS;
  if (nranks <= 1 || path == NULL) {
    g_world_rank = 0;
    g_world_size = 1;
    g_seg = NULL;
    g_single_slot = (unsigned char *)malloc(SLOT_BYTES);
  } else {
    if (nranks > MAX_RANKS) shim_fatal("too many ranks for the shim");
    attach(path, rank, nranks);
  }
  init_tables();
  g_initialized = 1;
  (void)g_single_hdr_only;
  return MPI_SUCCESS;
}

int MPI_Init(int *argc, char ***argv) {
  (void)argc;
  (void)argv;
  const char *path = getenv("PHPC_MPI_SHM");
  const char *r = getenv("PHPC_MPI_RANK"), *n = getenv("PHPC_MPI_SIZE");
  if (path && r && n) return phpc_mpi_init_explicit(path, atoi(r), atoi(n));
  return phpc_mpi_init_explicit(NULL, 0, 1); /* singleton, like `./main.out` without mpirun */
}

int MPI_Initialized(int *flag) {
  *flag = g_initialized;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  if (g_initialized && !g_finalized && g_world_size > 1) MPI_Barrier(MPI_COMM_WORLD);
  g_finalized = 1;
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode) {
  (void)comm;
  if (errorcode == 0) errorcode = 1;
  fflush(stdout);
  fflush(stderr);
  if (g_seg) atomic_store(&g_seg->abort_code, errorcode & 0xff ? errorcode & 0xff : 1);
  _exit(errorcode & 0xff ? errorcode & 0xff : 1);
}

double MPI_Wtime(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + tv.tv_usec / 1e6;
}

/* ------------------------------------------------------------------------- */
/* communicators                                                              */
/* ------------------------------------------------------------------------- */
int MPI_Comm_size(MPI_Comm comm, int *size) {
  *size = get_comm(comm)->n;
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  *rank = get_comm(comm)->rank;
  return MPI_SUCCESS;
}

static unsigned char *slot_data(const comm_t *c) { return g_seg ? g_seg->slots[c->slot].data : g_single_slot; }

static void comm_barrier(comm_t *c) {
  if (c->n <= 1) return;
  slot_t *s = &g_seg->slots[c->slot];
  c->sense ^= 1u;
  if (atomic_fetch_add(&s->count, 1u) == (unsigned)c->n - 1u) {
    atomic_store(&s->count, 0u);
    atomic_store(&s->sense, c->sense);
  } else {
    unsigned spins = 0;
    while (atomic_load(&s->sense) != c->sense) backoff(&spins);
  }
}

int MPI_Barrier(MPI_Comm comm) {
  comm_barrier(get_comm(comm));
  return MPI_SUCCESS;
}

static void seg_lock(void) {
  unsigned spins = 0;
  int expected = 0;
  while (!atomic_compare_exchange_weak(&g_seg->lock, &expected, 1)) {
    expected = 0;
    backoff(&spins);
  }
}
static void seg_unlock(void) { atomic_store(&g_seg->lock, 0); }

/* allocate `n` slots (called by one rank), reset their barrier words */
static void alloc_slots(int n, int *out) {
  seg_lock();
  for (int i = 0; i < n; ++i) {
    int s;
    if (g_seg->free_top > 0)
      s = g_seg->free_stack[--g_seg->free_top];
    else if (g_seg->next_slot < MAX_SLOTS)
      s = g_seg->next_slot++;
    else {
      seg_unlock();
      shim_fatal("out of communicator slots");
      return;
    }
    atomic_store(&g_seg->slots[s].count, 0u);
    atomic_store(&g_seg->slots[s].sense, 0u);
    out[i] = s;
  }
  seg_unlock();
}

static void bcast_bytes(comm_t *c, void *buf, size_t bytes, int root); /* fwd */

static MPI_Comm new_comm_handle(void) {
  for (int i = 1; i < MAX_COMMS; ++i)
    if (!g_comms[i].used) return i;
  shim_fatal("out of communicator handles");
  return MPI_COMM_NULL;
}

int MPI_Comm_free(MPI_Comm *comm) {
  if (*comm == MPI_COMM_NULL || *comm == MPI_COMM_WORLD) return MPI_SUCCESS;
  comm_t *c = get_comm(*comm);
  if (c->n > 1) {
    comm_barrier(c); /* nobody is still inside a collective on this slot */
    if (c->rank == 0) {
      seg_lock();
      g_seg->free_stack[g_seg->free_top++] = c->slot;
      seg_unlock();
    }
  }
  c->used = 0;
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

int MPI_Dims_create(int nnodes, int ndims, int dims[]) {
  if (ndims == 1) {
    if (dims[0] == 0) dims[0] = nnodes;
    return MPI_SUCCESS;
  }
  if (ndims != 2) shim_fatal("MPI_Dims_create: only 1 or 2 dimensions");
  if (dims[0] > 0 && dims[1] > 0) return MPI_SUCCESS;
  if (dims[0] > 0) {
    dims[1] = nnodes / dims[0];
    return MPI_SUCCESS;
  }
  if (dims[1] > 0) {
    dims[0] = nnodes / dims[1];
    return MPI_SUCCESS;
  }
  int small = 1; /* most balanced factorisation, larger factor first (8 -> 4 x 2) */
  for (int d = 1; (long)d * d <= nnodes; ++d)
    if (nnodes % d == 0) small = d;
  dims[0] = nnodes / small;
  dims[1] = small;
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm *comm_cart) {
  (void)reorder;
  comm_t *o = get_comm(comm_old);
  if (ndims < 1 || ndims > 2) shim_fatal("MPI_Cart_create: only 1 or 2 dimensions");
  const int d0 = dims[0], d1 = ndims == 2 ? dims[1] : 1;
  const int n = d0 * d1;
  if (n > o->n || n < 1) shim_fatal("MPI_Cart_create: grid larger than the communicator");
  int slot = 0;
  if (o->n > 1) {
    if (o->rank == 0) alloc_slots(1, &slot);
    bcast_bytes(o, &slot, sizeof slot, 0);
  }
  if (o->rank >= n) {
    *comm_cart = MPI_COMM_NULL;
    return MPI_SUCCESS;
  }
  MPI_Comm h = new_comm_handle();
  comm_t *c = &g_comms[h];
  memset(c, 0, sizeof *c);
  c->used = 1;
  c->slot = slot;
  c->n = n;
  c->rank = o->rank;
  for (int i = 0; i < n; ++i) c->world[i] = o->world[i];
  c->ndims = 2;
  c->dims[0] = d0;
  c->dims[1] = d1;
  c->periods[0] = periods ? periods[0] : 0;
  c->periods[1] = (periods && ndims == 2) ? periods[1] : 0;
  *comm_cart = h;
  return MPI_SUCCESS;
}

int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[]) {
  comm_t *c = get_comm(comm);
  if (!c->ndims) shim_fatal("MPI_Cart_coords on a non-Cartesian communicator");
  if (maxdims >= 1) coords[0] = rank / c->dims[1];
  if (maxdims >= 2) coords[1] = rank % c->dims[1];
  return MPI_SUCCESS;
}

int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[], int coords[]) {
  comm_t *c = get_comm(comm);
  if (!c->ndims) shim_fatal("MPI_Cart_get on a non-Cartesian communicator");
  for (int i = 0; i < maxdims && i < 2; ++i) {
    dims[i] = c->dims[i];
    periods[i] = c->periods[i];
  }
  return MPI_Cart_coords(comm, c->rank, maxdims, coords);
}

int MPI_Cart_sub(MPI_Comm comm, const int remain_dims[], MPI_Comm *newcomm) {
  comm_t *o = get_comm(comm);
  if (!o->ndims) shim_fatal("MPI_Cart_sub on a non-Cartesian communicator");
  const int my0 = o->rank / o->dims[1], my1 = o->rank % o->dims[1];
  const int keep0 = remain_dims[0] != 0, keep1 = remain_dims[1] != 0;
  /* colour = coordinates in the dropped dimensions */
  const int ncolors = (keep0 ? 1 : o->dims[0]) * (keep1 ? 1 : o->dims[1]);
  const int color = (keep0 ? 0 : my0) * (keep1 ? 1 : o->dims[1]) + (keep1 ? 0 : my1);
  int slots[MAX_RANKS];
  memset(slots, 0, sizeof slots);
  if (o->n > 1) {
    if (o->rank == 0) alloc_slots(ncolors, slots);
    bcast_bytes(o, slots, sizeof(int) * (size_t)ncolors, 0);
  }
  MPI_Comm h = new_comm_handle();
  comm_t *c = &g_comms[h];
  memset(c, 0, sizeof *c);
  c->used = 1;
  c->slot = slots[color];
  c->ndims = 2;
  c->dims[0] = keep0 ? o->dims[0] : 1;
  c->dims[1] = keep1 ? o->dims[1] : 1;
  c->periods[0] = o->periods[0];
  c->periods[1] = o->periods[1];
  c->n = 0;
  for (int r = 0; r < o->n; ++r) {
    const int c0 = r / o->dims[1], c1 = r % o->dims[1];
    if ((!keep0 && c0 != my0) || (!keep1 && c1 != my1)) continue;
    if (r == o->rank) c->rank = c->n;
    c->world[c->n++] = o->world[r];
  }
  *newcomm = h;
  return MPI_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* datatypes                                                                  */
/* ------------------------------------------------------------------------- */
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype *newtype) {
  type_t *o = get_type(oldtype);
  if (o->count != 1) shim_fatal("MPI_Type_vector: nested derived types are not supported");
  for (int i = MPI_UNSIGNED_LONG_LONG + 1; i < MAX_TYPES; ++i)
    if (!g_types[i].used) {
      g_types[i].used = 1;
      g_types[i].elem = o->elem;
      g_types[i].base = o->base;
      g_types[i].count = count;
      g_types[i].blocklen = blocklength;
      g_types[i].stride = stride;
      *newtype = i;
      return MPI_SUCCESS;
    }
  shim_fatal("out of datatype handles");
  return MPI_ERR_OTHER;
}

int MPI_Type_commit(MPI_Datatype *datatype) {
  get_type(*datatype);
  return MPI_SUCCESS;
}

int MPI_Type_free(MPI_Datatype *datatype) {
  if (*datatype > MPI_UNSIGNED_LONG_LONG) get_type(*datatype)->used = 0;
  *datatype = MPI_DATATYPE_NULL;
  return MPI_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* collectives                                                                */
/* ------------------------------------------------------------------------- */
static void bcast_stream(comm_t *c, void *buf, const type_t *t, size_t total, int root) {
  if (c->n <= 1 || total == 0) return;
  unsigned char *data = slot_data(c);
  for (size_t off = 0; off < total; off += SLOT_BYTES) {
    const size_t len = total - off < SLOT_BYTES ? total - off : SLOT_BYTES;
    if (c->rank == root) stream_copy(buf, t, off, len, data, 1);
    comm_barrier(c);
    if (c->rank != root) stream_copy(buf, t, off, len, data, 0);
    comm_barrier(c);
  }
}

static void bcast_bytes(comm_t *c, void *buf, size_t bytes, int root) { bcast_stream(c, buf, &g_types[MPI_BYTE], bytes, root); }

int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm) {
  comm_t *c = get_comm(comm);
  const type_t *t = get_type(datatype);
  if (root < 0 || root >= c->n) shim_fatal("MPI_Bcast: bad root");
  if (count > 1 && !(t->count == 1 || t->stride == t->blocklen)) {
    /* several strided items: stream them one by one with the right extent */
    type_t one = *t;
    const size_t per = type_bytes(t), ext = type_extent(t);
    for (int i = 0; i < count; ++i) bcast_stream(c, (unsigned char *)buffer + (size_t)i * ext, &one, per, root);
    return MPI_SUCCESS;
  }
  bcast_stream(c, buffer, t, (size_t)count * type_bytes(t), root);
  return MPI_SUCCESS;
}

#define REDUCE_LOOP(T)                                     \
  do {                                                     \
    T *acc = (T *)out;                                     \
    const T *in = (const T *)src;                          \
    for (int i = 0; i < count; ++i) {                      \
      if (op == MPI_SUM)                                   \
        acc[i] = first ? in[i] : (T)(acc[i] + in[i]);      \
      else if (op == MPI_MAX)                              \
        acc[i] = (first || in[i] > acc[i]) ? in[i] : acc[i]; \
      else                                                 \
        acc[i] = (first || in[i] < acc[i]) ? in[i] : acc[i]; \
    }                                                      \
  } while (0)

static void reduce_into(void *out, const void *src, int count, MPI_Datatype dt, MPI_Op op, int first) {
  switch (dt) {
    case MPI_FLOAT:
      REDUCE_LOOP(float);
      break;
    case MPI_DOUBLE:
      REDUCE_LOOP(double);
      break;
    case MPI_INT:
      REDUCE_LOOP(int);
      break;
    case MPI_LONG_LONG:
      REDUCE_LOOP(long long);
      break;
    case MPI_UNSIGNED_LONG_LONG:
      REDUCE_LOOP(unsigned long long);
      break;
    default:
      shim_fatal("MPI_Reduce: unsupported datatype");
  }
}

static void reduce_impl(comm_t *c, const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, int root, int all) {
  const type_t *t = get_type(dt);
  if (t->count != 1) shim_fatal("MPI_Reduce: basic datatypes only");
  if (op != MPI_SUM && op != MPI_MAX && op != MPI_MIN) shim_fatal("MPI_Reduce: unsupported op");
  const size_t bytes = (size_t)count * t->elem;
  const int i_am_dst = all || c->rank == root;
  const void *mine = (sendbuf == MPI_IN_PLACE) ? recvbuf : sendbuf;
  if (c->n <= 1) {
    if (i_am_dst && mine != recvbuf) memcpy(recvbuf, mine, bytes);
    return;
  }
  if (bytes * (size_t)c->n > SLOT_BYTES) shim_fatal("MPI_Reduce: message too large for the shim");
  unsigned char *data = slot_data(c);
  memcpy(data + (size_t)c->rank * bytes, mine, bytes);
  comm_barrier(c);
  if (i_am_dst)
    for (int r = 0; r < c->n; ++r) reduce_into(recvbuf, data + (size_t)r * bytes, count, dt, op, r == 0); /* rank order */
  comm_barrier(c);
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm) {
  reduce_impl(get_comm(comm), sendbuf, recvbuf, count, datatype, op, root, 0);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm) {
  reduce_impl(get_comm(comm), sendbuf, recvbuf, count, datatype, op, 0, 1);
  return MPI_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* point to point                                                             */
/* ------------------------------------------------------------------------- */
int MPI_Send(const void *buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm) {
  comm_t *c = get_comm(comm);
  const type_t *t = get_type(datatype);
  if (dest < 0 || dest >= c->n) shim_fatal("MPI_Send: bad destination");
  if (dest == c->rank || !g_seg) shim_fatal("MPI_Send to self is not supported by the shim");
  if (count != 1 && !(t->count == 1 || t->stride == t->blocklen)) shim_fatal("MPI_Send: count > 1 of a strided type");
  const int wdst = c->world[dest];
  mail_t *m = &g_seg->mail[g_world_rank];
  const size_t total = (size_t)count * type_bytes(t);
  size_t off = 0;
  do { /* a zero-byte message still travels as one empty chunk */
    const size_t len = total - off < MAIL_BYTES ? total - off : MAIL_BYTES;
    unsigned spins = 0;
    while (atomic_load_explicit(&m->full, memory_order_acquire)) backoff(&spins);
    if (len) stream_copy((void *)buf, t, off, len, m->data, 1);
    m->dst = wdst;
    m->tag = tag;
    m->nbytes = len;
    m->total = total;
    atomic_store_explicit(&m->full, 1, memory_order_release);
    off += len;
  } while (off < total);
  return MPI_SUCCESS;
}

int MPI_Recv(void *buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status *status) {
  comm_t *c = get_comm(comm);
  const type_t *t = get_type(datatype);
  if (source < 0 || source >= c->n) shim_fatal("MPI_Recv: bad source (MPI_ANY_SOURCE is not supported)");
  if (!g_seg) shim_fatal("MPI_Recv in a singleton world");
  if (count != 1 && !(t->count == 1 || t->stride == t->blocklen)) shim_fatal("MPI_Recv: count > 1 of a strided type");
  mail_t *m = &g_seg->mail[c->world[source]];
  const size_t capacity = (size_t)count * type_bytes(t);
  size_t total = capacity; /* replaced by the sender's length once the first chunk is seen */
  size_t off = 0;
  do {
    unsigned spins = 0;
    while (!(atomic_load_explicit(&m->full, memory_order_acquire) && m->dst == g_world_rank)) backoff(&spins);
    if (tag != MPI_ANY_TAG && m->tag != tag) shim_fatal("MPI_Recv: tag mismatch (the shim matches messages in order)");
    const size_t len = m->nbytes;
    total = m->total;
    if (total > capacity || off + len > total) shim_fatal("MPI_Recv: message longer than the receive buffer");
    if (len) stream_copy(buf, t, off, len, m->data, 0);
    atomic_store_explicit(&m->full, 0, memory_order_release);
    off += len;
  } while (off < total);
  if (status) {
    status->MPI_SOURCE = source;
    status->MPI_TAG = tag;
    status->MPI_ERROR = MPI_SUCCESS;
  }
  return MPI_SUCCESS;
}
