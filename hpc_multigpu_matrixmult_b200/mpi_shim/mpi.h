/*
 * mpi.h — single-node MPI shim for boxes without an MPI installation.
 *
 * The SUMMA hot path uses MPI only as its control plane and process model
 * (reference src/main.c:21-62,90-120 and src/phpc_summa.c:26-34,75-119).  This
 * header + mpi_shim.c implement exactly that surface — the 19 functions and the
 * constants the reference's main.c / phpc_summa.c reference — over one POSIX
 * shared-memory segment, so that (a) `main.out` runs under `bin/mpirun -n P`
 * with the reference's command lines, and (b) the UNMODIFIED reference sources
 * compile against it as the on-box CUDA+MPI baseline (oracle/_ref).  With a
 * real MPI installed, build with MPI=system and this directory is not used.
 *
 * Not a general MPI: blocking calls only, one node, basic + vector datatypes.
 */
#ifndef PHPC_MPI_SHIM_H
#define PHPC_MPI_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHPC_MPI_SHIM 1

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15

#define MPI_COMM_NULL ((MPI_Comm) - 1)
#define MPI_COMM_WORLD ((MPI_Comm)0)

#define MPI_DATATYPE_NULL ((MPI_Datatype)0)
#define MPI_BYTE ((MPI_Datatype)1)
#define MPI_CHAR ((MPI_Datatype)2)
#define MPI_INT ((MPI_Datatype)3)
#define MPI_FLOAT ((MPI_Datatype)4)
#define MPI_DOUBLE ((MPI_Datatype)5)
#define MPI_LONG_LONG ((MPI_Datatype)6)
#define MPI_UNSIGNED_LONG_LONG ((MPI_Datatype)7)

#define MPI_SUM ((MPI_Op)1)
#define MPI_MAX ((MPI_Op)2)
#define MPI_MIN ((MPI_Op)3)

#define MPI_IN_PLACE ((void *)-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_ANY_TAG (-1)

int MPI_Init(int *argc, char ***argv);
int MPI_Initialized(int *flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);

int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_free(MPI_Comm *comm);

int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[], int reorder, MPI_Comm *comm_cart);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[]);
int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[], int coords[]);
int MPI_Cart_sub(MPI_Comm comm, const int remain_dims[], MPI_Comm *newcomm);

int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *datatype);
int MPI_Type_free(MPI_Datatype *datatype);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status *status);

/* ---- shim-only helpers (not MPI) ----------------------------------------- */
/* Create and initialise the shared segment for `nranks` ranks at `path` (a file
 * under /dev/shm or /tmp).  Called by bin/mpirun before forking, or by rank 0 of
 * a torchrun-launched job before the other ranks attach.  Returns 0 on success. */
int phpc_mpi_segment_create(const char *path, int nranks);
int phpc_mpi_segment_unlink(const char *path);
/* MPI_Init for hosts that are not C mains (ctypes): attach to `path` as `rank`
 * of `nranks`; nranks == 1 with path == NULL gives a singleton world. */
int phpc_mpi_init_explicit(const char *path, int rank, int nranks);

#ifdef __cplusplus
}
#endif

#endif
