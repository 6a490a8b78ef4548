"""Worker of tests/test_multiprocess_cpu.py: world_size-N CPU run of the N>1 host logic.
Launched by torch.distributed.run (gloo).  Exercises exactly what bench.py does before it
touches a GPU — gloo rendezvous, shim segment hand-off, MPI_Init, Cartesian grid — and then
walks the product's SUMMA schedule with the shim's MPI_Bcast moving the chunks between the
processes and the ORACLE's block GEMM standing in for the CUDA kernel (test code only)."""
import ctypes
import os
import sys
import time

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hpc_multigpu_matrixmult_b200 import capi  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    r, c = (int(x) for x in sys.argv[1].split("x"))
    N, kc = int(sys.argv[2]), int(sys.argv[3])
    assert r * c == world
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = [None]
    if rank == 0:
        box[0] = f"/dev/shm/phpc_test_{os.getpid()}_{int(time.time())}"
        capi.mpi_segment_create(box[0], world)
    dist.broadcast_object_list(box, src=0)
    capi.mpi_init(rank, world, box[0])
    M = capi.mpi()
    comm = capi.cart_create((r, c))
    pi, pj = rank // c, rank % c
    rows_keep = (ctypes.c_int * 2)(0, 1)
    cols_keep = (ctypes.c_int * 2)(1, 0)
    row_comm, col_comm = ctypes.c_int(), ctypes.c_int()
    M.MPI_Cart_sub(comm, rows_keep, ctypes.byref(row_comm))
    M.MPI_Cart_sub(comm, cols_keep, ctypes.byref(col_comm))

    steps, m, n = capi.summa_schedule(N, r, c, pi, pj, kc)
    A = oracle.fill(N, N, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(N, N, kind=1, seed=oracle.SEED_B)
    blk = np.zeros((m, n))
    for s in steps:
        a = np.ascontiguousarray(A[pi * m:(pi + 1) * m, s.k0:s.k0 + s.width]) if s.own_a else np.empty((m, s.width))
        b = np.ascontiguousarray(B[s.k0:s.k0 + s.width, pj * n:(pj + 1) * n]) if s.own_b else np.empty((s.width, n))
        M.MPI_Bcast(a.ctypes.data, a.size, capi.MPI_DOUBLE, s.a_root, row_comm.value)
        M.MPI_Bcast(b.ctypes.data, b.size, capi.MPI_DOUBLE, s.b_root, col_comm.value)
        blk = oracle.gemm_block(a, b, blk)

    # gather with the strided datatype of reference src/phpc_summa.c:57,97-110
    C = np.zeros((N, N))
    C[pi * m:(pi + 1) * m, pj * n:(pj + 1) * n] = blk
    t = ctypes.c_int()
    M.MPI_Type_vector(m, n, N, capi.MPI_DOUBLE, ctypes.byref(t))
    M.MPI_Type_commit(ctypes.byref(t))
    if rank == 0:
        for i in range(1, world):
            co = (ctypes.c_int * 2)()
            M.MPI_Cart_coords(comm, i, 2, co)
            M.MPI_Recv(C.ctypes.data + (co[0] * m * N + co[1] * n) * 8, 1, t.value, i, 0, comm, None)
    else:
        M.MPI_Send(C.ctypes.data + (pi * m * N + pj * n) * 8, 1, t.value, 0, 0, comm)
    # reduce like src/main.c:97
    val = ctypes.c_float(float(rank + 1))
    out = ctypes.c_float(0)
    M.MPI_Reduce(ctypes.byref(val), ctypes.byref(out), 1, capi.MPI_FLOAT, capi.MPI_SUM, 0, capi.MPI_COMM_WORLD)
    ok = True
    if rank == 0:
        ok = oracle.rel_frobenius(C, oracle.summa(A, B, r, c)) < 1e-15 and out.value == world * (world + 1) / 2
    M.MPI_Barrier(capi.MPI_COMM_WORLD)
    dist.barrier()
    if rank == 0:
        os.unlink(box[0])
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
