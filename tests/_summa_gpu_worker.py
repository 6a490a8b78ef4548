"""Worker of tests/test_multigpu.py, one process per GPU under bin/mpirun:
    mpirun -n P python tests/_summa_gpu_worker.py <RxC> <N> <fill> <out.npy> [kc]
Calls the reference-facing C-ABI phpc_gemm_summa_cuda on full host matrices (what the
reference's main.c does, src/main.c:94) and, with kc given, the device-resident k-loop."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hpc_multigpu_matrixmult_b200 import capi  # noqa: E402


def rect(r, c, M, K, N, out, kc):
    """General C[M x N] += A[M x K] * B[K x N] through the object API (phpc_summa_create_mkn): device-generated blocks, then
    blocks uploaded from full host matrices (nonzero C), both with the default (tcgen05) local GEMM; gathered to rank 0."""
    L = capi.load()
    M_ = capi.mpi()
    M_.MPI_Init(None, None)
    rank = int(os.environ.get("PHPC_MPI_RANK", "0"))
    comm = capi.cart_create((r, c))
    s = capi.Summa(comm, N, kc, m=M, k=K)
    assert s.mkn == (M, K, N) and s.block == (M // r, N // c)
    s.fill(capi.FILL_SEEDED)
    s.run()
    Cd = np.zeros((M, N))
    s.download_c(Cd, gather=True)
    A = np.empty((M, K))
    B = np.empty((K, N))
    C0 = np.empty((M, N))
    L.phpc_fill_host(capi._dp(A), K, M, K, 0, 0, K, capi.FILL_SEEDED, 91)
    L.phpc_fill_host(capi._dp(B), N, K, N, 0, 0, N, capi.FILL_SEEDED, 92)
    L.phpc_fill_host(capi._dp(C0), N, M, N, 0, 0, N, capi.FILL_SEEDED, 93)
    Ch = C0.copy()
    s.run_host(A, B, Ch)
    s.destroy()
    if rank == 0:
        np.save(out, np.stack([Cd, Ch]))
    M_.MPI_Finalize()


def main():
    if os.environ.get("PHPC_TEST_RECT"):
        r, c = (int(x) for x in sys.argv[1].split("x"))
        M, K, N = (int(x) for x in os.environ["PHPC_TEST_RECT"].split(","))
        return rect(r, c, M, K, N, sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    r, c = (int(x) for x in sys.argv[1].split("x"))
    N, fill, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    kc = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    L = capi.load()
    M = capi.mpi()
    M.MPI_Init(None, None)
    rank = int(os.environ.get("PHPC_MPI_RANK", "0"))
    comm = capi.cart_create((r, c))
    A = np.empty((N, N))
    B = np.empty((N, N))
    L.phpc_fill_host(capi._dp(A), N, N, N, 0, 0, N, fill, capi.SEED_A)
    L.phpc_fill_host(capi._dp(B), N, N, N, 0, 0, N, fill, capi.SEED_B)
    C = np.zeros((N, N))
    secs = capi.phpc_gemm_summa_cuda(comm, A, B, C)
    assert secs > 0
    Cb = np.zeros((N, N))
    capi.phpc_gemm_summa_cublas(comm, A, B, Cb)
    # the same entry point with rank 0's C in node-shared pinned memory: the band-pipelined multi-rank run, every rank
    # delivers its bands into rank 0's matrix itself (on one rank this is the ordinary path)
    shared = capi.host_array_shared(N, N) if rank == 0 else None
    Cs = shared[0] if shared else np.zeros((N, N))
    Cs[:] = 0.0
    capi.phpc_gemm_summa_cuda(comm, A, B, Cs)
    Cs_copy = Cs.copy()
    capi.phpc_gemm_summa_cuda(comm, A, B, Cs)  # second pass on the same C: C += A*B with a NONZERO caller block (added at the end of each band)
    Cs2 = Cs.copy()
    del Cs
    if shared:
        L.phpc_host_free_shared(shared[1])
    # device-resident loop with its own chunking and on-device generation of the blocks
    s = capi.Summa(comm, N, kc)
    s.fill(fill)
    st = s.run(capi.BACKEND_DMMA)
    Cd = np.zeros((N, N))
    s.download_c(Cd, gather=True)
    # the same loop with the tcgen05 (Ozaki) local GEMM
    s.zero_c()
    M.MPI_Barrier(capi.MPI_COMM_WORLD)
    s.run(capi.BACKEND_OZAKI)
    Co = np.zeros((N, N))
    s.download_c(Co, gather=True)
    s.destroy()
    if rank == 0:
        np.save(out, np.stack([C, Cb, Cd, Co, Cs_copy, Cs2 / 2.0]))
        print(f"steps={st.steps} launches={st.launches} bcasts={st.broadcasts} rx={st.bytes_received} "
              f"total_ms={st.total_ms:.3f} gemm_ms={st.gemm_ms:.3f} exposed_ms={st.exposed_ms:.3f}")
    L.phpc_summa_release_cache()
    M.MPI_Finalize()


if __name__ == "__main__":
    main()
