"""Panel transports of the device-resident SUMMA loop side by side, same kernel, same chunking:
    bin/mpirun -n P python tools/panel_transports.py <RxC> <N> [steps]
prints one JSON line per rank and transport:
    pull            copy-engine pulls over CUDA IPC (default)
    nccl            one ncclGroup of ncclBroadcast per K chunk on the row / column communicator (PHPC_PANEL=nccl)
    nccl+register   the same with stores and receive rings registered (ncclCommRegister, PHPC_NCCL_REGISTER=1)
`ms` is the device time of one whole k-loop on that rank (the step time is the max over ranks); every transport's C block is
checked element for element against the pull transport's (the transports move the same bits)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hpc_multigpu_matrixmult_b200 import capi  # noqa: E402


def main():
    r, c = (int(x) for x in sys.argv[1].split("x"))
    N = int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    L = capi.load()
    M = capi.mpi()
    M.MPI_Init(None, None)
    rank = int(os.environ.get("PHPC_MPI_RANK", "0"))
    comm = capi.cart_create((r, c))
    want = None
    for name, env in (("pull", {}), ("nccl", {"PHPC_PANEL": "nccl"}), ("nccl+register", {"PHPC_PANEL": "nccl", "PHPC_NCCL_REGISTER": "1"})):
        for k in ("PHPC_PANEL", "PHPC_NCCL_REGISTER"):
            os.environ.pop(k, None)
        os.environ.update(env)
        s = capi.Summa(comm, N, 0)
        s.fill(capi.FILL_SEEDED)
        for _ in range(2):
            s.run(stats=False)
        L.phpc_device_synchronize()
        M.MPI_Barrier(comm)
        s.zero_c()
        ms, exposed = [], []
        for _ in range(steps):
            st = s.run()
            ms.append(st.total_ms)
            exposed.append(st.exposed_ms)
        rows = s.read_c_block(0, 0, 2, s.block[1]) / steps
        if want is None:
            want = rows
        same = bool(np.array_equal(rows, want))
        s.destroy()
        M.MPI_Barrier(comm)
        print(json.dumps({"transport": name, "grid": f"{r}x{c}", "N": N, "rank": rank, "ms": float(np.median(ms)),
                          "exposed_ms": float(np.median(exposed)), "tflops_if_slowest": 2.0 * N ** 3 / (float(np.median(ms)) * 1e-3) / 1e12,
                          "same_bits_as_pull": same}), flush=True)
    M.MPI_Finalize()


if __name__ == "__main__":
    main()
