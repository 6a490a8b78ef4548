"""one_gemm.py — one device-resident local GEMM of the given size, for profiler captures:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum ... python tools/one_gemm.py 32768 dmma
backend: dmma | tcgen05 | cublas.  Prints the event-timed TFLOP/s of a second, unprofiled-quality run."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hpc_multigpu_matrixmult_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1])
    backend = {"dmma": 0, "cublas": 1, "tcgen05": 2}[sys.argv[2]]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    L = capi.load()
    L.phpc_b200_set_device(0)
    dA, dB, dC = (L.phpc_device_malloc(n * n * 8) for _ in range(3))
    L.phpc_fill_device(dA, n, n, n, 0, 0, n, 1, 11, None)
    L.phpc_fill_device(dB, n, n, n, 0, 0, n, 1, 22, None)
    L.phpc_device_memset(dC, 0, n * n * 8)
    ms = L.phpc_gemm_device_timed(dA, n, dB, n, dC, n, n, n, n, 0, reps, backend)
    print(json.dumps({"n": n, "backend": sys.argv[2], "ms": ms, "tflops": 2.0 * n ** 3 / ms / 1e9}))


if __name__ == "__main__":
    main()
