"""quick_gemm_bench.py — device-resident local GEMM timing, DMMA kernel vs cuBLAS Dgemm.
Usage: python tools/quick_gemm_bench.py [N ...]   (prints one JSON line per shape)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpc_multigpu_matrixmult_b200 import capi  # noqa: E402


def run(m, k, n, reps=3):
    L = capi.load()
    ld_a, ld_b = (k + 15) // 16 * 16, (n + 15) // 16 * 16
    dA = L.phpc_device_malloc(m * ld_a * 8)
    dB = L.phpc_device_malloc(k * ld_b * 8)
    dC = L.phpc_device_malloc(m * ld_b * 8)
    L.phpc_fill_device(dA, ld_a, m, k, 0, 0, k, capi.FILL_SEEDED, 1, None)
    L.phpc_fill_device(dB, ld_b, k, n, 0, 0, n, capi.FILL_SEEDED, 2, None)
    L.phpc_device_memset(dC, 0, m * ld_b * 8)
    out = {"m": m, "k": k, "n": n}
    flops = 2.0 * m * k * n
    for name, cublas in (("dmma", 0), ("cublas", 1)):
        L.phpc_gemm_device_timed(dA, ld_a, dB, ld_b, dC, ld_b, m, k, n, 0, 1, cublas)  # warm-up
        ms = L.phpc_gemm_device_timed(dA, ld_a, dB, ld_b, dC, ld_b, m, k, n, 0, reps, cublas)
        out[name + "_ms"] = round(ms, 3)
        out[name + "_tflops"] = round(flops / ms / 1e9, 2)
    for p in (dA, dB, dC):
        L.phpc_device_free(p)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    L = capi.load()
    L.phpc_b200_set_device(0)
    shapes = sys.argv[1:] or ["2048", "4096", "8192", "16384"]
    for s in shapes:
        if "x" in s:
            m, k, n = (int(v) for v in s.split("x"))
        else:
            m = k = n = int(s)
        run(m, k, n)
