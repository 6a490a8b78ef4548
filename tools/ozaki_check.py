"""ozaki_check.py — Ozaki (int8 tcgen05) GEMM vs the DMMA kernel on the same device inputs.
Usage: python tools/ozaki_check.py check | time [N ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpc_multigpu_matrixmult_b200 import capi  # noqa: E402


def one(L, m, k, n, kind, slices=0):
    lda, ldb = (k + 15) // 16 * 16, (n + 15) // 16 * 16
    dA, dB = L.phpc_device_malloc(m * lda * 8), L.phpc_device_malloc(k * ldb * 8)
    dC1, dC2 = L.phpc_device_malloc(m * ldb * 8), L.phpc_device_malloc(m * ldb * 8)
    L.phpc_fill_device(dA, lda, m, k, 0, 0, k, kind, 11, None)
    L.phpc_fill_device(dB, ldb, k, n, 0, 0, n, kind, 22, None)
    L.phpc_device_memset(dC1, 0, m * ldb * 8)
    L.phpc_device_memset(dC2, 0, m * ldb * 8)
    L.phpc_gemm_device(dA, lda, dB, ldb, dC1, ldb, m, k, n, 0, None)
    L.phpc_gemm_device_ozaki(dA, lda, dB, ldb, dC2, ldb, m, k, n, slices, None)
    L.phpc_device_synchronize()
    rows = min(m, 512)
    c1 = capi.device_window(dC1, ldb, 0, 0, rows, n)
    c2 = capi.device_window(dC2, ldb, 0, 0, rows, n)
    for p in (dA, dB, dC1, dC2):
        L.phpc_device_free(p)
    denom = np.linalg.norm(c1)
    rel = float(np.linalg.norm(c2 - c1) / denom) if denom > 0 else float(np.linalg.norm(c2))
    return rel, bool(np.array_equal(c1, c2)), c1, c2


def main():
    L = capi.load()
    L.phpc_b200_set_device(0)
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        shapes = [(128, 128, 256), (128, 256, 256), (256, 512, 512), (100, 77, 50), (384, 1000, 300), (1024, 1024, 1024), (2048, 9000, 1024)]
        for m, k, n in shapes:
            for kind in (0, 1):
                rel, exact, c1, c2 = one(L, m, k, n, kind)
                print(json.dumps({"m": m, "k": k, "n": n, "fill": "index" if kind == 0 else "seeded", "rel_vs_dmma": rel, "bit_equal": exact,
                                  "c_dmma": c1[0, :3].tolist(), "c_ozaki": c2[0, :3].tolist()}), flush=True)
        for s in (4, 6, 7, 8):
            rel, exact, _, _ = one(L, 512, 2048, 512, 1, s)
            print(json.dumps({"slices": s, "rel_vs_dmma": rel}), flush=True)
    elif mode == "sustain":
        import subprocess
        import threading
        import time
        n = int(sys.argv[2])
        secs = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
        m = k = n
        dA, dB, dC = (L.phpc_device_malloc(n * n * 8) for _ in range(3))
        L.phpc_fill_device(dA, n, m, k, 0, 0, k, 1, 11, None)
        L.phpc_fill_device(dB, n, k, n, 0, 0, n, 1, 22, None)
        L.phpc_device_memset(dC, 0, n * n * 8)
        for name, be in (("dmma", 0), ("ozaki", 2)):
            ms1 = L.phpc_gemm_device_timed(dA, n, dB, n, dC, n, m, k, n, 0, 1, be)
            reps = max(2, int(secs * 1000 / ms1))
            samples = []
            stop = threading.Event()

            def poll():
                while not stop.is_set():
                    out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
                    samples.append(out)
                    time.sleep(0.2)

            th = threading.Thread(target=poll)
            th.start()
            ms = L.phpc_gemm_device_timed(dA, n, dB, n, dC, n, m, k, n, 0, reps, be)
            stop.set()
            th.join()
            print(json.dumps({"n": n, "backend": name, "burst_tflops": round(2.0 * n ** 3 / ms1 / 1e9, 2), "sustained_tflops": round(2.0 * n ** 3 / ms / 1e9, 2),
                              "reps": reps, "clock_power_samples": samples[1:-1][:12]}), flush=True)
    else:
        for a in sys.argv[2:] or ["4096", "8192", "16384"]:
            m = k = n = int(a)
            lda = ldb = n
            dA, dB, dC = (L.phpc_device_malloc(n * n * 8) for _ in range(3))
            L.phpc_fill_device(dA, lda, m, k, 0, 0, k, 1, 11, None)
            L.phpc_fill_device(dB, ldb, k, n, 0, 0, n, 1, 22, None)
            L.phpc_device_memset(dC, 0, n * n * 8)
            out = {"n": n}
            for name, be in (("dmma", 0), ("ozaki", 2)):
                L.phpc_gemm_device_timed(dA, lda, dB, ldb, dC, ldb, m, k, n, 0, 1, be)
                ms = L.phpc_gemm_device_timed(dA, lda, dB, ldb, dC, ldb, m, k, n, 0, 2, be)
                out[name + "_ms"] = round(ms, 3)
                out[name + "_tflops"] = round(2.0 * n ** 3 / ms / 1e9, 2)
            print(json.dumps(out), flush=True)
            for p in (dA, dB, dC):
                L.phpc_device_free(p)


if __name__ == "__main__":
    main()
