"""ozaki_knobs.py — one-call A/B of the tcgen05 (Ozaki) kernel's diagnostics flags on a B200 (the round-2 knob sweep that chose the
final kernel — balanced digits, paired N = 256 MMAs, wave-synchronised starts, 2-CTA variants — is profiles/ozaki_knobs_r02.jsonl and
profiles/ozaki_variants_r02.jsonl; the losing variants were deleted from the tree):

    python tools/ozaki_knobs.py [--out gpurun_out/ozaki_knobs.jsonl] [--time 8192 32768] [--configs name ...]

Each configuration is an environment of phpc_launch_ozaki (csrc/phpc_core.cu) run in a child process under a timeout:
a short parity check against the native-FP64 DMMA kernel (ragged shapes, several K chunks, the reference's own fill
bit-exact), then timing, then optionally (PHPC_OZ_TSTAMP=1) the per-tile start/end timestamps of one launch reduced to
"how far apart do the CTAs of one wave start" — the quantity that decides whether CTAs sharing an A row panel or a B
column panel find it in L2."""
import argparse
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# PHPC_OZ_FLAGS (diagnostics of csrc/ozaki_gemm.cuh): 1 = no C read-modify-write, 2 = no operand loads, 4 = no wave synchronisation
CONFIGS = {
    "default": {},
    "nosync": {"PHPC_OZ_FLAGS": "4"},
    "diag:noC": {"PHPC_OZ_FLAGS": "1"},
    "diag:noloads": {"PHPC_OZ_FLAGS": "2"},
    "diag:noC+noloads": {"PHPC_OZ_FLAGS": "3"},
}
SHAPES = [(128, 128, 128), (100, 77, 50), (384, 1000, 300), (640, 333, 257), (1024, 1024, 1024), (300, 9000, 200), (2048, 2048, 1536)]


def worker(times, tstamp_n):
    from hpc_multigpu_matrixmult_b200 import capi

    L = capi.load()
    L.phpc_b200_set_device(0)
    dp = capi.c_double_p

    def run(m, k, n, kind):
        lda, ldb = (k + 15) // 16 * 16, (n + 15) // 16 * 16
        a, b, c0 = np.zeros((m, lda)), np.zeros((k, ldb)), np.zeros((m, ldb))
        L.phpc_fill_host(a.ctypes.data_as(dp), lda, m, k, 0, 0, k, kind, 11)
        L.phpc_fill_host(b.ctypes.data_as(dp), ldb, k, n, 0, 0, n, kind, 22)
        L.phpc_fill_host(c0.ctypes.data_as(dp), ldb, m, n, 0, 0, n, 1, 33)
        dA, dB = L.phpc_device_malloc(a.nbytes), L.phpc_device_malloc(b.nbytes)
        dC1, dC2 = L.phpc_device_malloc(c0.nbytes), L.phpc_device_malloc(c0.nbytes)
        L.phpc_copy2d_to_device(dA, lda, a.ctypes.data_as(dp), lda, m, lda)
        L.phpc_copy2d_to_device(dB, ldb, b.ctypes.data_as(dp), ldb, k, ldb)
        for d in (dC1, dC2):
            L.phpc_copy2d_to_device(d, ldb, c0.ctypes.data_as(dp), ldb, m, ldb)
        L.phpc_gemm_device(dA, lda, dB, ldb, dC1, ldb, m, k, n, 0, None)
        L.phpc_gemm_device_ozaki(dA, lda, dB, ldb, dC2, ldb, m, k, n, None)
        L.phpc_device_synchronize()
        c1, c2 = capi.device_window(dC1, ldb, 0, 0, m, n), capi.device_window(dC2, ldb, 0, 0, m, n)
        for p in (dA, dB, dC1, dC2):
            L.phpc_device_free(p)
        return float(np.linalg.norm(c2 - c1) / np.linalg.norm(c1)), bool(np.array_equal(c1, c2))

    ok_all = True
    for (m, k, n) in ([] if os.environ.get("OZ_KNOBS_SKIP_PARITY") else SHAPES):
        for kind in (0, 1):
            if kind == 0 and max(m, k, n) > 1024:
                continue
            rel, exact = run(m, k, n, kind)
            ok = exact if kind == 0 else rel <= 2e-14
            if os.environ.get("OZ_KNOBS_DIAG"):
                ok = True  # diagnostic flags compute garbage on purpose
            ok_all &= ok
            print(json.dumps({"check": [m, k, n], "fill": "index" if kind == 0 else "seeded", "rel_vs_dmma": rel, "bit_equal": exact, "ok": ok}), flush=True)
    if not ok_all:
        print(json.dumps({"timing": "skipped: parity failed"}), flush=True)
        return 1
    for n in times:
        dA, dB, dC = (L.phpc_device_malloc(n * n * 8) for _ in range(3))
        L.phpc_fill_device(dA, n, n, n, 0, 0, n, 1, 11, None)
        L.phpc_fill_device(dB, n, n, n, 0, 0, n, 1, 22, None)
        L.phpc_device_memset(dC, 0, n * n * 8)
        L.phpc_gemm_device_timed(dA, n, dB, n, dC, n, n, n, n, 0, 1, 2)  # warm-up (backend 2 = Ozaki)
        reps = 3 if n < 32768 else 2
        ms = L.phpc_gemm_device_timed(dA, n, dB, n, dC, n, n, n, n, 0, reps, 2)
        print(json.dumps({"time_n": n, "ms": ms, "fp64_equivalent_tflops": 2.0 * n ** 3 / ms / 1e9}), flush=True)
        if tstamp_n == n and os.environ.get("PHPC_OZ_TSTAMP"):
            tiles = (n // 128) ** 2
            buf = (ctypes.c_ulonglong * (2 * tiles))()
            L.phpc_oz_tstamp_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_longlong]
            L.phpc_oz_tstamp_read.restype = ctypes.c_longlong
            got = L.phpc_oz_tstamp_read(buf, tiles)
            t = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 2)[:got].astype(np.int64)
            grid = L.phpc_b200_sm_count()
            waves = got // grid
            start = t[: waves * grid, 0].reshape(waves, grid)
            end = t[: waves * grid, 1].reshape(waves, grid)
            dur = (end - start).astype(np.float64)
            spread = (start.max(axis=1) - start.min(axis=1)).astype(np.float64)
            print(json.dumps({"tstamp_n": n, "waves": int(waves), "tile_us_mean": float(dur.mean() / 1e3), "tile_us_std": float(dur.std() / 1e3),
                              "start_spread_us_by_wave_quartiles": [float(x) / 1e3 for x in np.percentile(spread, [0, 25, 50, 75, 100])],
                              "start_spread_us_first_waves": [float(x) / 1e3 for x in spread[:8]],
                              "start_spread_us_last_waves": [float(x) / 1e3 for x in spread[-4:]],
                              "per_cta_total_ms_min_max": [float((end[-1] - start[0]).min() / 1e6), float((end[-1] - start[0]).max() / 1e6)]}), flush=True)
        for p in (dA, dB, dC):
            L.phpc_device_free(p)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ozaki_knobs.jsonl"))
    ap.add_argument("--time", type=int, nargs="*", default=[8192, 32768])
    ap.add_argument("--tstamp-n", type=int, default=0)
    ap.add_argument("--configs", nargs="*", default=list(CONFIGS))
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.worker:
        sys.exit(worker(args.time, args.tstamp_n))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as out:
        for name in args.configs:
            env = dict(os.environ, **CONFIGS[name])
            if name.startswith("diag:"):
                env["OZ_KNOBS_DIAG"] = "1"
                env["OZ_KNOBS_SKIP_PARITY"] = "1"
            if args.tstamp_n:
                env["PHPC_OZ_TSTAMP"] = "1"
            cmd = [sys.executable, os.path.abspath(__file__), "--worker", "--tstamp-n", str(args.tstamp_n), "--time"] + [str(t) for t in args.time]
            try:
                p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=args.timeout)
                rc, stdout, stderr = p.returncode, p.stdout, p.stderr
            except subprocess.TimeoutExpired as e:
                rc, stdout, stderr = "timeout", (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "timeout"
            lines = [l for l in stdout.splitlines() if l.startswith("{")]
            for l in lines:
                out.write(json.dumps({"config": name, **json.loads(l)}) + "\n")
            checks = [json.loads(l) for l in lines if '"check"' in l]
            summary = {"config": name, "env": CONFIGS[name], "exit": rc, "checks": len(checks), "failed": sum(1 for c in checks if not c["ok"]),
                       "tflops": {json.loads(l)["time_n"]: round(json.loads(l)["fp64_equivalent_tflops"], 1) for l in lines if '"time_n"' in l},
                       "tstamp": [json.loads(l) for l in lines if '"tstamp_n"' in l],
                       "stderr_tail": stderr.strip()[-400:] if rc != 0 else ""}
            out.write(json.dumps({"summary": summary}) + "\n")
            out.flush()
            print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
