"""ozaki_variants.py — one-call validation of the EXPERIMENTAL Ozaki kernel variants on a B200 (they were written
at the end of round 1 with no GPU budget left and have never run):

    python tools/ozaki_variants.py [--out gpurun_out/ozaki_variants.jsonl] [--time 4096 8192 16384]

For each variant (environment of phpc_launch_ozaki, csrc/phpc_core.cu)
    default            8 truncated 7-bit digits, 1-CTA kernel (the validated round-1 kernel: the control)
    balanced           PHPC_OZAKI_DIGITS=balanced                      7 balanced base-256 digits, 28 products
    2cta               PHPC_OZAKI_KERNEL=2cta                          CTA pairs, cta_group::2, M = 256
    2cta+balanced      both
    2cta-tma[+balanced] PHPC_OZAKI_KERNEL=2cta-tma                     the 2-CTA kernel loading through tensor maps (no relay warp)
    ...+kc16384        PHPC_OZ_KC=16384                                K chunks of 16384: half as many epilogues
a child process (bounded by a timeout: a hanging kernel must not take the GPU box with it) checks the variant
against the native-FP64 DMMA kernel on the same device inputs over shapes that exercise odd tile counts (the
2-CTA padding tile), ragged edges, several K chunks, zero and rescaled rows/columns and the reference's own fill
(bit-exact up to N = 1024), then times it.  One JSON line per check; a summary line per variant at the end."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "default": {},
    "balanced": {"PHPC_OZAKI_DIGITS": "balanced"},
    "2cta": {"PHPC_OZAKI_KERNEL": "2cta", "PHPC_OZ_PROGRESS": "1"},
    "2cta+balanced": {"PHPC_OZAKI_KERNEL": "2cta", "PHPC_OZAKI_DIGITS": "balanced", "PHPC_OZ_PROGRESS": "1"},
    # the 2-CTA kernel with cp.async.bulk.tensor.cta_group::2 loads instead of the relay warp
    "2cta-tma": {"PHPC_OZAKI_KERNEL": "2cta-tma", "PHPC_OZ_PROGRESS": "1"},
    "2cta-tma+balanced": {"PHPC_OZAKI_KERNEL": "2cta-tma", "PHPC_OZAKI_DIGITS": "balanced", "PHPC_OZ_PROGRESS": "1"},
    # K chunks of 16384 instead of 8192 (int32 stays exact up to 16643 / 18724): half as many epilogues per GEMM
    "default+kc16384": {"PHPC_OZ_KC": "16384"},
    "2cta+balanced+kc16384": {"PHPC_OZAKI_KERNEL": "2cta", "PHPC_OZAKI_DIGITS": "balanced", "PHPC_OZ_KC": "16384", "PHPC_OZ_PROGRESS": "1"},
}
SHAPES = [(128, 128, 128), (256, 128, 128), (128, 64, 256), (100, 77, 50), (384, 1000, 300), (640, 333, 257), (1024, 1024, 1024),
          (300, 9000, 200), (200, 17000, 130), (2048, 2048, 1536)]


ROLES = ["producer", "mma/relay", "epi0", "epi1", "epi2", "epi3", "setup", "-"]
MARKS = {0: "not started", 1: "waiting", 2: "passed", 3: "finished", 4: "waiting(tempty/cluster)", 5: "own full ok, waiting peer_full"}


def wait_or_report_hang(L, shape, seconds=20.0):
    """Poll the compute stream; if the kernel is still running after `seconds`, print where every warp role of the first
    CTA pairs stands (progress words of the 2-CTA kernel, PHPC_OZ_PROGRESS=1) and leave: the parent records a hang."""
    import ctypes
    import time

    t0 = time.time()
    while time.time() - t0 < seconds:
        if L.phpc_compute_stream_idle():
            return
        time.sleep(0.01)
    words = (ctypes.c_uint * 4096)()
    L.phpc_oz_progress_read.argtypes = [ctypes.POINTER(ctypes.c_uint), ctypes.c_int]
    n = L.phpc_oz_progress_read(words, 4096)
    ctas = []
    for c in range(min(n // 8, 6)):
        ctas.append({ROLES[r]: f"{MARKS.get(words[c * 8 + r] >> 28, words[c * 8 + r] >> 28)}@{words[c * 8 + r] & 0x0fffffff}" for r in range(7)})
    stuck = sum(1 for c in range(n // 8) if (words[c * 8 + 6] >> 28) != 3)
    print(json.dumps({"hang": list(shape), "seconds": seconds, "ctas_not_finished": stuck, "ctas_total": n // 8, "first_ctas": ctas}), flush=True)
    os._exit(3)


def worker(times):
    from hpc_multigpu_matrixmult_b200 import capi

    L = capi.load()
    L.phpc_b200_set_device(0)

    def run(m, k, n, kind, tweak=None):
        lda, ldb = (k + 15) // 16 * 16, (n + 15) // 16 * 16
        a = np.zeros((m, lda))
        b = np.zeros((k, ldb))
        dp = capi.c_double_p
        L.phpc_fill_host(a.ctypes.data_as(dp), lda, m, k, 0, 0, k, kind, 11)
        L.phpc_fill_host(b.ctypes.data_as(dp), ldb, k, n, 0, 0, n, kind, 22)
        if tweak == "scaled":
            a[:, :k] *= np.ldexp(1.0, (np.arange(m) % 13) * 40 - 250)[:, None]
            b[:k, :n] *= np.ldexp(1.0, (np.arange(n) % 11) * 30 - 100)[None, :]
            a[m // 2, :] = 0.0
            b[:, n // 3] = 0.0
        c0 = np.zeros((m, ldb))
        L.phpc_fill_host(c0.ctypes.data_as(dp), ldb, m, n, 0, 0, n, 1, 33)
        dA, dB = L.phpc_device_malloc(a.nbytes), L.phpc_device_malloc(b.nbytes)
        dC1, dC2 = L.phpc_device_malloc(c0.nbytes), L.phpc_device_malloc(c0.nbytes)
        L.phpc_copy2d_to_device(dA, lda, a.ctypes.data_as(dp), lda, m, lda)
        L.phpc_copy2d_to_device(dB, ldb, b.ctypes.data_as(dp), ldb, k, ldb)
        for d in (dC1, dC2):
            L.phpc_copy2d_to_device(d, ldb, c0.ctypes.data_as(dp), ldb, m, ldb)
        L.phpc_gemm_device(dA, lda, dB, ldb, dC1, ldb, m, k, n, 0, None)
        L.phpc_device_synchronize()
        L.phpc_gemm_device_ozaki(dA, lda, dB, ldb, dC2, ldb, m, k, n, 0, None)  # enqueues only
        wait_or_report_hang(L, (m, k, n))
        L.phpc_device_synchronize()
        c1 = capi.device_window(dC1, ldb, 0, 0, m, n)
        c2 = capi.device_window(dC2, ldb, 0, 0, m, n)
        for p in (dA, dB, dC1, dC2):
            L.phpc_device_free(p)
        # per-row relative difference: rows differ by 2^±250 under "scaled", a global norm would hide the small ones
        num = np.linalg.norm(c2 - c1, axis=1)
        den = np.linalg.norm(c1, axis=1)
        rel_rows = np.where(den > 0, num / np.where(den > 0, den, 1.0), num)
        return float(rel_rows.max()), bool(np.array_equal(c1, c2))

    ok_all = True
    for (m, k, n) in SHAPES:
        for kind, tweak in ((0, None), (1, None), (1, "scaled")):
            if kind == 0 and max(m, k, n) > 1024:
                continue
            rel, exact = run(m, k, n, kind, tweak)
            # index fill with integer partial sums < 2^53: every summation order is exact -> must be bit-equal
            ok = exact if (kind == 0 and m * k <= (1 << 20) and k <= 1024 and n <= 1024) else rel <= 2e-14
            ok_all &= ok
            print(json.dumps({"check": [m, k, n], "fill": "index" if kind == 0 else "seeded", "tweak": tweak, "max_row_rel_vs_dmma": rel,
                              "bit_equal": exact, "ok": ok}), flush=True)
    if not ok_all:
        print(json.dumps({"timing": "skipped: parity failed"}), flush=True)
        return 1
    for n in times:
        dA, dB, dC = (L.phpc_device_malloc(n * n * 8) for _ in range(3))
        L.phpc_fill_device(dA, n, n, n, 0, 0, n, 1, 11, None)
        L.phpc_fill_device(dB, n, n, n, 0, 0, n, 1, 22, None)
        L.phpc_device_memset(dC, 0, n * n * 8)
        L.phpc_gemm_device_timed(dA, n, dB, n, dC, n, n, n, n, 0, 1, 2)  # warm-up (backend 2 = Ozaki)
        ms = L.phpc_gemm_device_timed(dA, n, dB, n, dC, n, n, n, n, 0, 3, 2)  # mean of 3
        print(json.dumps({"time_n": n, "ms": ms, "fp64_equivalent_tflops": 2.0 * n ** 3 / ms / 1e9}), flush=True)
        for p in (dA, dB, dC):
            L.phpc_device_free(p)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ozaki_variants.jsonl"))
    ap.add_argument("--time", type=int, nargs="*", default=[4096, 8192, 16384])
    ap.add_argument("--variants", nargs="*", default=list(VARIANTS))
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.worker:
        sys.exit(worker(args.time))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as out:
        for name in args.variants:
            env = dict(os.environ, **VARIANTS[name])
            cmd = [sys.executable, os.path.abspath(__file__), "--worker", "--time"] + [str(t) for t in args.time]
            try:
                p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=args.timeout)
                rc, stdout, stderr = p.returncode, p.stdout, p.stderr
            except subprocess.TimeoutExpired as e:
                rc, stdout, stderr = "timeout", (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "timeout"
            lines = [l for l in stdout.splitlines() if l.startswith("{")]
            for l in lines:
                out.write(json.dumps({"variant": name, **json.loads(l)}) + "\n")
            checks = [json.loads(l) for l in lines if '"check"' in l]
            hangs = [json.loads(l) for l in lines if '"hang"' in l]
            summary = {"variant": name, "exit": rc, "checks": len(checks), "failed": sum(1 for c in checks if not c["ok"]),
                       "tflops": {json.loads(l)["time_n"]: round(json.loads(l)["fp64_equivalent_tflops"], 1) for l in lines if '"time_n"' in l},
                       "hang": hangs[0] if hangs else None, "stderr_tail": stderr.strip()[-400:] if rc != 0 else ""}
            out.write(json.dumps({"summary": summary}) + "\n")
            out.flush()
            print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
