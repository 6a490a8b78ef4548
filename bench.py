#!/usr/bin/env python
"""
bench.py — SUMMA GEMM TFLOP/s at N=32768 on 1/2/4/8 B200 (BASELINE.json's metric).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 4 --steps 3 --warmup 3
    python bench.py --impl reference ...      # the reference's own CPU GEMM on the host cores

A "step" is one whole SUMMA  C += A*B  (FP64, N x N, 2*N^3 flop) on the r x c process
grid (1x1, 1x2, 2x2, 2x4 for 1, 2, 4, 8 GPUs), one process per GPU.  N is fixed as the
GPU count grows ("scaling": "strong").
  value  device-resident: every rank's owned A/B blocks already in HBM, timed with CUDA
         events on the launching stream around K steps, max over ranks; afterwards every rank checks
         sampled elements of its C block against exactly summed FP64 dot products (`value_verified`).
  e2e    the reference-facing C-ABI call phpc_gemm_summa_cuda() on FULL N x N HOST
         matrices: upload of the owned blocks, the k-loop, download of C and the gather
         to rank 0 are all inside the timed region (wall clock around a synchronous call,
         bracketed by barriers, max over ranks); verified on every rank, rank 0 across every block.
  roofline  the tcgen05 GEMM kernel (phpc::oz::ozaki_gemm_kernel): 2*m*k*n*28 int8 operations per
         local GEMM / mean local-GEMM duration (CUDA events recorded around every launch inside the
         timed region) against the int8 tensor peaks measured on this pool with tools/umma_rate.cu
         (burst: profiles/umma_rate_r01.jsonl; sustained on random operands: profiles/umma_rate_sustained_r02.jsonl);
         the native-FP64 DMMA kernel and cuBLAS Dgemm are measured in the same run and reported beside it.
  cpu_baseline  the reference's iterative.c (oracle/_ref, compiled unchanged with the
         reference's flags) on ONE host core at a bounded size; beside it the reference's own
         SUMMA on all host cores and the reference's own CUDA+MPI build on the same GPUs.
torch is plumbing only (process group, events, synchronize); all math goes through
libphpc_b200.so.  The oracle is touched only by the cpu_baseline / --impl reference legs.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "summa_gemm_tflops"
UNIT = "TFLOP/s"
GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
CPU_SAMPLE_N = 1024
DMMA_N32768_DRAM_BYTES = 578.8e9 + 68.8e9  # ncu, the 8 K-chunk launches of one N=32768 GEMM (profiles/ncu_dmma_n32768_dram_r02.csv; one launch in r01: 1522 GB)


# ----------------------------------------------------------------------------
# CPU arm: the reference's own iterative.c
# ----------------------------------------------------------------------------
def _iterative_binary(opt):
    path = os.path.join(ROOT, "oracle", "_ref", f"iterative_{opt}.out")
    return path if os.path.exists(path) else None


def run_reference_cpu(n, opt="O0"):
    """One run of the reference's CPU GEMM; returns (tflops, seconds, kind)."""
    exe = _iterative_binary(opt)
    if exe:
        out = subprocess.run([exe, str(n)], check=True, capture_output=True, text=True).stdout.strip()
        secs = float(out.split(",")[1])  # reference prints "n,seconds" (src/iterative.c:39)
        kind = "reference"
    else:  # /root/reference was absent at build time: time the restatement instead
        from oracle import oracle

        a = oracle.fill(n, n, kind=0)
        t0 = time.perf_counter()
        oracle.gemm_iterative(a, a)
        secs = time.perf_counter() - t0
        kind = "port"
    return 2.0 * n ** 3 / secs / 1e12, secs, kind


def run_reference_summa_all_cores(n=4096):
    """The reference's own SUMMA (src/phpc_summa.c compiled unchanged, oracle/_ref/ref_summa_cpu.out) on as many host
    cores as it can use: one MPI rank per core (a power of two, at most 16) under the shim, the local GEMM plugin being
    the CPU restatement of the reference kernel's arithmetic (oracle/ref_cpu_plugin.c; the reference itself only ships a
    GPU plugin).  Reported beside the single-threaded iterative.c number; returns None if it cannot run here."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_summa_cpu.out")
    mpirun = os.path.join(ROOT, "bin", "mpirun")
    if not (os.path.exists(exe) and os.path.exists(mpirun)):
        return None
    ranks = 1
    while ranks * 2 <= min(os.cpu_count() or 1, 16):
        ranks *= 2
    try:
        p = subprocess.run([mpirun, "-n", str(ranks), exe, str(n), "1", "-"], capture_output=True, text=True, timeout=600)
        f = p.stdout.strip().split(",")
        secs = float(f[4])
    except (subprocess.TimeoutExpired, OSError, ValueError, IndexError):
        return None
    return {"value": 2.0 * n ** 3 / secs / 1e12, "unit": UNIT, "cores": ranks, "kind": "port", "seconds": secs,
            "sample": f"reference phpc_summa.c (unchanged) on a {f[2]}x{f[3]} grid of {ranks} MPI-shim ranks, N={n}, CPU gemm_t plugin = "
                      "restatement of the reference kernel's per-element loop (gcc -O2), host-memory broadcasts included"}


def run_reference_cuda_build(n=8192, tile=32, gw=64, gh=64, timeout=300, ranks=1):
    """The reference's OWN CUDA+MPI program on this box: oracle/_ref/ref_main.out = /root/reference/src/{main.c, phpc_summa.c,
    phpc_gemm.cu, utils.c} compiled unchanged against the MPI shim (oracle/Makefile), one rank, one GPU.  It times its CUDA pass
    (host-memory SUMMA + H2D + gemm_kernel + D2H, src/main.c:93-95) and its cuBLASXt pass (:105-107) itself and writes the
    reference's CSV record (src/utils.c:26-27); this returns those numbers as TFLOP/s.  Launch grid gw x gh CTAs of tile x tile
    threads (the reference's own sweeps use 1 x 1 .. 8 x 8 CTAs; 64 x 64 is the generous setting)."""
    import tempfile

    exe = os.path.join(ROOT, "oracle", "_ref", "ref_main.out")
    mpirun = os.path.join(ROOT, "bin", "mpirun")
    if not (os.path.exists(exe) and os.path.exists(mpirun)):
        return {"unavailable": "oracle/_ref/ref_main.out or bin/mpirun not built"}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "csv"))
        if ranks == 1:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
        else:  # the reference uses EVERY visible GPU in every rank (src/main.c:54-56): one GPU per rank through the launcher
            env = dict(os.environ, PHPC_GPU_POLICY="visible", PHPC_GPUS=str(ranks))
            env.pop("CUDA_VISIBLE_DEVICES", None)
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "PHPC_MPI_SHM", "PHPC_MPI_RANK", "PHPC_MPI_SIZE"):
            env.pop(k, None)  # the child world is the launcher's, not torchrun's
        try:
            p = subprocess.run([mpirun, "-n", str(ranks), exe, str(n), str(tile), str(gw), str(gh), "bench"], cwd=tmp, env=env, capture_output=True,
                               text=True, timeout=timeout)
        except subprocess.TimeoutExpired:
            return {"unavailable": f"timed out after {timeout} s at N={n}"}
        files = os.listdir(os.path.join(tmp, "csv"))
        if p.returncode != 0 or not files:
            return {"unavailable": f"exit {p.returncode}: {p.stderr.strip()[-200:]}"}
        rec = open(os.path.join(tmp, "csv", files[0])).read().strip().split(",")
    cuda_s, kernel_s, cublas_s = float(rec[6]), float(rec[7]), float(rec[8])
    fl = 2.0 * n ** 3 / 1e12
    return {"N": n, "launch": f"{gw}x{gh} CTAs of {tile}x{tile} threads", "ranks": ranks, "gpus": ranks,
            "cuda_pass_tflops": fl / cuda_s if cuda_s > 0 else None, "cuda_pass_s": cuda_s,
            "gemm_kernel_tflops": fl / kernel_s if kernel_s > 0 else None, "gemm_kernel_s": kernel_s,
            "cublasxt_pass_tflops": fl / cublas_s if cublas_s > 0 else None, "cublasxt_pass_s": cublas_s,
            "what": "the reference's unmodified main.c + phpc_summa.c + phpc_gemm.cu (oracle/_ref/ref_main.out), timed by itself: wall time of "
                    "its CUDA pass (includes CUDA context creation, host pinning, H2D/D2H per call) and of its cuBLASXt pass, "
                    "event time of its kernel"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = CPU_SAMPLE_N
    for _ in range(args.warmup):
        run_reference_cpu(n)
    vals, secs = [], []
    for _ in range(args.steps):
        v, s, kind = run_reference_cpu(n)
        vals.append(v)
        secs.append(s)
    value = sum(vals) / len(vals)
    sample = f"iterative.c (gcc -Wall, no -O: reference Makefile:9), N={n}, A[i]=B[i]=i, 1 thread; TFLOP/s = 2N^3/t"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SUMMA GEMM N={args.n} FP64 (bounded CPU sample N={n})", "N": args.n, "sample_N": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "host_cores": os.cpu_count(),
                         "reference_summa_all_cores": run_reference_summa_all_cores()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.lines:
            if not (t0 <= t <= t1 + 0.2):
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------
def fp64_peak():
    """FP64 roofline denominator measured on this pool (tools/fp64_peak.cu -> profiles/)."""
    path = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    nominal = 148 * 64 * 2 * 1.965e9 / 1e12  # 64 FP64 FMA/clk/SM at clocks.max.sm
    try:
        d = json.load(open(path))
        dgemm = max(x["burst_tflops"] for x in d["dgemm"])
        issue = max(max(d["dfma_tflops"].values()), max(d["dmma_tflops"].values()))
        return max(dgemm, issue), (f"measured FP64 peak on this pool: max(DFMA/DMMA issue-rate microbenchmark {issue:.2f}, cuBLAS Dgemm burst "
                                   f"{dgemm:.2f}) TFLOP/s, profiles/fp64_peak_r01.json; nominal 148 SM x 64 FMA/clk x 1965 MHz = {nominal:.1f}")
    except (OSError, KeyError, ValueError):
        return nominal, f"fallback: nominal FP64 148 SM x 64 FMA/clk x 1965 MHz = {nominal:.1f} TFLOP/s (profiles/fp64_peak_r01.json missing)"


def int8_peak():
    """int8 tensor roofline: tcgen05.mma kind::i8 issue-rate microbenchmark on this pool (tools/umma_rate.cu)."""
    path = os.path.join(ROOT, "profiles", "umma_rate_r01.jsonl")
    try:
        best = max(json.loads(l)["chip_tops_at_event_time"] for l in open(path) if '"kind": "i8"' in l)
        return best, f"measured int8 tensor peak {best:.0f} TOP/s (tcgen05.mma kind::i8 128x256x32 issue rate, all SMs, burst clocks; profiles/umma_rate_r01.jsonl)"
    except (OSError, ValueError, KeyError):
        return 2.0 * 1590.0, "fallback: 2 x bf16 fallback peak (profiles/umma_rate_r01.jsonl missing)"


def int8_sustained_peak():
    """Sustained (seconds-long, power-capped) int8 tensor peak on RANDOM operands, if it has been measured on this pool:
    UMMA_SUSTAIN=1 bin/umma_rate -> profiles/umma_rate_sustained*_r*.jsonl (scripts/gpu_session.sh, stage `peaks`)."""
    import glob

    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "umma_rate_sustained*.jsonl"))):
        try:
            lines = open(path).read().splitlines()
        except OSError:
            continue
        for l in lines:
            try:
                d = json.loads(l)
            except ValueError:
                continue  # a stray non-JSON line must not hide the measurement
            if isinstance(d, dict) and d.get("operands") == "random" and "chip_tops_sustained" in d:
                best = (d["chip_tops_sustained"], os.path.relpath(path, ROOT))
    return best


def ozaki_traffic(m, k, n):
    """DRAM bytes of one Ozaki GEMM launch of this shape from an `ncu --set full` capture, if one is committed:
    profiles/ozaki_traffic*.json written by tools/ncu_summary.py --traffic-json --shape m,k,n (scripts/gpu_session.sh)."""
    import glob

    hit = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ozaki_traffic*.json"))):
        try:
            d = json.load(open(path))
            if d.get("shape_mkn") == [int(m), int(k), int(n)]:
                hit = (d["dram_bytes"], os.path.relpath(path, ROOT), d.get("kernel"))
        except (OSError, ValueError, KeyError):
            continue
    return hit


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written on this pool: HBM copy GB/s, cuBLAS bf16 burst / sustained TFLOP/s), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return None


def sampled_check(capi, L, N, pairs, got, passes):
    """Self-verification of a bench result: `got[(i, j)]` = C[i][j] after `passes` accumulations of A*B, for global element
    positions `pairs`.  Rows of A / columns of B are regenerated on the host from the counter-based fill the device blocks
    were generated with (the library's own phpc_fill_host: same splitmix64 stream), the dot product is summed exactly
    (math.fsum over FP64 products), and the bound is the one the parity tests use: 4*sqrt(N)*2^-53 * sum|a||b| per pass.
    Returns (ok, worst error / bound)."""
    import math

    import numpy as np

    dp = capi.c_double_p
    rows = sorted({i for i, _ in pairs})
    cols = sorted({j for _, j in pairs})
    a = {}
    for i in rows:
        v = np.empty(N)
        L.phpc_fill_host(v.ctypes.data_as(dp), N, 1, N, int(i), 0, N, capi.FILL_SEEDED, capi.SEED_A)
        a[i] = v
    b = {}
    for j in cols:
        v = np.empty(N)
        L.phpc_fill_host(v.ctypes.data_as(dp), 1, N, 1, 0, int(j), N, capi.FILL_SEEDED, capi.SEED_B)
        b[j] = v
    worst = 0.0
    for (i, j) in pairs:
        prod = a[i] * b[j]
        want = passes * math.fsum(prod)
        bound = passes * 4.0 * math.sqrt(N) * 2.0 ** -53 * float(np.abs(prod).sum())
        err = abs(float(got[(i, j)]) - want)
        if not err <= bound:  # also catches NaN
            return False, float("inf")
        worst = max(worst, err / bound)
    return True, worst


def sample_positions(row0, rows, col0, cols, seed, count=8):
    """count x count element positions of the block [row0, row0+rows) x [col0, col0+cols): first and last row / column
    (first and last tile, band and chunk edges) plus seeded random ones."""
    import numpy as np

    rng = np.random.default_rng(seed)
    ri = sorted({row0, row0 + rows - 1, *(int(x) for x in rng.integers(row0, row0 + rows, count - 2))})
    ci = sorted({col0, col0 + cols - 1, *(int(x) for x in rng.integers(col0, col0 + cols, count - 2))})
    return [(i, j) for i in ri for j in ci]


def host_matrices(capi, N, dims, coords, pin=True):
    """FULL N x N host A, B, C as the reference API wants them; only the windows this rank
    owns are filled and page-locked (the rest of the address range is never touched)."""
    import numpy as np

    L = capi.load()
    r, c = dims
    pi, pj = coords
    m, n = N // r, N // c
    A = np.empty((N, N), dtype=np.float64)
    B = np.empty((N, N), dtype=np.float64)
    # rank 0's C receives every block: in node-shared page-locked memory (phpc_host_malloc_shared) all ranks write their block
    # into it over their own PCIe link; plain memory (the fallback when /dev/shm is too small) gives the serial gather
    shared = capi.host_array_shared(N, N) if (r * c > 1 and pi == 0 and pj == 0) else None
    C = shared[0] if shared else np.empty((N, N), dtype=np.float64)
    dp = capi.c_double_p
    lcm = r * c // __import__("math").gcd(r, c)
    pk = N // lcm

    def fill_rows(mat, row0, rows, col0, cols, seed):
        nthreads = min(8, os.cpu_count() or 1, max(1, rows // 64))
        per = (rows + nthreads - 1) // nthreads

        def work(t):
            lo, hi = row0 + t * per, min(row0 + rows, row0 + (t + 1) * per)
            if lo < hi:
                ptr = ctypes.cast(mat.ctypes.data + (lo * N + col0) * 8, dp)
                L.phpc_fill_host(ptr, N, hi - lo, cols, lo, col0, N, capi.FILL_SEEDED, seed)

        ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
        [t.start() for t in ts]
        [t.join() for t in ts]

    regs = []
    # A: rows of this process row, the K panels this process column owns
    for k in range(lcm):
        if k % c == pj:
            fill_rows(A, pi * m, m, k * pk, pk, capi.SEED_A)
    # B: the K panels this process row owns, columns of this process column
    for k in range(lcm):
        if k % r == pi:
            fill_rows(B, k * pk, pk, pj * n, n, capi.SEED_B)
            regs.append((B.ctypes.data + (k * pk * N) * 8, pk * N * 8))
    regs.append((A.ctypes.data + (pi * m * N) * 8, m * N * 8))
    C[pi * m:(pi + 1) * m, pj * n:(pj + 1) * n] = 0.0
    if shared:
        pass  # page-locked by the allocator
    elif pi == 0 and pj == 0:
        regs.append((C.ctypes.data, N * N * 8))  # the gather root receives every block: keep its whole C page-locked
    else:
        regs.append((C.ctypes.data + (pi * m * N) * 8, m * N * 8))
    if pin:
        for ptr, nbytes in regs:
            L.phpc_host_register(ptr, nbytes)
    return A, B, C, regs, (shared[1] if shared else None)


def e2e_measure(args, capi, L, comm, dims, rank, world, barrier, max_over_ranks):
    """`e2e`: the reference-facing phpc_gemm_summa_cuda() on page-locked FULL N x N host matrices, wall clock
    around the synchronous call (H2D of the owned blocks and of C, the k-loop, D2H of C and the gather to
    rank 0 all inside).  On one GPU sampled elements of the host result are checked against FP64 dot
    products of the host rows/columns; a number is only reported for a result that passed."""
    import numpy as np
    import torch

    N = args.n
    flops = 2.0 * N ** 3
    coords = (rank // dims[1], rank % dims[1])
    A, B, C, regs, shared_c = host_matrices(capi, N, dims, coords)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    pi, pj = coords
    m_blk, n_blk = N // dims[0], N // dims[1]

    def leg():
        """Warm-up call + timed calls on a zeroed C; returns (mean seconds, number of C += A*B passes)."""
        C[pi * m_blk:(pi + 1) * m_blk, pj * n_blk:(pj + 1) * n_blk] = 0.0
        capi.phpc_gemm_summa_cuda(comm, A, B, C)  # warm-up: allocates the cached device blocks
        ts = []
        for _ in range(e2e_steps):
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            capi.phpc_gemm_summa_cuda(comm, A, B, C)
            torch.cuda.synchronize()
            barrier()
            ts.append(time.perf_counter() - t0)
        return sum(ts) / len(ts), e2e_steps + 1

    def verified(passes):
        """Rank 0 holds the full gathered C: sampled elements of EVERY rank's block (so the gather is checked too); the other
        ranks check their own block, which the entry point leaves at its global offset in their C (reference :44)."""
        pairs = []
        if rank == 0:
            for bi in range(dims[0]):
                for bj in range(dims[1]):
                    pairs += sample_positions(bi * m_blk, m_blk, bj * n_blk, n_blk, 2026 + bi * 16 + bj, 5 if world > 1 else 8)
        else:
            pairs = sample_positions(pi * m_blk, m_blk, pj * n_blk, n_blk, 2026 + rank, 5)
        ok, worst = sampled_check(capi, L, N, pairs, {(i, j): C[i, j] for (i, j) in pairs}, passes)
        return ok, worst, len(pairs)

    secs, passes = leg()
    ok, worst, nsamples = verified(passes)
    ok = max_over_ranks(0.0 if ok else 1.0) == 0.0  # every rank's check must pass
    worst = max_over_ranks(worst)
    secs = max_over_ranks(secs)
    for ptr, _ in regs:
        L.phpc_host_unregister(ptr)
    L.phpc_summa_release_cache()
    gather = "n/a (one rank)" if world == 1 else ("parallel: rank 0's C in node-shared page-locked memory, every rank writes its block over its own PCIe link"
                                                 if max_over_ranks(1.0 if (rank == 0 and shared_c) else 0.0) > 0 else "serial through rank 0's GPU and PCIe link")
    if shared_c:
        del C
        L.phpc_host_free_shared(shared_c)
    bands = os.environ.get("PHPC_HOST_BANDS", "4 (default for blocks of >= 8192 rows)") if world == 1 else "n/a"
    return {"value": flops / secs / 1e12 if ok is not False else None, "unit": UNIT, "h2d_bytes_per_step": 3 * 8 * N * N,
            "d2h_bytes_per_step": 8 * N * N, "ms_per_step": secs * 1e3, "steps": e2e_steps, "verified": ok,
            "verify": {"elements_checked_rank0": nsamples, "worst_error_over_bound": worst, "passes_accumulated": passes,
                       "how": "sampled C elements of every rank's block in rank 0's gathered host C (and each rank's own block) vs exactly summed FP64 "
                              "dot products of regenerated rows/columns; bound 4*sqrt(N)*2^-53*sum|a||b| per pass"},
            "host_row_bands": bands, "gather": gather,
            "api": "phpc_gemm_summa_cuda(grid_comm, A, B, C, N, ...) on page-locked full N x N host matrices; owned blocks H2D, C block "
                   "H2D and D2H + gather to rank 0 inside the timed region (one GPU: C row bands pipelined under the GEMMs)"}


def e2e_child(args):
    """`bench.py --e2e-child`: the single-GPU e2e leg in its own process; prints the e2e object as one JSON line."""
    import torch

    from hpc_multigpu_matrixmult_b200 import capi

    L = capi.load()
    if not torch.cuda.is_available() or L.phpc_b200_device_count() < 1:
        raise SystemExit("bench.py needs a B200: no CUDA device visible and there is no CPU fallback")
    torch.cuda.set_device(0)
    L.phpc_b200_set_device(0)
    capi.mpi_init(0, 1, None)
    comm = capi.cart_create((1, 1))
    out = e2e_measure(args, capi, L, comm, (1, 1), 0, 1, lambda: None, lambda x: x)
    print("E2E_JSON " + json.dumps(out), flush=True)


def e2e_in_child(args):
    """Run the e2e leg as a child process.  If the band-pipelined host path crashes or fails its check, the
    chunk-pipelined loop (PHPC_HOST_BANDS=1, the path the round-1 numbers were taken with) is measured instead."""
    cmd = [sys.executable, os.path.abspath(__file__), "--e2e-child", "--n", str(args.n), "--steps", str(args.steps),
           "--e2e-steps", str(args.e2e_steps)]
    attempts = [dict(os.environ)]
    if os.environ.get("PHPC_HOST_BANDS") != "1":
        attempts.append(dict(os.environ, PHPC_HOST_BANDS="1"))
    note = None
    for env in attempts:
        try:
            p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=420)
            lines = [l for l in p.stdout.splitlines() if l.startswith("E2E_JSON ")]
            if p.returncode == 0 and lines:
                out = json.loads(lines[-1][len("E2E_JSON "):])
                if out.get("verified") is not False:
                    if note:
                        out["note"] = note
                    return out
                note = "band pipeline failed its sampled verification; re-measured with PHPC_HOST_BANDS=1"
            else:
                note = f"e2e child exited {p.returncode}: {p.stderr.strip()[-300:]}; re-measured with PHPC_HOST_BANDS=1"
        except subprocess.TimeoutExpired:
            note = "e2e child timed out; re-measured with PHPC_HOST_BANDS=1"
    return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "verified": False, "note": note}


def product_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from hpc_multigpu_matrixmult_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if world not in GRIDS:
        raise SystemExit("supported GPU counts: 1, 2, 4, 8")
    L = capi.load()  # raises if the CUDA library is missing: no fallback
    if not torch.cuda.is_available() or L.phpc_b200_device_count() < 1:
        raise SystemExit("bench.py needs a B200: no CUDA device visible and there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    L.phpc_b200_set_device(local_rank)

    seg = None
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        box = [None]
        if rank == 0:
            seg = f"/dev/shm/phpc_bench_{os.getpid()}_{int(time.time())}"
            capi.mpi_segment_create(seg, world)
            box[0] = seg
        dist.broadcast_object_list(box, src=0)
        seg = box[0]
    capi.mpi_init(rank, world, seg)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    N = args.n
    dims = GRIDS[world]
    comm = capi.cart_create(dims)
    flops = 2.0 * N ** 3

    # ---------------- device-resident SUMMA: `value` ----------------
    # primary = the kernel the reference-named entry points run (tcgen05 Ozaki unless PHPC_GEMM=dmma);
    # the other kernel is measured right after it on the same blocks and reported beside it.
    primary = capi.BACKEND_DMMA if os.environ.get("PHPC_GEMM") == "dmma" else capi.BACKEND_OZAKI
    secondary = capi.BACKEND_OZAKI if primary == capi.BACKEND_DMMA else capi.BACKEND_DMMA
    names = {capi.BACKEND_DMMA: "native_fp64_dmma", capi.BACKEND_OZAKI: "tcgen05_ozaki"}
    s = capi.Summa(comm, N, args.kc)
    s.fill(capi.FILL_SEEDED)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    oz = capi.ozaki_config()  # the fixed arithmetic of the library's tcgen05 path
    slices, pairs = oz["digits"], oz["products"]
    oz_kernel = ("phpc::oz::ozaki_gemm_kernel (tcgen05.mma.cta_group::1.kind::i8 M=128, N=256 over adjacent digit pairs, int32 accumulators in all 512 "
                 "TMEM columns, cp.async.bulk mbarrier ring, warp-specialised, wave-synchronised tile starts)")
    fp64_pk, fp64_src = fp64_peak()
    int8_pk, int8_src = int8_peak()

    passes_on_c = [0]  # C += A*B passes accumulated in the device C blocks since s.fill() zeroed them
    primary_summa = s

    def measure(backend, warmup, steps, sample_clocks, s=s):
        if s is primary_summa:
            passes_on_c[0] += warmup + steps
        for _ in range(warmup):
            s.run(backend, 0, sptr, stats=False)
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0 and sample_clocks:
            sampler.start()
            time.sleep(0.3)
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.perf_counter()
        e0.record(stream)
        st = None
        for i in range(steps):
            st = s.run(backend, 0, sptr, stats=(i == steps - 1))  # per-launch events are read on the last step
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        t_end = time.perf_counter()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        clocks = sampler.stop(t_begin, t_end) if (rank == 0 and sample_clocks) else None
        m_blk, n_blk = s.block
        k_per_gemm = N / st.steps
        gemm_ms = st.gemm_ms / st.steps  # mean duration of one local GEMM (device events around every launch)
        gemm_tflops = 2.0 * m_blk * n_blk * k_per_gemm / (gemm_ms * 1e-3) / 1e12
        if backend == capi.BACKEND_OZAKI:
            roof = {"bound": "tensor", "achieved": gemm_tflops * pairs, "peak": int8_pk, "unit": "TOP/s", "frac": gemm_tflops * pairs / int8_pk,
                    "traffic": None, "kernel": oz_kernel,
                    "ops_per_launch": 2.0 * m_blk * n_blk * k_per_gemm * pairs, "kernel_ms": gemm_ms, "fp64_equivalent_tflops": gemm_tflops,
                    "fp64_equivalent_vs_fp64_peak": gemm_tflops / fp64_pk, "peak_source": int8_src,
                    "note": f"{pairs} int8 MMAs per FP64 MMA ({slices} balanced base-256 digits); kernel_ms spans the whole local GEMM (exponent, guard and split "
                            "kernels included, < 3 % at this size)",
                    "traffic_note": "no ncu capture of this launch shape committed under profiles/ozaki_traffic*.json"}
            k_launch = min(int(k_per_gemm), oz["k_chunk"])  # the launcher cuts K into chunks of this size
            tr = ozaki_traffic(m_blk, k_launch, n_blk)
            if tr:
                roof["traffic"] = tr[0]
                roof["traffic_source"] = (f"ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch of {tr[2]} at m,k,n = "
                                          f"{m_blk},{k_launch},{n_blk} ({tr[1]}); algorithmic bytes of that launch: "
                                          f"{(m_blk + n_blk) * k_launch * slices + 2 * 8 * m_blk * n_blk} (digits read once + one read and one write of C); with a "
                                          "126 MB L2 each wave of 148 tiles has to stream its 16 + ~9 digit panels once: 443 waves x ~25 panels x 12 MiB "
                                          "= 146 GB is the floor of this tiling")
                roof.pop("traffic_note", None)
            sus = int8_sustained_peak()
            if sus:
                roof["peak_sustained"] = sus[0]
                roof["frac_of_sustained"] = gemm_tflops * pairs / sus[0]
                roof["peak_sustained_source"] = f"tcgen05.mma kind::i8 128x256x32 on random operand bytes, seconds-long run under the power cap ({sus[1]})"
            mp = measured_peaks()
            if mp and mp.get("bf16_tflops") and mp.get("bf16_tflops_sustained"):
                # the step is seconds long and power capped: scale the burst int8 peak by the driver's sustained/burst bf16 ratio,
                # and show the fraction against 2 x the driver's own bf16 numbers (int8 = 2 x bf16 nominally) beside it
                ratio = mp["bf16_tflops_sustained"] / mp["bf16_tflops"]
                roof["frac_of_sustained_estimate"] = gemm_tflops * pairs / (int8_pk * ratio)
                roof["sustained_estimate_source"] = (f"burst int8 peak x MEASURED_PEAKS.json bf16 sustained/burst ({ratio:.3f}); a sustained int8 "
                                                     "microbenchmark with random operands is tools/umma_rate.cu UMMA_SUSTAIN=1 (not yet run)")
                roof["frac_of_2x_measured_bf16_sustained"] = gemm_tflops * pairs / (2.0 * mp["bf16_tflops_sustained"])
        else:
            roof = {"bound": "tensor", "achieved": gemm_tflops, "peak": fp64_pk, "unit": UNIT, "frac": gemm_tflops / fp64_pk,
                    "traffic": DMMA_N32768_DRAM_BYTES if (world == 1 and N == 32768 and st.steps == 1) else None,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 8 K-chunk launches (kc = 4096) of this GEMM, "
                                      "profiles/ncu_dmma_n32768_dram_r02.csv; algorithmic bytes of the chunked GEMM: 8 x (A chunk 1.07 + B chunk 1.07 "
                                      "+ C read+write 17.2 GB) = 155 GB (a single launch in round 1 moved 1522 GB)",
                    "kernel": "phpc::dmma_gemm_kernel (FP64 DMMA.8x8x4, TMA + mbarrier pipeline)",
                    "flops_per_launch": 2.0 * m_blk * n_blk * k_per_gemm, "kernel_ms": gemm_ms, "peak_source": fp64_src}
        return {"tflops": flops / (ms * 1e-3) / 1e12, "ms": ms, "clocks": clocks, "stats": st, "roofline": roof,
                "exposed": max_over_ranks(st.exposed_ms / st.total_ms if st.total_ms > 0 else 0.0)}

    prim = measure(primary, args.warmup, args.steps, True)

    def throttled(c):
        """A number taken under a hardware / thermal slowdown, or with SM clocks far below max for no stated reason (a leftover
        clock lock), is not kept: measure once more.  sw_power_cap is normal for a dense tensor-core kernel on a 1 kW part."""
        if not c or c.get("sm_mhz") is None:
            return False
        bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c.get("reasons") or [])
        stuck = c.get("sm_max_mhz") and c["sm_mhz"] < 0.6 * c["sm_max_mhz"] and not c.get("reasons")
        return bool(bad or stuck)

    if max_over_ranks(1.0 if (rank == 0 and throttled(prim["clocks"])) else 0.0) > 0:
        first_clocks = prim["clocks"]
        prim = measure(primary, 1, args.steps, True)
        if prim["clocks"] is not None:
            prim["clocks"]["remeasured_after"] = first_clocks
    value, ms_per_step, clocks, st = prim["tflops"], prim["ms"], prim["clocks"], prim["stats"]
    exposed_frac = prim["exposed"]

    # ---------------- self-verification of `value`: every rank checks sampled elements of ITS C block ----------------
    # (the device blocks hold passes x A*B after the warm-up and timed steps; the driver's pytest box has one GPU, so this is
    # the multi-GPU parity check that runs with every bench line)
    m_blk, n_blk = s.block
    pi, pj = s.coords
    vpos = sample_positions(pi * m_blk, m_blk, pj * n_blk, n_blk, 4242 + rank, 8)
    got = {}
    for i in sorted({i for i, _ in vpos}):
        row = s.read_c_block(i - pi * m_blk, 0, 1, n_blk)[0]
        for (i2, j) in vpos:
            if i2 == i:
                got[(i, j)] = row[j - pj * n_blk]
    v_ok, v_worst = sampled_check(capi, L, N, vpos, got, passes_on_c[0])
    value_verified = max_over_ranks(0.0 if v_ok else 1.0) == 0.0
    value_verify = {"elements_checked_per_rank": len(vpos), "ranks": world, "worst_error_over_bound": max_over_ranks(v_worst),
                    "passes_accumulated": passes_on_c[0],
                    "how": "after the timed steps every rank reads sampled elements of its device C block (first/last row and column of the block + "
                           "random ones) and compares them with passes x the exactly summed FP64 dot product of the regenerated row of A and "
                           "column of B; bound 4*sqrt(N)*2^-53*sum|a||b| per pass (the parity tests' bound)"}
    launches_per_step = st.launches
    bytes_rx = st.bytes_received
    steps_per_summa = st.steps
    kc_used = getattr(s, "kc", int(N / st.steps))
    other = None
    if not args.no_secondary:
        sec = measure(secondary, 1, max(1, min(2, args.steps)), False)
        other = {"value": sec["tflops"], "unit": UNIT, "ms_per_step": sec["ms"], "exposed_frac": sec["exposed"],
                 "kernels_per_step": sec["stats"].launches, "roofline": sec["roofline"]}
    cublas = None
    if not args.no_secondary:
        # the strongest on-box library baseline: the same device-resident SUMMA with cuBLAS Dgemm as the local GEMM
        cb = measure(capi.BACKEND_CUBLAS, 1, 1, False)
        cublas = {"value": cb["tflops"], "unit": UNIT, "ms_per_step": cb["ms"], "what": "same SUMMA loop, local GEMM = cublasDgemm (library call, not the product)"}
    s.destroy()

    # ---------------- the north star's transport, measured with the same kernel: ncclBroadcast on row / column communicators ----------------
    nccl = None
    if world > 1 and not args.no_secondary:
        saved = os.environ.get("PHPC_PANEL")
        os.environ["PHPC_PANEL"] = "nccl"
        try:
            s2 = capi.Summa(comm, N, args.kc)
            s2.fill(capi.FILL_SEEDED)
            nb = measure(primary, 1, max(1, min(2, args.steps)), False, s=s2)
            nccl = {"value": nb["tflops"], "unit": UNIT, "ms_per_step": nb["ms"], "exposed_frac": nb["exposed"],
                    "what": "same device-resident SUMMA and local GEMM, panels moved by one ncclGroup of ncclBroadcast on the row and the column "
                            "communicator (ncclCommSplit) per K chunk instead of the default copy-engine pulls over CUDA IPC (PHPC_PANEL=nccl)"}
            s2.destroy()
        finally:
            if saved is None:
                os.environ.pop("PHPC_PANEL", None)
            else:
                os.environ["PHPC_PANEL"] = saved

    # ---------------- alternative schedule: stationary C, all panel transfers issued up front ----------------
    prefetch_all = None
    if world > 1 and not args.no_secondary:
        saved = os.environ.get("PHPC_SCHEDULE")
        os.environ["PHPC_SCHEDULE"] = "prefetch-all"
        try:
            s3 = capi.Summa(comm, N, args.kc)
            s3.fill(capi.FILL_SEEDED)
            pa = measure(primary, 1, max(1, min(2, args.steps)), False, s=s3)
            prefetch_all = {"value": pa["tflops"], "unit": UNIT, "ms_per_step": pa["ms"], "exposed_frac": pa["exposed"],
                            "what": "PHPC_SCHEDULE=prefetch-all: the receive ring holds every K chunk, all copy-engine pulls are issued at the start "
                                    "(all-gather of the panels), GEMMs consume chunks as they land; default = 3-slot ring"}
            s3.destroy()
        finally:
            if saved is None:
                os.environ.pop("PHPC_SCHEDULE", None)
            else:
                os.environ["PHPC_SCHEDULE"] = saved

    # ---------------- e2e through the reference-facing C-ABI on host matrices ----------------
    e2e = None
    if not args.no_e2e:
        if world == 1:
            e2e = e2e_in_child(args)  # own process: a failure of the host path cannot take the bench line with it
        else:
            e2e = e2e_measure(args, capi, L, comm, dims, rank, world, barrier, max_over_ranks)

    # ---------------- CPU baseline beside it (rank 0, single GPU run only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v0, s0, kind = run_reference_cpu(CPU_SAMPLE_N, "O0")
        cpu = {"value": v0, "unit": UNIT, "cores": 1, "kind": kind, "host_cores": os.cpu_count(),
               "sample": f"iterative.c N={CPU_SAMPLE_N} (reference fill), 1 thread, reference flags gcc -Wall (no -O): {s0:.3f} s"}
        if _iterative_binary("O3"):
            v3, s3, _ = run_reference_cpu(CPU_SAMPLE_N, "O3")
            cpu["value_O3"] = v3
            cpu["sample"] += f"; same source with -O3: {s3:.3f} s"
        cpu["reference_summa_all_cores"] = run_reference_summa_all_cores()

    ref_cuda = None
    barrier()
    if rank == 0 and world > 1 and not args.no_refcuda:
        # the reference's own CUDA + MPI program on the same number of ranks = GPUs, sane launch grid
        try:
            ref_cuda = [run_reference_cuda_build(n=8192, gw=148, gh=4, ranks=world, timeout=420)]
        except Exception as e:
            ref_cuda = [{"unavailable": f"{type(e).__name__}: {e}"}]
    if rank == 0 and world == 1 and not args.no_refcuda:
        # the reference's build with (a) its default launch, one 32x32 CTA (what its own CSV sweeps use),
        # and (b) a sane launch, 148x4 CTAs
        ref_cuda = []
        for n_ref, gw, gh in ((4096, 1, 1), (8192, 148, 4)):
            try:
                ref_cuda.append(run_reference_cuda_build(n=n_ref, gw=gw, gh=gh))
            except Exception as e:  # a reported baseline must never cost the bench line
                ref_cuda.append({"unavailable": f"{type(e).__name__}: {e}"})

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"SUMMA C+=A*B, N={N}, FP64 in/out, splitmix64-seeded uniform(-1,1) A/B generated in HBM, owned blocks device-resident",
                       "local_gemm": names[primary],
                       "arithmetic": (f"emulated FP64: rebuilt from {slices} balanced base-256 digits per operand (54 bits below each row / column maximum "
                                      f"of a K chunk of {oz['k_chunk']}), {pairs} s8 x s8 -> s32 products on tcgen05, exact FP64 recombination; normwise error "
                                      "model (rel. Frobenius difference to native FP64 ~1e-15); chunks with Inf/NaN, near-range exponents or > 2^40 "
                                      "spread inside a row / column run on the native-FP64 DMMA kernel")
                       if primary == capi.BACKEND_OZAKI else "native FP64 DMMA",
                       "N": N, "process_grid": f"{dims[0]}x{dims[1]}", "k_chunk": kc_used, "k_chunk_first": getattr(s, "kc_first", 0),
                       "summa_steps": steps_per_summa,
                       "cache": "inputs_larger_than_l2 (operands are GiBs; L2 is 126 MB)", "exposed_broadcast_frac": exposed_frac,
                       "nvlink_bytes_received_rank0_per_step": bytes_rx},
            "value_verified": value_verified,
            "value_verify": value_verify,
            "clocks": clocks,
            "e2e": e2e,
            "e2e_steps_averaged": (e2e or {}).get("steps"),
            "gpu_launches": launches_per_step * args.steps,
            "roofline": prim["roofline"],
            "cpu_baseline": cpu,
            "reference_cuda_build": ref_cuda,
            "local_gemm": names[primary],
            names[secondary]: other,
            "cublas_dgemm": cublas,
            "nccl_broadcast_transport": nccl,
            "schedule_prefetch_all": prefetch_all,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        if rank == 0 and seg:
            try:
                os.unlink(seg)
            except OSError:
                pass
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("PHPC_BENCH_N", "32768")))
    ap.add_argument("--kc", type=int, default=0, help="K chunk (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-child", action="store_true", help="internal: run only the single-GPU e2e leg and print it")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-refcuda", action="store_true", help="skip the run of the reference's own CUDA build (oracle/_ref/ref_main.out)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the second local-GEMM kernel")
    args = ap.parse_args()
    if args.e2e_child:
        e2e_child(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        product_arm(args)


if __name__ == "__main__":
    main()
