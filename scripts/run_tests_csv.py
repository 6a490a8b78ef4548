#!/usr/bin/env python
"""
run_tests_csv.py — local runner for run-configuration CSVs in the reference's format
(header `matrix_size,n_proc,n_gpu,tile_width,grid_width,grid_height`, reference tests/*.csv).

Stands in for the reference's SLURM submitter (scripts/run.sh:34-86) and local sweep
(scripts/tests.sh:46-93) on a box without SLURM/MPI/bc: one `bin/mpirun -n <n_proc> bin/main.out
<matrix_size> <tile_width> <grid_width> <grid_height> <test_name>` per row, then merges the one-line
CSV records into `csv/<test_name>.csv` with the reference's merged schema (scripts/tests.sh:17)
    matrix_size,n_proc,n_gpu,n_block,n_thread_per_block,n_thread,time,time_kernel,time_cublas,
    speedup,speedup_kernel,speedup_cublas,efficiency,efficiency_kernel
plus tflops, tflops_cublas.  Speedups are against `csv/iterative.csv` ("n,seconds" lines of
iterative.out, measured for N <= --cpu-max and extrapolated with N^3 beyond, flagged in the log).

    python scripts/run_tests_csv.py tests/configs/b200_configs.csv --name b200 [--dry-run] [--pgrid 2x4]
"""
import argparse
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = ["matrix_size", "n_proc", "n_gpu", "tile_width", "grid_width", "grid_height"]
PGRID = {2: "1x2", 8: "2x4"}  # BASELINE.json's orientations (MPI_Dims_create would give 2x1 / 4x2)


def read_rows(path):
    with open(path) as f:
        rows = list(csv.DictReader(line for line in f if line.strip() and not line.lstrip().startswith("#")))
    if not rows or any(h not in rows[0] for h in HEADER):
        raise SystemExit(f"{path}: expected header {','.join(HEADER)}")
    return [{h: int(r[h]) for h in HEADER} for r in rows]


def command(row, name):
    exe = os.path.join(ROOT, "bin", "main.out")
    args = [str(row["matrix_size"]), str(row["tile_width"]), str(row["grid_width"]), str(row["grid_height"]), name]
    if row["n_proc"] == 1:
        return [exe] + args
    return [os.path.join(ROOT, "bin", "mpirun"), "--oversubscribe", "-n", str(row["n_proc"]), exe] + args


def iterative_seconds(n, cache, cpu_max):
    """Seconds of the CPU baseline for size n: measured up to cpu_max, N^3-extrapolated beyond."""
    exe = os.path.join(ROOT, "oracle", "_ref", "iterative_O0.out")
    base = min(n, cpu_max)
    if base not in cache:
        if not os.path.exists(exe):
            return None, False
        out = subprocess.run([exe, str(base)], check=True, capture_output=True, text=True).stdout.strip()
        cache[base] = float(out.split(",")[1])
    return cache[base] * (n / base) ** 3, n > base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--name", default=None)
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--cpu-max", type=int, default=1024)
    ap.add_argument("--pgrid", default=None, help="force PHPC_PGRID=RxC for every multi-rank row")
    ap.add_argument("--verify", action="store_true", help="PHPC_VERIFY=1: main.out checks sampled elements of C")
    args = ap.parse_args()
    name = args.name or os.path.splitext(os.path.basename(args.config))[0]
    rows = read_rows(args.config)
    os.makedirs(os.path.join(ROOT, "csv"), exist_ok=True)
    cache, merged = {}, []
    for row in rows:
        cmd = command(row, name)
        env = dict(os.environ)
        grid = args.pgrid or PGRID.get(row["n_proc"])
        if grid and row["n_proc"] > 1:
            env["PHPC_PGRID"] = grid
        if args.verify:
            env["PHPC_VERIFY"] = "1"
        print(("PHPC_PGRID=%s " % grid if grid and row["n_proc"] > 1 else "") + " ".join(cmd), flush=True)
        if args.dry_run:
            continue
        res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True)
        if res.returncode != 0:
            print(f"  FAILED rc={res.returncode}: {res.stderr.strip()[-300:]}", flush=True)
            continue
        pattern = os.path.join(ROOT, "csv", f"{name}_N{row['matrix_size']}_T{row['n_proc']}_G*_TW{row['tile_width']}_GW{row['grid_width']}_GH{row['grid_height']}.csv")
        rec = open(sorted(glob.glob(pattern))[-1]).read().strip().split(",")
        n, procs = int(rec[0]), int(rec[1])
        t, tk, tc = float(rec[6]), float(rec[7]), float(rec[8])
        base, extrapolated = iterative_seconds(n, cache, args.cpu_max)
        threads = int(rec[5])
        extra = [""] * 5
        if base:
            extra = [f"{base / t:.2f}", f"{base / tk:.2f}" if tk > 0 else "", f"{base / tc:.2f}", f"{base / t / max(threads, 1):.6f}",
                     f"{base / tk / max(threads, 1):.6f}" if tk > 0 else ""]
            if extrapolated:
                print(f"  CPU baseline for N={n} extrapolated from N={min(n, args.cpu_max)} with N^3", flush=True)
        flops = 2.0 * n ** 3
        merged.append(rec + extra + [f"{flops / t / 1e12:.3f}", f"{flops / tc / 1e12:.3f}"])
        side = pattern.replace("*", rec[2]) + ".json"
        if os.path.exists(side):
            print("  " + json.dumps(json.load(open(side))), flush=True)
    if not args.dry_run:
        out = os.path.join(ROOT, "csv", f"{name}.csv")
        with open(out, "w") as f:
            f.write("matrix_size,n_proc,n_gpu,n_block,n_thread_per_block,n_thread,time,time_kernel,time_cublas,"
                    "speedup,speedup_kernel,speedup_cublas,efficiency,efficiency_kernel,tflops,tflops_cublas\n")
            for m in merged:
                f.write(",".join(m) + "\n")
        print(f"wrote {out} ({len(merged)} rows)")


if __name__ == "__main__":
    main()
