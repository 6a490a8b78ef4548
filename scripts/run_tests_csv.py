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

plus tflops, tflops_cublas, and with --plots the three figures of the reference's scripts/plots.py:174-237 (time, speedup,
efficiency against matrix size or rank count) as dependency-free SVG files under <workdir>/plots/ (matplotlib is not in the image).

    python scripts/run_tests_csv.py tests/configs/b200_configs.csv --name b200 [--dry-run] [--pgrid 2x4] [--workdir DIR] [--plots]
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


def svg_plot(path, title, ylabel, xlabel, xs, series, logy=False):
    """Minimal line chart: series = [(label, [y...])], one point per x (categorical x axis, as the reference's plots are)."""
    import math

    W, H, L, R, T, B = 640, 400, 70, 150, 40, 50
    vals = [v for _, ys in series for v in ys if v is not None and v > 0]
    if not vals:
        return
    f = (lambda v: math.log10(v)) if logy else (lambda v: v)
    lo, hi = (f(min(vals)), f(max(vals))) if logy else (0.0, max(vals))
    if hi <= lo:
        hi = lo + 1.0
    X = lambda i: L + (W - L - R) * (i / max(len(xs) - 1, 1))
    Y = lambda v: H - B - (H - T - B) * ((f(v) - lo) / (hi - lo))
    colors = ["#1f77b4", "#ff7f0e", "#2ca02c", "#d62728"]
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif" font-size="12">',
           f'<rect width="{W}" height="{H}" fill="white"/><text x="{W / 2}" y="20" text-anchor="middle" font-size="14">{title}</text>',
           f'<line x1="{L}" y1="{H - B}" x2="{W - R}" y2="{H - B}" stroke="black"/><line x1="{L}" y1="{T}" x2="{L}" y2="{H - B}" stroke="black"/>',
           f'<text x="{(L + W - R) / 2}" y="{H - 10}" text-anchor="middle">{xlabel}</text>',
           f'<text x="15" y="{H / 2}" text-anchor="middle" transform="rotate(-90 15 {H / 2})">{ylabel}{" (log)" if logy else ""}</text>']
    for i, x in enumerate(xs):
        out.append(f'<text x="{X(i)}" y="{H - B + 16}" text-anchor="middle">{x}</text>')
    for t in range(5):
        v = lo + (hi - lo) * t / 4
        label = f"{10 ** v:.3g}" if logy else f"{v:.3g}"
        out.append(f'<text x="{L - 6}" y="{H - B - (H - T - B) * t / 4 + 4}" text-anchor="end">{label}</text>')
    for k, (label, ys) in enumerate(series):
        pts = [(X(i), Y(v)) for i, v in enumerate(ys) if v is not None and v > 0]
        if not pts:
            continue
        c = colors[k % len(colors)]
        out.append(f'<polyline fill="none" stroke="{c}" stroke-width="2" points="{" ".join(f"{a:.1f},{b:.1f}" for a, b in pts)}"/>')
        out += [f'<circle cx="{a:.1f}" cy="{b:.1f}" r="3" fill="{c}"/>' for a, b in pts]
        out.append(f'<text x="{W - R + 10}" y="{T + 18 * k + 10}" fill="{c}">{label}</text>')
    out.append("</svg>")
    with open(path, "w") as fh:
        fh.write("\n".join(out))


def write_plots(workdir, name, merged):
    """time / speedup / efficiency against the column that varies (matrix size, else rank count): reference scripts/plots.py:174-237."""
    os.makedirs(os.path.join(workdir, "plots"), exist_ok=True)
    col = 0 if len({m[0] for m in merged}) > 1 else 1
    xs = [m[col] for m in merged]
    num = lambda m, i: float(m[i]) if m[i] not in ("", None) else None
    base = os.path.join(workdir, "plots", name)
    xlabel = "matrix size N" if col == 0 else "MPI ranks (GPUs)"
    svg_plot(base + ".svg", f"{name}: time", "seconds", xlabel, xs,
             [("CUDA (SUMMA call)", [num(m, 6) for m in merged]), ("CUDA kernels", [num(m, 7) for m in merged]), ("cuBLAS", [num(m, 8) for m in merged])], logy=True)
    svg_plot(base + "_speedup.svg", f"{name}: speedup over iterative.c (1 core)", "speedup", xlabel, xs,
             [("CUDA", [num(m, 9) for m in merged]), ("CUDA kernels", [num(m, 10) for m in merged]), ("cuBLAS", [num(m, 11) for m in merged])], logy=True)
    svg_plot(base + "_efficiency.svg", f"{name}: efficiency (speedup / threads)", "efficiency", xlabel, xs,
             [("CUDA", [num(m, 12) for m in merged]), ("CUDA kernels", [num(m, 13) for m in merged])], logy=True)
    return [base + ".svg", base + "_speedup.svg", base + "_efficiency.svg"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--name", default=None)
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--cpu-max", type=int, default=1024)
    ap.add_argument("--pgrid", default=None, help="force PHPC_PGRID=RxC for every multi-rank row")
    ap.add_argument("--verify", action="store_true", help="PHPC_VERIFY=1: main.out checks sampled elements of C")
    ap.add_argument("--workdir", default=ROOT, help="directory the runs start in (csv/ and plots/ are created there)")
    ap.add_argument("--plots", action="store_true", help="write time / speedup / efficiency SVG plots of the merged CSV")
    args = ap.parse_args()
    work = os.path.abspath(args.workdir)
    name = args.name or os.path.splitext(os.path.basename(args.config))[0]
    rows = read_rows(args.config)
    os.makedirs(os.path.join(work, "csv"), exist_ok=True)
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
        res = subprocess.run(cmd, cwd=work, env=env, capture_output=True, text=True)
        if res.returncode != 0:
            print(f"  FAILED rc={res.returncode}: {res.stderr.strip()[-300:]}", flush=True)
            continue
        pattern = os.path.join(work, "csv", f"{name}_N{row['matrix_size']}_T{row['n_proc']}_G*_TW{row['tile_width']}_GW{row['grid_width']}_GH{row['grid_height']}.csv")
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
        out = os.path.join(work, "csv", f"{name}.csv")
        with open(out, "w") as f:
            f.write("matrix_size,n_proc,n_gpu,n_block,n_thread_per_block,n_thread,time,time_kernel,time_cublas,"
                    "speedup,speedup_kernel,speedup_cublas,efficiency,efficiency_kernel,tflops,tflops_cublas\n")
            for m in merged:
                f.write(",".join(m) + "\n")
        print(f"wrote {out} ({len(merged)} rows)")
        if args.plots and merged:
            for path in write_plots(work, name, merged):
                print(f"wrote {path}")


if __name__ == "__main__":
    main()
