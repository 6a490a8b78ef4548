"""the CSV harness.  scripts/run_tests_csv.py reads run configurations in the reference's tests/*.csv format
(header matrix_size,n_proc,n_gpu,tile_width,grid_width,grid_height), launches `main.out` per row as the reference's
scripts/tests.sh:7-8 does, and merges the one-line records into the reference's merged schema (scripts/tests.sh:17) with
speedup / efficiency against iterative.c; --plots writes the three figures of scripts/plots.py:174-237 as SVG."""
import csv
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MERGED = ("matrix_size,n_proc,n_gpu,n_block,n_thread_per_block,n_thread,time,time_kernel,time_cublas,speedup,speedup_kernel,"
          "speedup_cublas,efficiency,efficiency_kernel,tflops,tflops_cublas").split(",")


def test_svg_plots_are_written_without_matplotlib(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import run_tests_csv as r

    merged = [["256", "1", "1", "1", "1024", "1024", "0.01", "0.001", "0.002", "3.0", "30.0", "15.0", "0.003", "0.03", "0.1", "0.2"],
              ["512", "1", "1", "1", "1024", "1024", "0.02", "0.004", "0.006", "12.0", "60.0", "40.0", "0.012", "0.06", "0.3", "0.4"]]
    files = r.write_plots(str(tmp_path), "t", merged)
    assert len(files) == 3
    for f in files:
        text = open(f).read()
        assert text.startswith("<svg") and "polyline" in text and text.rstrip().endswith("</svg>")


@pytest.mark.gpu
def test_runner_executes_the_rows_and_merges_the_reference_schema(gpu, tmp_path):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_tests_csv.py"), os.path.join(ROOT, "tests", "configs", "smoke_configs.csv"),
                          "--name", "smoke", "--verify", "--workdir", str(tmp_path), "--plots", "--cpu-max", "512"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "FAILED" not in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    rows = list(csv.reader(open(tmp_path / "csv" / "smoke.csv")))
    assert rows[0] == MERGED and len(rows) == 4
    for row, n in zip(rows[1:], (256, 512, 1024)):
        assert int(row[0]) == n and row[1] == "1" and len(row) == len(MERGED)
        assert float(row[6]) > 0 and float(row[7]) > 0 and float(row[8]) > 0  # time, time_kernel, time_cublas
        assert float(row[9]) > 0 and float(row[14]) > 0                          # speedup over iterative.c, TFLOP/s
    assert "extrapolated" in res.stdout  # N = 1024 > --cpu-max: the CPU baseline is N^3-extrapolated and says so
    for suffix in ("", "_speedup", "_efficiency"):
        assert (tmp_path / "plots" / f"smoke{suffix}.svg").exists()
