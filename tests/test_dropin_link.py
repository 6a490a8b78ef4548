"""Drop-in link test: the reference's UNMODIFIED driver (src/main.c + src/utils.c, and for recipe (a) also src/phpc_summa.c),
compiled in the build container from the sources where they lie (oracle/Makefile `ref` -> oracle/_ref/dropin_{a,b}.out) and
linked against libphpc_b200.so + the MPI shim exactly as INTEGRATION.md section 2 says, run on a B200:

  dropin_b.out  recipe (b): the reference's main() calls the LIBRARY's phpc_gemm_summa_cuda / phpc_gemm_summa_cublas
  dropin_a.out  recipe (a): the reference's own host-memory SUMMA (phpc_summa.c, MPI_Bcast through the shim) calls the
                library's phpc_gemm_cuda / phpc_gemm_cublas as its gemm_t plugin (src/phpc_summa.c:7,93)

The reference never outputs C, so both binaries carry oracle/dropin_wrap.c in front of the two SUMMA calls of main.c
(linker --wrap: no reference source line changes): zero C, call through, dump rank 0's C.  Checked: exit status, the CSV
file name and 9-column record of src/main.c:75 / src/utils.c:26-27, and C bit for bit against the closed form of the
reference's fill A[i] = B[i] = i at N = 1024 (every summation order is exact there)."""
import glob
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _run(recipe, ranks, tmp_path, n=1024, env=None):
    exe = os.path.join(REF, f"dropin_{recipe}.out")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} was not built (needs /root/reference in the build container)")
    (tmp_path / "csv").mkdir(exist_ok=True)
    e = dict(os.environ, PHPC_DROPIN_DUMP=str(tmp_path / "c"))
    e.update(env or {})
    cmd = [os.path.join(ROOT, "bin", "mpirun"), "--oversubscribe", "-n", str(ranks), exe, str(n), "32", "1", "1", "dropin"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=tmp_path, env=e)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    files = glob.glob(str(tmp_path / "csv" / f"dropin_N{n}_T{ranks}_G*_TW32_GW1_GH1.csv"))
    assert len(files) == 1, os.listdir(tmp_path / "csv")
    rec = open(files[0]).read().strip().split(",")
    assert len(rec) == 9 and rec[0] == str(n) and rec[1] == str(ranks) and rec[3] == "1" and rec[4] == "1024"
    assert float(rec[6]) > 0 and float(rec[8]) > 0  # cuda_time, cublas_time (wall seconds of the two passes)
    out = {}
    for which in ("cuda", "cublas"):
        out[which] = np.fromfile(str(tmp_path / f"c.{which}"), dtype=np.float64).reshape(n, n)
    return out, rec


@pytest.mark.parametrize("recipe", ["b", "a"])
def test_reference_driver_links_and_runs_single_rank(gpu, oracle, tmp_path, recipe):
    out, rec = _run(recipe, 1, tmp_path)
    exact = oracle.index_fill_exact(1024)
    assert np.array_equal(out["cuda"], exact)
    assert np.array_equal(out["cublas"], exact)
    if recipe == "b":
        assert float(rec[7]) > 0  # cuda_gpu_time: device seconds of the local GEMMs


@pytest.mark.multigpu
@pytest.mark.parametrize("recipe,ranks", [("b", 2), ("a", 2), ("b", 4), ("a", 4)])
def test_reference_driver_links_and_runs_multi_rank(gpu, oracle, tmp_path, recipe, ranks):
    """MPI_Dims_create grids of the reference (2 -> 2x1, 4 -> 2x2), one GPU per rank."""
    have = gpu.phpc_b200_device_count()
    if have < ranks:
        pytest.skip(f"needs {ranks} GPUs, {have} visible")
    # recipe (a): the reference hands every rank ALL visible GPUs (src/main.c:54-56); give each rank its own, as its SLURM
    # deployment does (--gpus-per-task), through the launcher's per-rank CUDA_VISIBLE_DEVICES
    env = {"PHPC_GPU_POLICY": "visible", "PHPC_GPUS": str(ranks)} if recipe == "a" else {}
    out, _ = _run(recipe, ranks, tmp_path, env=env)
    exact = oracle.index_fill_exact(1024)
    assert np.array_equal(out["cuda"], exact)
    assert np.array_equal(out["cublas"], exact)
