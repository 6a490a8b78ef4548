"""N>1 host logic on CPU: gloo world_size 2 and 4, MPI shim under bin/mpirun, main.out's
argument checks.  No CUDA call is made (the CUDA path has no CPU fallback)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket

    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _torchrun(nproc, args, port=None):
    port = port or _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_gloo_worker.py")] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)


@pytest.mark.parametrize("grid,nproc", [("1x2", 2), ("2x1", 2), ("2x2", 4)])
def test_gloo_world_walks_the_schedule(built, grid, nproc):
    res = _torchrun(nproc, [grid, "48", "5"])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_main_out_usage_and_divisibility_messages(built):
    """Same abort conditions and messages as reference src/main.c:26-31,47-52."""
    exe = os.path.join(ROOT, "bin", "main.out")
    res = subprocess.run([exe, "128"], capture_output=True, text=True, timeout=60)
    assert res.returncode != 0
    assert "Usage: " in res.stderr and "<matrix_size> <tile_width> <grid_width> <grid_height> <test_name>" in res.stderr
    mpirun = os.path.join(ROOT, "bin", "mpirun")
    res = subprocess.run([mpirun, "-n", "4", exe, "127", "32", "1", "1", "t"], capture_output=True, text=True, timeout=60)
    assert res.returncode != 0
    assert "Error: Matrix size N (127) must be divisible by process grid dimensions (2 x 2)." in res.stderr


def test_mpirun_propagates_failure_and_rank_env(built, tmp_path):
    mpirun = os.path.join(ROOT, "bin", "mpirun")
    out = subprocess.run([mpirun, "--oversubscribe", "-n", "3", "sh", "-c", "echo $PHPC_MPI_RANK/$PHPC_MPI_SIZE"], capture_output=True,
                         text=True, timeout=60)
    assert out.returncode == 0
    assert sorted(out.stdout.split()) == ["0/3", "1/3", "2/3"]
    bad = subprocess.run([mpirun, "-n", "2", "sh", "-c", "exit 7"], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 7


@pytest.mark.parametrize("ranks", [1, 2, 4, 8])
def test_shim_collectives_under_stress(built, tmp_path, ranks):
    """The 19-function MPI subset under load, with more ranks than some boxes have cores."""
    exe = str(tmp_path / "shim_stress")
    shim = os.path.join(ROOT, "hpc_multigpu_matrixmult_b200", "mpi_shim")
    lib = os.path.join(ROOT, "hpc_multigpu_matrixmult_b200", "lib")
    subprocess.run(["gcc", "-O2", "-I", shim, "-o", exe, os.path.join(ROOT, "tests", "csrc", "shim_stress.c"), "-L", lib, "-lphpcmpi",
                    f"-Wl,-rpath,{lib}"], check=True)
    res = subprocess.run([os.path.join(ROOT, "bin", "mpirun"), "--oversubscribe", "-n", str(ranks), exe, "1500"], capture_output=True, text=True,
                         timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr


def test_no_cpu_fallback_without_a_gpu(built):
    """On a box without a CUDA device the compute entry points abort loudly; nothing computes on the CPU."""
    import ctypes

    from hpc_multigpu_matrixmult_b200 import capi

    if capi.load().phpc_b200_device_count() > 0:
        pytest.skip("a GPU is visible here")
    code = ("import numpy as np, sys; sys.path.insert(0, %r); from hpc_multigpu_matrixmult_b200 import capi; "
            "a = np.ones((4, 4)); c = np.zeros((4, 4)); capi.phpc_gemm_cuda(a, a, c); print('COMPUTED', c.sum())" % ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode != 0
    assert "COMPUTED" not in res.stdout
    assert "no CUDA device visible" in res.stderr and "no CPU fallback" in res.stderr
