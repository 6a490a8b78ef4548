"""
hpc_multigpu_matrixmult_b200 — B200-native SUMMA GEMM behind the C entry points of
Redy1908/HPC-MultiGPU-MatrixMult (phpc_gemm.cuh, phpc_summa.h, main.out CLI).

  csrc/      CUDA kernels (dmma_gemm.cuh), the C-ABI (phpc_core.cu, phpc_summa.cu),
             the drop-in driver main.c and utils.c
  mpi_shim/  single-node MPI subset + mpirun launcher for boxes without MPI
  lib/       built shared libraries (git-ignored, shipped to the GPU box)
  capi.py    ctypes view of the C-ABI for tests/ and bench.py

Importing the package never touches the oracle and never falls back to CPU math.
"""
import os
import subprocess

from . import capi  # noqa: F401

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PACKAGE_DIR)


def build(verbose=False):
    """Compile the CUDA library for sm_100a, the MPI shim, main.out and mpirun (in-tree)."""
    res = subprocess.run(["make", "-C", PACKAGE_DIR, "all"], capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("build failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return capi.load()
