"""The executor of the band-pipelined host call (csrc/host_band_exec.h, phpc::band_execute) is written against an abstract
stream backend; the CUDA build plugs in streams, events and the tensor-core GEMM (csrc/phpc_summa.cu), this test plugs in
deferred queues on the CPU (tests/csrc/band_exec_test.cpp) and drains them in random interleavings.  The lines that
compute every source/destination window, pitch and dependency are therefore THE SAME lines the GPU path runs; host
buffers that stand in for HBM start as NaN, so a wrong window or a missing dependency cannot go unnoticed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bandexec") / "libband_exec_test.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "hpc_multigpu_matrixmult_b200", "mpi_shim"), "-o", out,
                    os.path.join(ROOT, "tests", "csrc", "band_exec_test.cpp")], check=True, capture_output=True)
    L = ctypes.CDLL(out)
    L.band_exec_sim.argtypes = [ctypes.c_int] * 4 + [dp, dp, dp, ctypes.c_uint, ctypes.c_int]
    L.band_exec_sim.restype = ctypes.c_int
    return L


@pytest.mark.parametrize("N,kc,bands,align", [(37, 10, 3, 1), (64, 16, 4, 8), (50, 50, 2, 1), (33, 7, 8, 4), (20, 3, 1, 1), (96, 32, 5, 16)])
def test_executor_computes_c_plus_ab_under_any_interleaving(sim, N, kc, bands, align):
    rng = np.random.default_rng(N * 1000 + kc)
    A = rng.integers(-9, 10, (N, N)).astype(np.float64)
    B = rng.integers(-9, 10, (N, N)).astype(np.float64)
    C0 = rng.integers(-9, 10, (N, N)).astype(np.float64)
    want = C0 + A @ B
    nsteps = -(-N // kc)
    nbands, left = 0, N  # bands of 1/2, 1/4, ... of what is left, rounded up to `align`; the last one takes the rest
    while left > 0:
        nbands += 1
        rows = left if nbands == bands else min(left, -(-(-(-left // 2)) // align) * align)
        left -= rows
    for order, seed in [(0, 0)] + [(1, s) for s in range(12)]:
        C = C0.copy()
        gemms = sim.band_exec_sim(N, kc, bands, align, A.ctypes.data_as(dp), B.ctypes.data_as(dp), C.ctypes.data_as(dp), seed, order)
        assert gemms == nbands * nsteps, f"order {order} seed {seed}: {gemms}"
        assert np.array_equal(C, want), f"order {order} seed {seed}"  # integers: exact, and NaN-free


def test_second_pass_accumulates_like_the_reference_main(sim):
    """reference main.c runs two passes on the same C (src/main.c:94,106): C += A*B twice."""
    N = 40
    rng = np.random.default_rng(5)
    A = rng.integers(-5, 6, (N, N)).astype(np.float64)
    B = rng.integers(-5, 6, (N, N)).astype(np.float64)
    C = np.zeros((N, N))
    for seed in (1, 2):
        assert sim.band_exec_sim(N, 16, 3, 1, A.ctypes.data_as(dp), B.ctypes.data_as(dp), C.ctypes.data_as(dp), seed, 1) > 0
    assert np.array_equal(C, 2 * (A @ B))
