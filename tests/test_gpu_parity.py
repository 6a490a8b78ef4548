"""Parity tests proper (-m gpu): the CUDA path, called through the C-ABI with host buffers
exactly as the reference's SUMMA loop calls its gemm_t plugin, against the oracle.

Tolerances.  Bit-exact (0 ulp) for the reference's own input A[i]=B[i]=i while every
partial sum is an integer below 2^53 (N < 1552).  Otherwise relative Frobenius
error <= 1e-14 against the oracle (which sums in the reference's order; the tensor-core
kernel sums in a different order, so equality is not expected) and, per element,
|c - c*| <= 4 * sqrt(k) * 2^-53 * sum_p |a_p||b_p|.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

U = 2.0 ** -53


def _check(oracle, c, want, a, b, tol=1e-14):
    assert oracle.rel_frobenius(c, want) <= tol
    bound = 4.0 * np.sqrt(max(a.shape[1], 1)) * U * (np.abs(a) @ np.abs(b)) + 1e-300
    assert np.all(np.abs(c - want) <= bound + 4 * U * np.abs(want))


SHAPES = [
    (1, 1, 1), (8, 4, 8), (64, 64, 64), (128, 16, 128), (128, 128, 128), (129, 17, 131), (100, 37, 53),
    (300, 215, 170), (257, 511, 255), (512, 1024, 384), (1000, 999, 1001), (130, 2048, 70),
]


@pytest.mark.parametrize("m,k,n", SHAPES)
def test_gemm_cuda_seeded_vs_oracle(gpu, capi, oracle, m, k, n):
    a = oracle.fill(m, k, kind=1, seed=101)
    b = oracle.fill(k, n, kind=1, seed=202)
    c0 = oracle.fill(m, n, kind=1, seed=303)
    c = c0.copy()
    secs = capi.phpc_gemm_cuda(a, b, c)
    assert secs >= 0.0
    _check(oracle, c, oracle.gemm_block(a, b, c0), a, b)


@pytest.mark.parametrize("n", [16, 32, 64, 128, 200, 512, 1024])
def test_gemm_cuda_index_fill_bit_exact(gpu, capi, oracle, n):
    """The reference's input (src/main.c:85-86): exact for any summation order at these sizes."""
    a = oracle.fill(n, n, kind=0)
    c = np.zeros((n, n))
    capi.phpc_gemm_cuda(a, a.copy(), c)
    assert np.array_equal(c, oracle.index_fill_exact(n))


def test_gemm_cuda_interior_pointers_and_leading_dimensions(gpu, capi, oracle):
    """The SUMMA owner passes interior pointers with ld = N (reference src/phpc_summa.c:72-73,82-83)."""
    N = 96
    A = oracle.fill(N, N, kind=1, seed=1)
    B = oracle.fill(N, N, kind=1, seed=2)
    C = oracle.fill(N, N, kind=1, seed=3)
    C0 = C.copy()
    a, b, c = A[24:72, 32:64], B[32:64, 48:96], C[24:72, 48:96]
    capi.phpc_gemm_cuda(a, b, c)
    want = C0.copy()
    want[24:72, 48:96] = oracle.gemm_block(a.copy(), b.copy(), C0[24:72, 48:96].copy())
    assert oracle.rel_frobenius(C, want) <= 1e-14
    mask = np.ones_like(C, dtype=bool)
    mask[24:72, 48:96] = False
    assert np.array_equal(C[mask], C0[mask])  # nothing outside the block is touched


def test_gemm_cuda_empty_and_degenerate(gpu, capi, oracle):
    c = np.ones((4, 6))
    capi.phpc_gemm_cuda(np.zeros((4, 0)), np.zeros((0, 6)), c)  # k = 0: C unchanged
    assert np.array_equal(c, np.ones((4, 6)))
    capi.phpc_gemm_cuda(np.zeros((0, 5)), np.zeros((5, 6)), np.zeros((0, 6)))  # m = 0: no-op


@pytest.mark.parametrize("tile_width,gw,gh", [(1, 1, 1), (16, 1, 1), (32, 2, 2), (32, 4, 4), (64, 148, 4), (256, 7, 3)])
def test_launch_parameters_never_change_the_result(gpu, capi, oracle, tile_width, gw, gh):
    """tile_width / grid_width / grid_height of the CLI and CSV sweeps (tests/*.csv) are accepted."""
    m, k, n = 260, 100, 390
    a = oracle.fill(m, k, kind=1, seed=5)
    b = oracle.fill(k, n, kind=1, seed=6)
    c = np.zeros((m, n))
    capi.phpc_gemm_cuda(a, b, c, 1, gw, gh, tile_width)
    ref = np.zeros((m, n))
    capi.phpc_gemm_cuda(a, b, ref)
    assert np.array_equal(c, ref)
    _check(oracle, c, oracle.gemm_block(a, b), a, b)


def test_gemm_cublas_entry_point(gpu, capi, oracle):
    m, k, n = 200, 150, 100
    a = oracle.fill(m, k, kind=1, seed=7)
    b = oracle.fill(k, n, kind=1, seed=8)
    c0 = oracle.fill(m, n, kind=1, seed=9)
    c = c0.copy()
    t = capi.phpc_gemm_cublas(a, b, c)
    assert t == 0.0  # reference src/phpc_gemm.cu:173
    _check(oracle, c, oracle.gemm_block(a, b, c0), a, b)


@pytest.mark.parametrize("n", [64, 256, 1024])
def test_summa_single_rank_bit_exact_on_reference_fill(gpu, capi, oracle, n):
    """BASELINE config 1 (N=1024, 1 rank) through phpc_gemm_summa_cuda, vs iterative.c's result."""
    comm = capi.cart_create((1, 1))
    A = oracle.fill(n, n, kind=0)
    C = np.zeros((n, n))
    secs = capi.phpc_gemm_summa_cuda(comm, A, A.copy(), C, 1, 1, 1, 32)
    assert secs > 0.0
    assert np.array_equal(C, oracle.index_fill_exact(n))
    if n <= 256:
        assert np.array_equal(C, oracle.gemm_iterative(A, A))


def test_summa_single_rank_matches_golden_and_accumulates(gpu, capi, oracle):
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz"))
    n = 48
    comm = capi.cart_create((1, 1))
    A = oracle.fill(n, n, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(n, n, kind=1, seed=oracle.SEED_B)
    C = np.zeros((n, n))
    capi.phpc_gemm_summa_cuda(comm, A, B, C)
    assert oracle.rel_frobenius(C, g["summa_N48_P1_F1"]) <= 1e-14
    first = C.copy()
    capi.phpc_gemm_summa_cuda(comm, A, B, C)  # C += A*B again (reference main.c runs both passes on one C)
    assert oracle.rel_frobenius(C, 2 * first) <= 1e-15
    Cb = np.zeros((n, n))
    capi.phpc_gemm_summa_cublas(comm, A, B, Cb)
    assert oracle.rel_frobenius(Cb, g["summa_N48_P1_F1"]) <= 1e-14


def test_summa_chunked_k_loop_device_resident(gpu, capi, oracle):
    """K chunking (the multi-buffered loop) must not change the result beyond rounding."""
    n = 384
    comm = capi.cart_create((1, 1))
    A = oracle.fill(n, n, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(n, n, kind=1, seed=oracle.SEED_B)
    want = oracle.gemm_block(A, B)
    for kc in (0, 37, 128):
        s = capi.Summa(comm, n, kc)
        s.fill(capi.FILL_SEEDED)
        st = s.run(capi.BACKEND_DMMA)
        assert st.steps == (1 if kc == 0 else -(-n // kc)) and st.launches == st.steps
        assert oracle.rel_frobenius(s.read_c_block(), want) <= 1e-14
        s.zero_c()
        s.run(capi.BACKEND_CUBLAS)
        assert oracle.rel_frobenius(s.read_c_block(), want) <= 1e-14
        s.destroy()


def test_rectangular_summa_object_single_rank(gpu, capi, oracle):
    """SURVEY 8(f1) on one GPU: C[M x N] += A[M x K] * B[K x N] with M, K, N all different through phpc_summa_create_mkn: blocks
    generated in HBM, blocks uploaded from full host matrices (chunk loop and row bands, nonzero C), every backend."""
    import os

    M, K, N = 700, 1100, 450
    comm = capi.cart_create((1, 1))
    A = oracle.fill(M, K, N=K, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(K, N, N=N, kind=1, seed=oracle.SEED_B)
    want = oracle.gemm_block(A, B)
    s = capi.Summa(comm, N, 300, m=M, k=K)
    assert s.mkn == (M, K, N) and s.block == (M, N)
    s.fill(capi.FILL_SEEDED)
    for backend in (capi.BACKEND_OZAKI, capi.BACKEND_DMMA, capi.BACKEND_CUBLAS):
        s.zero_c()
        st = s.run(backend)
        assert st.steps == 4
        assert oracle.rel_frobenius(s.read_c_block(), want) <= 1e-14
    C0 = oracle.fill(M, N, N=N, kind=1, seed=5)
    saved = os.environ.get("PHPC_HOST_BANDS")
    try:
        for bands in ("1", "3"):
            os.environ["PHPC_HOST_BANDS"] = bands
            C = C0.copy()
            s.run_host(A, B, C)
            assert oracle.rel_frobenius(C, oracle.gemm_block(A, B, C0)) <= 1e-14, bands
    finally:
        if saved is None:
            os.environ.pop("PHPC_HOST_BANDS", None)
        else:
            os.environ["PHPC_HOST_BANDS"] = saved
    s.destroy()
    # the reference's fill on a rectangular problem: A[i][j] = i*K + j, B[i][j] = i*N + j, exact while sums stay below 2^53
    M, K, N = 96, 160, 224
    s = capi.Summa(comm, N, 0, m=M, k=K)
    s.fill(capi.FILL_INDEX)
    s.run()
    Ai = oracle.fill(M, K, N=K, kind=0)
    Bi = oracle.fill(K, N, N=N, kind=0)
    assert np.array_equal(s.read_c_block(), oracle.gemm_block(Ai, Bi))
    s.destroy()


def _device_gemm(capi, lib, n, kind, scale=1.0, cublas=False):
    """C = (scale*A) * B for the N x N synthetic matrices, all in HBM; returns device pointer of C."""
    ld = n
    bytes_ = n * ld * 8
    dA, dB, dC = (lib.phpc_device_malloc(bytes_) for _ in range(3))
    lib.phpc_fill_device(dA, ld, n, n, 0, 0, n, kind, capi.SEED_A, None)
    lib.phpc_fill_device(dB, ld, n, n, 0, 0, n, kind, capi.SEED_B, None)
    lib.phpc_device_memset(dC, 0, bytes_)
    return dA, dB, dC


@pytest.mark.parametrize("kernel", ["tcgen05", "dmma"])
def test_full_size_n16384_properties(gpu, capi, oracle, kernel):
    """BASELINE config 2 (N=16384, one GPU) on BOTH kernels — the tcgen05 (Ozaki) default of the entry points, whose
    launcher cuts K = 16384 into two chunks, and the native-FP64 DMMA kernel: size-independent properties instead of
    a CPU GEMM.
    (a) 256 sampled elements against correctly rounded dot products of regenerated rows/columns;
    (b) against cuBLAS Dgemm on the same device inputs (rel Frobenius on a 512x2048 window);
    (c) linearity: running the GEMM twice into the same C doubles it (C += semantics)."""
    lib = gpu
    n = 16384
    dA, dB, dC = _device_gemm(capi, lib, n, capi.FILL_SEEDED)

    def gemm():
        if kernel == "tcgen05":
            assert lib.phpc_gemm_device_ozaki(dA, n, dB, n, dC, n, n, n, n, None) >= 2  # at least two K chunks
        else:
            assert lib.phpc_gemm_device(dA, n, dB, n, dC, n, n, n, n, 0, None) == 4  # K chunks of 4096

    gemm()
    lib.phpc_device_synchronize()

    def window(ptr, r0, c0, rows, cols):
        return capi.device_window(ptr, n, r0, c0, rows, cols)

    rng = np.random.default_rng(7)
    rows = rng.integers(0, n, 16)
    cols = rng.integers(0, n, 16)
    worst = 0.0
    for r in rows:
        crow = window(dC, int(r), 0, 1, n)[0]
        a_row = oracle.fill(1, n, row0=int(r), col0=0, N=n, kind=1, seed=oracle.SEED_A)[0]
        for c in cols:
            b_col = oracle.fill(n, 1, row0=0, col0=int(c), N=n, kind=1, seed=oracle.SEED_B)[:, 0]
            exact = oracle.dot_exact(a_row, b_col)
            bound = 4.0 * np.sqrt(n) * U * float(np.abs(a_row) @ np.abs(b_col))
            assert abs(crow[c] - exact) <= bound
            worst = max(worst, abs(crow[c] - exact) / bound)
    # (b) cuBLAS on the same inputs
    dC2 = lib.phpc_device_malloc(n * n * 8)
    lib.phpc_device_memset(dC2, 0, n * n * 8)
    lib.phpc_gemm_device_cublas(dA, n, dB, n, dC2, n, n, n, n, None)
    lib.phpc_device_synchronize()
    w1 = window(dC, 4096, 8192, 512, 2048)
    w2 = window(dC2, 4096, 8192, 512, 2048)
    assert oracle.rel_frobenius(w1, w2) <= 1e-14
    # (c) C += : second pass doubles every element up to one rounding per K chunk (x + x itself is exact)
    gemm()
    lib.phpc_device_synchronize()
    w3 = window(dC, 4096, 8192, 512, 2048)
    assert oracle.rel_frobenius(w3, 2 * w1) <= 1e-15
    for p in (dA, dB, dC, dC2):
        lib.phpc_device_free(p)
    print(f"N=16384 [{kernel}] sampled-dot worst error / bound = {worst:.3f}")


@pytest.mark.parametrize("m,k,n", [(260, 20000, 140), (128, 32768, 130), (140, 16385, 257)])
def test_ozaki_multi_k_chunk_launcher_vs_oracle(gpu, capi, oracle, m, k, n):
    """K larger than one int32-exact chunk: the launcher of the tcgen05 kernel walks several K chunks (exponents, digit
    split and MMA kernel per chunk, C += per chunk).  k = 32768 is the shape of the headline 1x1 run; 16385 leaves a
    chunk of a single column.  Against the oracle (reference summation order) and correctly rounded dot products."""
    a = oracle.fill(m, k, kind=1, seed=111)
    b = oracle.fill(k, n, kind=1, seed=222)
    c0 = oracle.fill(m, n, kind=1, seed=333)
    c, launched = _device_gemm_from_numpy(capi, gpu, a, b, c0, "ozaki")
    assert launched >= 16  # at least two chunks
    _check(oracle, c, oracle.gemm_block(a, b, c0), a, b)
    rng = np.random.default_rng(k)
    for i, j in zip(rng.integers(0, m, 24), rng.integers(0, n, 24)):
        exact = oracle.dot_exact(a[i], np.ascontiguousarray(b[:, j])) + c0[i, j]
        assert abs(c[i, j] - exact) <= 4.0 * np.sqrt(k) * U * float(np.abs(a[i]) @ np.abs(b[:, j]))


def test_index_fill_large_n_against_closed_form(gpu, capi, oracle):
    """N=4096 reference fill: values reach 2^48 * 4096, sums are no longer exact; compare with the
    int128 closed form (rounded once) under the relative tolerance."""
    n = 4096
    comm = capi.cart_create((1, 1))
    s = capi.Summa(comm, n, 1024)
    s.fill(capi.FILL_INDEX)
    s.run()
    got = s.read_c_block(1000, 3000, 64, 512)
    want = oracle.index_fill_exact(n, 1000, 3000, 64, 512)
    assert oracle.rel_frobenius(got, want) <= 1e-14
    s.destroy()


def test_main_out_cli_single_rank(gpu, tmp_path):
    """`main.out <matrix_size> <tile_width> <grid_width> <grid_height> <test_name>` (reference
    src/main.c:28), file name of :75 and the 9-column record of src/utils.c:26-27."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (tmp_path / "csv").mkdir()
    for mode in ("host", "device"):
        env = dict(os.environ, PHPC_VERIFY="1", PHPC_MODE=mode)
        res = subprocess.run([os.path.join(root, "bin", "main.out"), "1024", "32", "1", "1", "testA"], capture_output=True, text=True,
                             timeout=300, cwd=tmp_path, env=env)
        assert res.returncode == 0, res.stdout + res.stderr
        rec = open(tmp_path / "csv" / "testA_N1024_T1_G1_TW32_GW1_GH1.csv").read().strip().split(",")
        assert rec[:6] == ["1024", "1", "1", "1", "1024", "1024"] and len(rec) == 9
        assert float(rec[6]) > 0 and float(rec[7]) > 0 and float(rec[8]) > 0
    res = subprocess.run([os.path.join(root, "bin", "main.out"), "64", "32", "1", "1", "nodir"], capture_output=True, text=True,
                         timeout=120, cwd=tmp_path / "csv")
    assert res.returncode != 0 and "Could not create CSV file" in res.stderr  # csv/ must pre-exist (reference :77-80)


# ----------------------------------------------------------------------------------------------
# FP64 on the tcgen05 tensor cores (Ozaki scheme, int8 digit products, int32 accumulators in TMEM)
# ----------------------------------------------------------------------------------------------
def _device_gemm_from_numpy(capi, lib, a, b, c0, backend):
    m, k = a.shape
    n = b.shape[1]
    lda, ldb = (k + 15) // 16 * 16, (n + 15) // 16 * 16
    dA, dB, dC = lib.phpc_device_malloc(max(m * lda, 2) * 8), lib.phpc_device_malloc(max(k * ldb, 2) * 8), lib.phpc_device_malloc(max(m * ldb, 2) * 8)
    lib.phpc_copy2d_to_device(dA, lda, capi._dp(a), k, m, k)
    lib.phpc_copy2d_to_device(dB, ldb, capi._dp(b), n, k, n)
    lib.phpc_copy2d_to_device(dC, ldb, capi._dp(c0), n, m, n)
    if backend == "ozaki":
        launched = lib.phpc_gemm_device_ozaki(dA, lda, dB, ldb, dC, ldb, m, k, n, None)
    else:
        launched = lib.phpc_gemm_device(dA, lda, dB, ldb, dC, ldb, m, k, n, 0, None)
    lib.phpc_device_synchronize()
    out = capi.device_window(dC, ldb, 0, 0, m, n)
    for p in (dA, dB, dC):
        lib.phpc_device_free(p)
    return out, launched


@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (128, 128, 128), (129, 33, 257), (100, 77, 50), (300, 215, 170), (512, 1024, 384), (257, 4100, 130)])
def test_ozaki_gemm_seeded_vs_oracle(gpu, capi, oracle, m, k, n):
    a = oracle.fill(m, k, kind=1, seed=101)
    b = oracle.fill(k, n, kind=1, seed=202)
    c0 = oracle.fill(m, n, kind=1, seed=303)
    c, launched = _device_gemm_from_numpy(capi, gpu, a, b, c0, "ozaki")
    assert launched >= 8  # exponent, guard, split and MMA kernels of at least one K chunk
    _check(oracle, c, oracle.gemm_block(a, b, c0), a, b)


@pytest.mark.parametrize("n", [32, 200, 512, 1024])
def test_ozaki_gemm_index_fill_bit_exact(gpu, capi, oracle, n):
    """Digits carry the integers exactly and every int32 / FP64 partial sum is exact: 0 ulp."""
    a = oracle.fill(n, n, kind=0)
    c, _ = _device_gemm_from_numpy(capi, gpu, a, a.copy(), np.zeros((n, n)), "ozaki")
    assert np.array_equal(c, oracle.index_fill_exact(n))


def test_ozaki_gemm_row_and_column_scaling(gpu, capi, oracle):
    """Per-row / per-column power-of-two scaling: rows and columns of wildly different magnitude
    (2^-300 .. 2^+300), zero rows and zero columns keep full relative accuracy per C element."""
    m, k, n = 96, 300, 80
    a = oracle.fill(m, k, kind=1, seed=7)
    b = oracle.fill(k, n, kind=1, seed=8)
    a *= np.ldexp(1.0, np.linspace(-300, 300, m).astype(int))[:, None]
    b *= np.ldexp(1.0, np.linspace(200, -200, n).astype(int))[None, :]
    a[5, :] = 0.0
    b[:, 9] = 0.0
    c, _ = _device_gemm_from_numpy(capi, gpu, a, b, np.zeros((m, n)), "ozaki")
    want = oracle.gemm_block(a, b)
    scale = np.abs(a) @ np.abs(b) + 1e-300
    assert np.all(np.abs(c - want) <= 4.0 * np.sqrt(k) * U * scale)
    assert np.all(c[5, :] == 0.0) and np.all(c[:, 9] == 0.0)


def test_summa_ozaki_backend_single_rank(gpu, capi, oracle):
    n = 384
    comm = capi.cart_create((1, 1))
    A = oracle.fill(n, n, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(n, n, kind=1, seed=oracle.SEED_B)
    want = oracle.gemm_block(A, B)
    for kc in (0, 100):
        s = capi.Summa(comm, n, kc)
        s.fill(capi.FILL_SEEDED)
        st = s.run(capi.BACKEND_OZAKI)
        assert st.launches >= 8 * st.steps
        assert oracle.rel_frobenius(s.read_c_block(), want) <= 1e-14
        s.destroy()
    s = capi.Summa(comm, 1024, 0)
    s.fill(capi.FILL_INDEX)
    s.run(capi.BACKEND_OZAKI)
    assert np.array_equal(s.read_c_block(), oracle.index_fill_exact(1024))
    s.destroy()


@pytest.mark.parametrize("m,k,n", [(15000, 7500, 7500), (7500, 3750, 15000)])
def test_config5_edge_tile_shapes_vs_cublas_and_exact_dots(gpu, capi, oracle, m, k, n):
    """BASELINE config 5 (N=30000): the per-step local GEMM shapes of the 2x4 / 1x2 grids — none of
    m, k, n is a multiple of the 128 / 16 tile.  DMMA and Ozaki kernels vs cuBLAS on the same device
    inputs, plus sampled elements against correctly rounded dot products."""
    lib = gpu
    lda, ldb = (k + 15) // 16 * 16, (n + 15) // 16 * 16
    dA, dB = lib.phpc_device_malloc(m * lda * 8), lib.phpc_device_malloc(k * ldb * 8)
    lib.phpc_fill_device(dA, lda, m, k, 0, 0, k, capi.FILL_SEEDED, 31, None)
    lib.phpc_fill_device(dB, ldb, k, n, 0, 0, n, capi.FILL_SEEDED, 32, None)
    outs = {}
    for name in ("cublas", "dmma", "ozaki"):
        dC = lib.phpc_device_malloc(m * ldb * 8)
        lib.phpc_device_memset(dC, 0, m * ldb * 8)
        if name == "cublas":
            lib.phpc_gemm_device_cublas(dA, lda, dB, ldb, dC, ldb, m, k, n, None)
        elif name == "dmma":
            lib.phpc_gemm_device(dA, lda, dB, ldb, dC, ldb, m, k, n, 0, None)
        else:
            lib.phpc_gemm_device_ozaki(dA, lda, dB, ldb, dC, ldb, m, k, n, None)
        lib.phpc_device_synchronize()
        # last rows / last columns: the edge tiles
        outs[name] = (capi.device_window(dC, ldb, m - 200, n - 300, 200, 300), capi.device_window(dC, ldb, 0, 0, 64, n))
        lib.phpc_device_free(dC)
    for name in ("dmma", "ozaki"):
        for w, wref in zip(outs[name], outs["cublas"]):
            assert oracle.rel_frobenius(w, wref) <= 1e-14, name
    a_rows = capi.device_window(dA, lda, m - 3, 0, 3, k)
    b_all = capi.device_window(dB, ldb, 0, n - 300, k, 300)
    for name in ("dmma", "ozaki"):
        edge = outs[name][0]
        for i in range(3):
            for j in (0, 150, 299):
                exact = oracle.dot_exact(a_rows[i], np.ascontiguousarray(b_all[:, j]))
                bound = 4.0 * np.sqrt(k) * U * float(np.abs(a_rows[i]) @ np.abs(b_all[:, j]))
                assert abs(edge[200 - 3 + i, j] - exact) <= bound, name
    lib.phpc_device_free(dA)
    lib.phpc_device_free(dB)


def test_entry_points_default_to_tcgen05_and_dmma_is_selectable(gpu, capi, oracle):
    """phpc_gemm_cuda runs the tcgen05 (Ozaki) kernel unless PHPC_GEMM=dmma; both meet the same bar."""
    import os

    m, k, n = 200, 333, 150
    a = oracle.fill(m, k, kind=1, seed=41)
    b = oracle.fill(k, n, kind=1, seed=42)
    want = oracle.gemm_block(a, b)
    results = {}
    for mode in ("ozaki", "dmma"):
        os.environ["PHPC_GEMM"] = mode
        c = np.zeros((m, n))
        capi.phpc_gemm_cuda(a, b, c)
        _check(oracle, c, want, a, b)
        results[mode] = c
    del os.environ["PHPC_GEMM"]
    c = np.zeros((m, n))
    capi.phpc_gemm_cuda(a, b, c)
    assert np.array_equal(c, results["ozaki"])           # the default is the tcgen05 kernel
    assert not np.array_equal(results["ozaki"], results["dmma"])  # different summation orders, same tolerance


def test_tcgen05_path_keeps_fp64_semantics_for_special_values(gpu, capi, oracle):
    """Inf / NaN, products near the underflow and overflow thresholds, and rows dominated by one huge entry: the guard of the
    tcgen05 launcher hands such K chunks to the native-FP64 kernel, so the result is what FP64 arithmetic gives (numpy here),
    not an emulation artefact.  ADVICE r01: rows AND columns scaled by 2^-500; an entry 2^67 above the rest of its row."""
    lib = gpu
    m, k, n = 64, 96, 80
    a = oracle.fill(m, k, kind=1, seed=51)
    b = oracle.fill(k, n, kind=1, seed=52)
    zero = np.zeros((m, n))
    lib.phpc_ozaki_fallback_chunks()  # reset the counter
    # (1) non-finite inputs propagate exactly like FP64
    a1, b1 = a.copy(), b.copy()
    a1[3, 7] = np.inf
    b1[11, 5] = np.nan
    c, _ = _device_gemm_from_numpy(capi, gpu, a1, b1, zero, "ozaki")
    with np.errstate(invalid="ignore"):
        want = a1 @ b1
    assert np.array_equal(np.isnan(c), np.isnan(want)) and np.array_equal(np.isinf(c), np.isinf(want))
    ok = np.isfinite(want)
    assert oracle.rel_frobenius(c[ok], want[ok]) <= 1e-14
    assert lib.phpc_ozaki_fallback_chunks() == 1
    # (2) rows and columns both scaled by 2^-500: products around 2^-1000 are normal doubles and must come out right
    a2, b2 = a * 2.0 ** -500, b * 2.0 ** -500
    c, _ = _device_gemm_from_numpy(capi, gpu, a2, b2, zero, "ozaki")
    assert oracle.rel_frobenius(c, oracle.gemm_block(a2, b2)) <= 1e-14 and np.all(c != 0.0)
    assert lib.phpc_ozaki_fallback_chunks() == 1
    # (3) 2^-600 each: every product underflows; FP64 gives (signed) zeros, so must we
    c, _ = _device_gemm_from_numpy(capi, gpu, a * 2.0 ** -600, b * 2.0 ** -600, zero, "ozaki")
    assert np.all(c == 0.0)
    # (4) overflow to +-Inf as in FP64
    c, _ = _device_gemm_from_numpy(capi, gpu, np.abs(a) * 2.0 ** 600, np.abs(b) * 2.0 ** 500, zero, "ozaki")
    assert np.all(np.isposinf(c))  # all products positive and far above 2^1024
    lib.phpc_ozaki_fallback_chunks()
    # (5) one entry 2^67 above the rest of its row, facing a zero in B: C[0][j] is made of the small entries only
    a5, b5 = a.copy(), b.copy()
    a5[0, 0] = 2.0 ** 67
    b5[0, :] = 0.0
    c, _ = _device_gemm_from_numpy(capi, gpu, a5, b5, zero, "ozaki")
    _check(oracle, c, oracle.gemm_block(a5, b5), a5, b5)
    assert lib.phpc_ozaki_fallback_chunks() == 1
    # (6) ordinary data never takes the fallback
    c, _ = _device_gemm_from_numpy(capi, gpu, a, b, zero, "ozaki")
    _check(oracle, c, oracle.gemm_block(a, b), a, b)
    assert lib.phpc_ozaki_fallback_chunks() == 0


def test_gemm_launches_on_different_streams_are_ordered(gpu, capi, oracle):
    """the launchers share per-device scratch (tile counter, digit stores, exponents, wave counters); launches on
    different streams of one device must not overlap.  Two tcgen05 GEMMs and a DMMA GEMM enqueued back to back on three
    streams, results checked."""
    import ctypes

    lib = gpu
    rt = ctypes.CDLL("libcudart.so.12")
    streams = []
    for _ in range(3):
        sp = ctypes.c_void_p()
        assert rt.cudaStreamCreateWithFlags(ctypes.byref(sp), 1) == 0
        streams.append(sp)
    m, k, n = 1500, 700, 1300
    lda, ldb = (k + 15) // 16 * 16, (n + 15) // 16 * 16
    mats = []
    for i in range(3):
        a = oracle.fill(m, k, kind=1, seed=61 + i)
        b = oracle.fill(k, n, kind=1, seed=71 + i)
        dA, dB, dC = lib.phpc_device_malloc(m * lda * 8), lib.phpc_device_malloc(k * ldb * 8), lib.phpc_device_malloc(m * ldb * 8)
        lib.phpc_copy2d_to_device(dA, lda, capi._dp(a), k, m, k)
        lib.phpc_copy2d_to_device(dB, ldb, capi._dp(b), n, k, n)
        lib.phpc_device_memset(dC, 0, m * ldb * 8)
        mats.append((a, b, dA, dB, dC))
    lib.phpc_device_synchronize()
    for i, (a, b, dA, dB, dC) in enumerate(mats):
        if i < 2:
            lib.phpc_gemm_device_ozaki(dA, lda, dB, ldb, dC, ldb, m, k, n, streams[i])
        else:
            lib.phpc_gemm_device(dA, lda, dB, ldb, dC, ldb, m, k, n, 0, streams[i])
    lib.phpc_device_synchronize()
    for (a, b, dA, dB, dC) in mats:
        got = capi.device_window(dC, ldb, 0, 0, m, n)
        assert oracle.rel_frobenius(got, a @ b) <= 1e-14
        for p in (dA, dB, dC):
            lib.phpc_device_free(p)
    for sp in streams:
        rt.cudaStreamDestroy(sp)


@pytest.mark.parametrize("mode", ["ozaki", "dmma"])
def test_host_bands_bit_identical_to_chunk_pipeline(gpu, capi, oracle, mode):
    """The band pipeline of the single-GPU host path (PHPC_HOST_BANDS; default 4 bands once the block has
    >= 8192 rows) only reorders independent work: per element the K chunks are still added in ascending
    order, so the result must equal the chunk-pipelined loop bit for bit, and the reference's own
    input must still come out exact.  Its operation list is proven race free in tests/test_host_plan.py."""
    import os

    lib = gpu
    n = 1024
    comm = capi.cart_create((1, 1))
    A = oracle.fill(n, n, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(n, n, kind=1, seed=oracle.SEED_B)
    C0 = oracle.fill(n, n, kind=1, seed=77)
    Ai = oracle.fill(n, n, kind=0)
    saved = {k: os.environ.get(k) for k in ("PHPC_GEMM", "PHPC_KC", "PHPC_HOST_BANDS")}
    out = {}
    try:
        os.environ["PHPC_GEMM"] = mode
        os.environ["PHPC_KC"] = "256"  # 4 K chunks
        for bands in (1, 3):  # 3 bands of 384, 384, 256 rows
            os.environ["PHPC_HOST_BANDS"] = str(bands)
            lib.phpc_summa_release_cache()
            C = C0.copy()
            capi.phpc_gemm_summa_cuda(comm, A, B, C)
            capi.phpc_gemm_summa_cuda(comm, A, B, C)  # second pass on the same C, as reference main.c does
            Ci = np.zeros((n, n))
            capi.phpc_gemm_summa_cuda(comm, Ai, Ai.copy(), Ci)
            out[bands] = (C, Ci)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        lib.phpc_summa_release_cache()
    assert np.array_equal(out[1][0], out[3][0])
    assert np.array_equal(out[3][1], oracle.index_fill_exact(n))
    want = C0 + 2.0 * oracle.gemm_block(A, B)
    assert oracle.rel_frobenius(out[3][0], want) <= 1e-14
