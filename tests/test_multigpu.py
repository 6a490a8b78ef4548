"""SUMMA on several GPUs (NCCL row/column broadcasts over NVLink) against the oracle's
restatement of reference src/phpc_summa.c.  Needs >= 2 GPUs: run with `gpurun --gpus N`."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need(gpu, n):
    have = gpu.phpc_b200_device_count()
    if have < n:
        pytest.skip(f"needs {n} GPUs, {have} visible")


def _run(grid, N, fill, tmp_path, kc=0, env=None):
    r, c = grid
    out = str(tmp_path / f"c_{r}x{c}_{N}_{fill}.npy")
    cmd = [os.path.join(ROOT, "bin", "mpirun"), "-n", str(r * c), sys.executable, os.path.join(ROOT, "tests", "_summa_gpu_worker.py"),
           f"{r}x{c}", str(N), str(fill), out, str(kc)]
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=e)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return np.load(out), res.stdout


GRID_CASES = [((1, 2), 2), ((2, 1), 2), ((2, 2), 4), ((2, 4), 8), ((4, 2), 8)]


@pytest.mark.parametrize("grid,ngpu", GRID_CASES)
def test_summa_matches_reference_summa_seeded(gpu, oracle, tmp_path, grid, ngpu):
    _need(gpu, ngpu)
    N = 480  # not a multiple of the 128 tile on any grid: edge tiles everywhere
    Cs, log = _run(grid, N, 1, tmp_path, kc=50)
    A = oracle.fill(N, N, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(N, N, kind=1, seed=oracle.SEED_B)
    want = oracle.summa(A, B, *grid)
    names = ("host-entry tcgen05 (default)", "host-entry cublas", "device-resident dmma", "device-resident tcgen05",
             "host-entry, rank 0's C node-shared (band-pipelined run, parallel gather)", "the same, second C += pass, halved")
    assert len(Cs) == len(names)
    for name, C in zip(names, Cs):
        assert oracle.rel_frobenius(C, want) <= 1e-14, name
    assert np.array_equal(Cs[4], Cs[0])  # bands only regroup rows: same arithmetic per element as the chunk loop
    assert "bcasts=" in log


@pytest.mark.parametrize("grid,ngpu", GRID_CASES)
def test_summa_bit_exact_on_reference_fill(gpu, oracle, tmp_path, grid, ngpu):
    """A[i]=B[i]=i at N=1024 (BASELINE config 1 input): exact integers, any grid, any order."""
    _need(gpu, ngpu)
    N = 1024
    Cs, _ = _run(grid, N, 0, tmp_path, kc=96)
    exact = oracle.index_fill_exact(N)
    for C in Cs:
        assert np.array_equal(C, exact)


@pytest.mark.parametrize("grid,ngpu", [((1, 2), 2), ((2, 1), 2), ((2, 2), 4), ((2, 4), 8)])
def test_rectangular_summa_object_api(gpu, oracle, tmp_path, grid, ngpu):
    """general C[M x N] += A[M x K] * B[K x N] behind the device-resident object API (the reference is square
    only, src/phpc_summa.c:36-39).  M, K, N all different, none a multiple of the 128 tile per block; device-generated blocks
    and host-uploaded blocks with a nonzero C, against the oracle's block GEMM on the regenerated full matrices."""
    _need(gpu, ngpu)
    M, K, N = 2 * 4 * 75, 8 * 135, 4 * 2 * 99  # divisible by every grid dimension and lcm used here
    Cs, _ = _run(grid, N, 1, tmp_path, kc=100, env={"PHPC_TEST_RECT": f"{M},{K},{N}"})
    A = oracle.fill(M, K, N=K, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(K, N, N=N, kind=1, seed=oracle.SEED_B)
    assert oracle.rel_frobenius(Cs[0], oracle.gemm_block(A, B)) <= 1e-14
    A2 = oracle.fill(M, K, N=K, kind=1, seed=91)
    B2 = oracle.fill(K, N, N=N, kind=1, seed=92)
    C0 = oracle.fill(M, N, N=N, kind=1, seed=93)
    assert oracle.rel_frobenius(Cs[1], oracle.gemm_block(A2, B2, C0)) <= 1e-14


@pytest.mark.parametrize("bands", [1, 3, 4])
def test_band_pipelined_multi_rank_host_run(gpu, oracle, tmp_path, bands):
    """The multi-rank host-sourced call with rank 0's C in node-shared memory (phpc_host_malloc_shared): the C block travels in
    row bands, the k-loop runs once per band, every rank writes its finished bands into rank 0's matrix.  Any band count gives
    the reference's result exactly on its own input (N = 1024) and the chunk loop's result bit for bit."""
    _need(gpu, 2)
    grid = (2, 2) if gpu.phpc_b200_device_count() >= 4 else (1, 2)
    Cs, _ = _run(grid, 1024, 0, tmp_path, kc=96, env={"PHPC_HOST_BANDS": str(bands)})
    exact = oracle.index_fill_exact(1024)
    for C in Cs:
        assert np.array_equal(C, exact)
    Cr, _ = _run(grid, 640, 1, tmp_path, kc=100, env={"PHPC_HOST_BANDS": str(bands)})
    assert np.array_equal(Cr[4], Cr[0])
    A = oracle.fill(640, 640, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(640, 640, kind=1, seed=oracle.SEED_B)
    assert oracle.rel_frobenius(Cr[5], oracle.summa(A, B, *grid)) <= 1e-14


def test_golden_reference_summa_fixture(gpu, oracle, tmp_path):
    """The committed outputs of the reference's own SUMMA (4 ranks -> 2x2, N=48)."""
    _need(gpu, 4)
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
    Cs, _ = _run((2, 2), 48, 1, tmp_path)
    for C in Cs:
        assert oracle.rel_frobenius(C, g["summa_N48_P4_F1"]) <= 1e-14


def test_nccl_broadcast_transport_equals_pull_transport(gpu, oracle, tmp_path):
    """PHPC_PANEL=nccl (ncclBroadcast on row/column communicators) vs the default copy-engine pull:
    the same chunks reach the same GEMMs, so the results are bit-identical."""
    _need(gpu, 2)
    grid = (2, 2) if gpu.phpc_b200_device_count() >= 4 else (1, 2)
    a, _ = _run(grid, 384, 1, tmp_path, kc=50)
    b, _ = _run(grid, 384, 1, tmp_path, kc=50, env={"PHPC_PANEL": "nccl"})
    for idx, name in ((0, "host-entry tcgen05 (default)"), (2, "device-resident dmma"), (3, "device-resident tcgen05")):
        diff = np.abs(a[idx] - b[idx])
        assert np.array_equal(a[idx], b[idx]), f"{name}: {np.count_nonzero(diff)} elements differ, max {diff.max():.3e}"
    # cuBLAS picks its kernel per call; only require agreement to rounding
    assert oracle.rel_frobenius(a[1], b[1]) <= 1e-14


def test_nccl_registered_buffers_equal_pull_transport(gpu, tmp_path):
    """PHPC_NCCL_REGISTER=1: stores and receive rings registered with the row / column communicators
    (ncclCommRegister). Registration changes how NCCL moves the bytes, never the bytes."""
    _need(gpu, 2)
    grid = (2, 2) if gpu.phpc_b200_device_count() >= 4 else (1, 2)
    a, _ = _run(grid, 384, 1, tmp_path, kc=50)
    b, _ = _run(grid, 384, 1, tmp_path, kc=50, env={"PHPC_PANEL": "nccl", "PHPC_NCCL_REGISTER": "1"})
    for idx in (0, 2, 3):
        assert np.array_equal(a[idx], b[idx]), idx


def test_prefetch_all_schedule_equals_default(gpu, oracle, tmp_path):
    """PHPC_SCHEDULE=prefetch-all (stationary C, every panel transfer issued up front into a ring that holds
    all chunks) moves the same chunks into the same GEMMs in the same order: bit-identical to the 3-slot ring."""
    _need(gpu, 2)
    grid = (2, 2) if gpu.phpc_b200_device_count() >= 4 else (1, 2)
    a, _ = _run(grid, 384, 1, tmp_path, kc=50)
    b, log = _run(grid, 384, 1, tmp_path, kc=50, env={"PHPC_SCHEDULE": "prefetch-all"})
    for idx in (0, 2, 3, 4):
        assert np.array_equal(a[idx], b[idx]), idx


def test_mpi_gather_path_equals_nvlink_gather(gpu, oracle, tmp_path):
    _need(gpu, 2)
    a, _ = _run((1, 2), 256, 1, tmp_path)
    b, _ = _run((1, 2), 256, 1, tmp_path, env={"PHPC_GATHER": "mpi"})
    assert np.array_equal(a, b)


def test_main_out_cli_multi_rank(gpu, tmp_path):
    """The drop-in driver under the reference's command line, self-verifying (PHPC_VERIFY)."""
    _need(gpu, 2)
    (tmp_path / "csv").mkdir()
    for mode in ("host", "device"):
        env = dict(os.environ, PHPC_VERIFY="1", PHPC_MODE=mode)
        res = subprocess.run([os.path.join(ROOT, "bin", "mpirun"), "--oversubscribe", "-n", "2", os.path.join(ROOT, "bin", "main.out"),
                              "1024", "32", "1", "1", "testA"], capture_output=True, text=True, timeout=300, cwd=tmp_path, env=env)
        assert res.returncode == 0, res.stdout + res.stderr
        line = open(tmp_path / "csv" / "testA_N1024_T2_G1_TW32_GW1_GH1.csv").read().strip().split(",")
        assert line[:6] == ["1024", "2", "1", "1", "1024", "1024"] and len(line) == 9


@pytest.mark.parametrize("tile_width", [16, 32, 64, 128, 256])
def test_config5_n30000_tile_width_sweep_1x2(gpu, tmp_path, tile_width):
    """BASELINE config 5: non-divisible N=30000 on the rectangular 1x2 grid, every tile_width of the
    sweep; main.out verifies sampled elements of every rank's block against the closed form."""
    _need(gpu, 2)
    (tmp_path / "csv").mkdir()
    env = dict(os.environ, PHPC_VERIFY="1", PHPC_MODE="device", PHPC_PGRID="1x2")
    res = subprocess.run([os.path.join(ROOT, "bin", "mpirun"), "-n", "2", os.path.join(ROOT, "bin", "main.out"), "30000", str(tile_width), "1", "1",
                          "cfg5"], capture_output=True, text=True, timeout=300, cwd=tmp_path, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    rec = open(tmp_path / "csv" / f"cfg5_N30000_T2_G1_TW{tile_width}_GW1_GH1.csv").read().strip().split(",")
    assert rec[0] == "30000" and rec[1] == "2" and float(rec[6]) > 0


@pytest.mark.parametrize("gpus", [2, 3])
def test_gemm_cuda_column_split_over_local_gpus(gpu, capi, oracle, gpus):
    """phpc_gemm_cuda(..., gpu_count > 1): the reference replicates A and slices B and C by columns over the local GPUs,
    dev_n = n/g + (gpu < n%g) (src/phpc_gemm.cu:97-129).  Same split here, one in-process call, n not divisible by g."""
    _need(gpu, gpus)
    m, k, n = 260, 190, 301  # 301 = 2*150 + 1 = 3*100 + 1: uneven slices on 2 and on 3 GPUs
    a = oracle.fill(m, k, kind=1, seed=71)
    b = oracle.fill(k, n, kind=1, seed=72)
    c0 = oracle.fill(m, n, kind=1, seed=73)
    want = oracle.gemm_block(a, b, c0)
    c = c0.copy()
    secs = capi.phpc_gemm_cuda(a, b, c, gpus, 1, 1, 32)
    assert secs > 0.0  # mean over the GPUs of the kernel event time (reference :145)
    assert oracle.rel_frobenius(c, want) <= 1e-14
    one = c0.copy()
    capi.phpc_gemm_cuda(a, b, one, 1, 1, 1, 32)
    assert np.array_equal(c, one)  # a column split does not change any element's arithmetic
    ai = oracle.fill(128, 128, kind=0)
    ci = np.zeros((128, 128))
    capi.phpc_gemm_cuda(ai, ai.copy(), ci, gpus, 1, 1, 32)
    assert np.array_equal(ci, oracle.index_fill_exact(128))
    gpu.phpc_b200_set_device(0)
