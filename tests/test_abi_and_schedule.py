"""CPU-only checks of the boundary: the shared library loads without a GPU, exports every
symbol include/*.h declares, and the SUMMA schedule follows reference src/phpc_summa.c:36-95."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", text, flags=re.S)
    return [n for n in names if n not in ("defined",)]


@pytest.mark.parametrize("header", ["phpc_gemm.cuh", "phpc_summa.h", "phpc_b200.h", "utils.h"])
def test_library_exports_every_declared_symbol(capi, header):
    lib = capi.load()
    names = _declared_functions(header)
    assert names, f"no declarations parsed from {header}"
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/{header} but not exported"


def test_reference_entry_points_present(capi):
    lib = capi.load()
    for name in ("phpc_gemm_cuda", "phpc_gemm_cublas", "phpc_gemm_summa_cuda", "phpc_gemm_summa_cublas", "get_cur_time", "log_to_csv"):
        assert hasattr(lib, name)
    assert lib.phpc_b200_version() >= 100
    assert lib.phpc_b200_device_count() >= 0  # never aborts without a GPU


def test_log_to_csv_record_is_byte_compatible(capi, tmp_path):
    """reference src/utils.c:26-27: '%d,%d,%d,%d,%d,%d,%f,%f,%f\\n' with total_threads computed."""
    lib = capi.load()
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p
    libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [ctypes.c_void_p]
    lib.log_to_csv.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_double, ctypes.c_float, ctypes.c_double]
    path = str(tmp_path / "x.csv")
    f = libc.fopen(path.encode(), b"w")
    lib.log_to_csv(f, 2048, 4, 1, 64, 1024, 0.45774, 0.25, 1.5)
    libc.fclose(f)
    assert open(path).read() == "2048,4,1,64,1024,65536,0.457740,0.250000,1.500000\n"


GRIDS = [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (4, 2), (2, 3), (4, 4)]


@pytest.mark.parametrize("r,c", GRIDS)
@pytest.mark.parametrize("kc", [0, 5, 16])
def test_schedule_matches_reference_ownership(capi, oracle, r, c, kc):
    lcm = oracle.find_lcm(r, c)
    N = lcm * 12
    owner_col, owner_row = oracle.summa_owners(r, c)
    pk = N // lcm
    for pi in range(r):
        for pj in range(c):
            steps, m, n = capi.summa_schedule(N, r, c, pi, pj, kc)
            assert (m, n) == (N // r, N // c)
            # chunks tile K exactly once, in ascending order, never crossing a panel
            k = 0
            for s in steps:
                assert s.k0 == k and s.width > 0
                assert s.k0 // pk == (s.k0 + s.width - 1) // pk == s.panel
                k += s.width
                assert s.a_root == owner_col[s.panel] and s.b_root == owner_row[s.panel]
                assert s.own_a == int(pj == s.a_root) and s.own_b == int(pi == s.b_root)
            assert k == N
            # owned chunks are stored back to back without overlap
            a_end = 0
            for s in steps:
                if s.own_a:
                    assert s.a_off == a_end
                    a_end += m * ((s.width + 15) // 16 * 16)
            owned_b = sorted((s.b_off, s.width) for s in steps if s.own_b)
            ldn = (n + 15) // 16 * 16
            end = 0
            for off, w in owned_b:
                assert off == end
                end += w * ldn
            assert sum(s.width for s in steps if s.own_a) == N // c
            assert sum(s.width for s in steps if s.own_b) == N // r


def test_schedule_rejects_indivisible(capi):
    with pytest.raises(ValueError):
        capi.summa_schedule(30, 4, 2, 0, 0)


def test_schedule_drives_a_cpu_emulation_to_the_oracle_result(capi, oracle):
    """Execute the product's schedule with the oracle's block GEMM standing in for the
    kernel: every rank multiplies exactly the chunks the plan names."""
    N, r, c, kc = 48, 2, 4, 5
    A = oracle.fill(N, N, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(N, N, kind=1, seed=oracle.SEED_B)
    C = np.zeros((N, N))
    for pi in range(r):
        for pj in range(c):
            steps, m, n = capi.summa_schedule(N, r, c, pi, pj, kc)
            blk = np.zeros((m, n))
            for s in steps:
                a = A[pi * m:(pi + 1) * m, s.k0:s.k0 + s.width]
                b = B[s.k0:s.k0 + s.width, pj * n:(pj + 1) * n]
                blk = oracle.gemm_block(a, b, blk)
            C[pi * m:(pi + 1) * m, pj * n:(pj + 1) * n] = blk
    assert oracle.rel_frobenius(C, oracle.summa(A, B, r, c)) < 1e-15


def test_csv_runner_reads_the_reference_config_format(built, tmp_path):
    """scripts/run_tests_csv.py consumes run-configuration CSVs in the reference's format (tests/*.csv)."""
    import subprocess
    import sys

    cfg = tmp_path / "cfg.csv"
    cfg.write_text("matrix_size,n_proc,n_gpu,tile_width,grid_width,grid_height\n128,1,1,32,1,1\n128,4,1,32,1,1\n512,8,2,16,2,2\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_tests_csv.py"), str(cfg), "--name", "t", "--dry-run"],
                         capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0].endswith("bin/main.out 128 32 1 1 t")
    assert "mpirun --oversubscribe -n 4" in lines[1] and lines[1].endswith("main.out 128 32 1 1 t")
    assert lines[2].startswith("PHPC_PGRID=2x4 ") and lines[2].endswith("main.out 512 16 2 2 t")
    b200 = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_tests_csv.py"),
                           os.path.join(ROOT, "tests", "configs", "b200_configs.csv"), "--dry-run"], capture_output=True, text=True, timeout=60)
    assert b200.returncode == 0 and len(b200.stdout.strip().splitlines()) == 10


def test_ozaki_config_is_fixed(capi, monkeypatch):
    """One tcgen05 kernel, one arithmetic: 7 balanced base-256 digits, 28 products, K chunks of 8192; only PHPC_GEMM=dmma
    switches the reference-named entry points to the native-FP64 kernel."""
    assert capi.ozaki_config() == {"digits": 7, "products": 28, "k_chunk": 8192, "max_spread": 40}
    monkeypatch.delenv("PHPC_GEMM", raising=False)
    assert capi.default_backend() == capi.BACKEND_OZAKI
    monkeypatch.setenv("PHPC_GEMM", "dmma")
    assert capi.default_backend() == capi.BACKEND_DMMA


def test_rectangular_schedule_tiles_k_exactly_and_owners_follow_the_reference(capi):
    """phpc_summa_schedule_mkn: K panels of K/lcm owned as in the reference (A panel k by column k % c, B panel k by row k % r,
    src/phpc_summa.c:64-65), cut into chunks that tile [0, K) exactly on every rank; the square call is the M = K = N case."""
    M, K, N = 600, 1080, 792
    for (r, c) in ((1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (4, 2), (2, 3)):
        lcm = r * c // __import__("math").gcd(r, c)
        for pi in range(r):
            for pj in range(c):
                steps, m, n = capi.summa_schedule_mkn(M, K, N, r, c, pi, pj, 100)
                assert (m, n) == (M // r, N // c)
                assert steps[0].k0 == 0 and all(steps[i].k0 + steps[i].width == steps[i + 1].k0 for i in range(len(steps) - 1))
                assert steps[-1].k0 + steps[-1].width == K
                for st in steps:
                    assert st.panel == st.k0 // (K // lcm)
                    assert (st.a_root, st.b_root) == (st.panel % c, st.panel % r)
                    assert st.own_a == (st.a_root == pj) and st.own_b == (st.b_root == pi)
                    assert (st.a_off >= 0) == bool(st.own_a) and (st.b_off >= 0) == bool(st.own_b)
    sq, m, n = capi.summa_schedule(768, 2, 4, 1, 3, 100)
    rq, m2, n2 = capi.summa_schedule_mkn(768, 768, 768, 2, 4, 1, 3, 100)
    assert (m, n) == (m2, n2) and [(a.k0, a.width, a.a_off, a.b_off) for a in sq] == [(b.k0, b.width, b.a_off, b.b_off) for b in rq]
    with pytest.raises(ValueError):
        capi.summa_schedule_mkn(600, 1081, 792, 2, 4, 0, 0)


def test_short_first_chunk_schedule(capi):
    """Multi-rank objects start with a short K chunk (nothing can overlap the transfer of the first chunk): the chunks still
    tile [0, K) exactly, only the first one is shortened, and every rank of the grid sees the same chunk boundaries."""
    N, kc, kf = 32768, 8192, 2048
    ref = None
    for pi in range(2):
        for pj in range(4):
            steps = capi.summa_schedule_first(N, N, N, 2, 4, pi, pj, kc, kf)
            bounds = [(s.k0, s.width) for s in steps]
            assert bounds[0] == (0, kf) and bounds[1] == (kf, kc - kf) and all(w == kc for _, w in bounds[2:])
            assert sum(w for _, w in bounds) == N and all(a + w == b for (a, w), (b, _) in zip(bounds, bounds[1:]))
            ref = ref or bounds
            assert bounds == ref
            own = [s for s in steps if s.own_a]
            offs = [s.a_off for s in own]
            assert offs == sorted(offs) and offs[0] == 0  # owned chunks are stored back to back in step order
    assert [(s.k0, s.width) for s in capi.summa_schedule_first(N, N, N, 2, 4, 0, 0, kc, 0)] == [(s.k0, s.width) for s in capi.summa_schedule(N, 2, 4, 0, 0, kc)[0]]
