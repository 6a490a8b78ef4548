"""The experimental Ozaki split kernels (balanced base-256 digits; half-major B store of the 2-CTA kernel) have
__host__ __device__ bodies: this test compiles tests/csrc/oz_host_probe.cu with nvcc AS HOST CODE and runs the very
lines the GPU executes (digit extraction, store addressing) on the CPU against oracle/ozaki_model.py and against
the store layout the kernels' loads assume (csrc/ozaki_gemm.cuh, csrc/ozaki_gemm2.cuh).  No GPU needed."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import ozaki_model as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_int8_p = ctypes.POINTER(ctypes.c_int8)


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("ozprobe") / "liboz_probe.so")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "csrc", "oz_host_probe.cu")], check=True, capture_output=True)
    L = ctypes.CDLL(out)
    L.oz_probe_store_offset.restype = ctypes.c_longlong
    L.oz_probe_digits.argtypes = [ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_int, c_int8_p]
    L.oz_probe_split_a.argtypes = [ctypes.c_int, c_double_p, ctypes.c_longlong] + [ctypes.c_int] * 4 + [c_int_p, c_int8_p]
    L.oz_probe_split_b.argtypes = [ctypes.c_int, c_double_p, ctypes.c_longlong] + [ctypes.c_int] * 4 + [c_int_p, c_int8_p, ctypes.c_int]
    return L


def _exps_array(exps, zero):
    return np.array([zero if e is None else e for e in exps], dtype=np.int32)


def _model_digits(x, exps, axis, bal):
    return om.split_digits_balanced(x, exps, axis, 7) if bal else om.split_digits(x, exps, axis, 8)[0]


def _layout_offset(row, kbyte, t, S, ksteps, halves):
    """The tiled digit store as the GEMM kernels read it (header comments of ozaki_gemm.cuh / ozaki_gemm2.cuh):
    store[row tile][k step][half][digit][canonical K-major tile]; inside a tile 8-row x 16-byte core matrices,
    the two k chunks of a 32-byte step 128 B apart, 8-row groups 256 B apart."""
    tile, r, ks, kb = row // 128, row % 128, kbyte // 32, kbyte % 32
    rows_per_half = 128 // halves
    h, rh = r // rows_per_half, r % rows_per_half
    inner = (rh // 8) * 256 + (kb // 16) * 128 + (rh % 8) * 16 + kb % 16
    return ((((tile * ksteps + ks) * halves + h) * S + t) * rows_per_half * 32) + inner


@pytest.mark.parametrize("bal", [0, 1])
def test_digits_match_the_model(probe, oracle, bal):
    S = 7 if bal else 8
    x = oracle.fill(1, 4096, kind=1, seed=21) * np.ldexp(1.0, np.arange(4096) % 60 - 59)[None, :]
    x[0, :6] = [0.0, 0.999999999, -0.999999999, 2.0 ** -60, -(2.0 ** -54), 0.5]
    exps = [0]  # |x| < 2^0
    want = np.stack(_model_digits(x, exps, 1, bal), axis=-1)[0]  # [n][S]
    got = np.zeros((x.shape[1], S), dtype=np.int8)
    probe.oz_probe_digits(bal, x.ctypes.data_as(c_double_p), x.shape[1], 0, got.ctypes.data_as(c_int8_p))
    assert np.array_equal(got.astype(np.int64), want)
    if bal:  # 54 bits below the scale, correctly rounded: |x - sum d_t 256^(6-t) 2^-54| <= 2^-55
        recon = sum(got[:, t].astype(object) * (256 ** (6 - t)) for t in range(7))
        for v, q in zip(x[0, :64], recon[:64]):
            assert abs(int(q) - v * 2.0 ** 54) <= 0.5
    # an all-zero / non-finite row (ZERO_EXP, NONFINITE_EXP) yields zero digits
    for e in (probe.oz_probe_zero_exp(), 2147483647):
        probe.oz_probe_digits(bal, x.ctypes.data_as(c_double_p), 16, e, got.ctypes.data_as(c_int8_p))
        assert not got[:16].any()


@pytest.mark.parametrize("bal", [0, 1])
@pytest.mark.parametrize("pad256", [False, True])
def test_a_store_layout(probe, oracle, bal, pad256):
    S = 7 if bal else 8
    m, k = 200, 75
    kp = 128
    m_pad = 512 if pad256 else 256  # 200 rows -> 2 tiles; the 2-CTA kernel pads to an even tile count (here 4)
    a = oracle.fill(m, k, kind=1, seed=31) * np.ldexp(1.0, (np.arange(m) % 9) * 11 - 40)[:, None]
    a[5, :] = 0.0
    exps = om.exponents(a, 1)
    digits = _model_digits(a, exps, 1, bal)
    eA = _exps_array(exps, probe.oz_probe_zero_exp())
    TA = np.full(S * m_pad * kp, 0x55, dtype=np.int8)
    probe.oz_probe_split_a(bal, a.ctypes.data_as(c_double_p), k, m, m_pad, k, kp, eA.ctypes.data_as(c_int_p), TA.ctypes.data_as(c_int8_p))
    want = np.zeros_like(TA)
    rows, cols = np.meshgrid(np.arange(m), np.arange(k), indexing="ij")
    for t in range(S):
        off = np.vectorize(_layout_offset)(rows, cols, t, S, kp // 32, 1)
        want[off] = digits[t].astype(np.int8)
    assert np.array_equal(TA, want)  # every byte written; padding rows / k padding are zero digits
    assert probe.oz_probe_store_offset(130, 70, 3, S, kp // 32, 1) == _layout_offset(130, 70, 3, S, kp // 32, 1)


@pytest.mark.parametrize("bal", [0, 1])
@pytest.mark.parametrize("halves", [1, 2])
def test_b_store_layout(probe, oracle, bal, halves):
    S = 7 if bal else 8
    k, n = 75, 300
    kp, n_pad = 128, 384
    b = oracle.fill(k, n, kind=1, seed=32) * np.ldexp(1.0, (np.arange(n) % 7) * 13 - 30)[None, :]
    b[:, 17] = 0.0
    exps = om.exponents(b, 0)
    digits = _model_digits(b, exps, 0, bal)
    eB = _exps_array(exps, probe.oz_probe_zero_exp())
    TB = np.full(S * n_pad * kp, 0x55, dtype=np.int8)
    probe.oz_probe_split_b(bal, b.ctypes.data_as(c_double_p), n, k, n, n_pad, kp, eB.ctypes.data_as(c_int_p), TB.ctypes.data_as(c_int8_p), halves)
    want = np.zeros_like(TB)
    ks_, cols = np.meshgrid(np.arange(k), np.arange(n), indexing="ij")
    for t in range(S):
        off = np.vectorize(_layout_offset)(cols, ks_, t, S, kp // 32, halves)  # B^T: the store row is the output column
        want[off] = digits[t].astype(np.int8)
    assert np.array_equal(TB, want)
    if halves == 2:
        # what CTA `rank` of a pair loads for one k step: ONE contiguous range of S * 2 KiB holding rows rank*64 .. +63
        tile, ks, rank = 1, 2, 1
        base = ((tile * (kp // 32) + ks) * 2 + rank) * S * 2048
        for t in (0, S - 1):
            for r in (0, 63):
                for kb in (0, 31):
                    col, kk = tile * 128 + rank * 64 + r, ks * 32 + kb
                    expect = digits[t][kk, col] if (kk < k and col < n) else 0
                    assert TB[base + t * 2048 + (r // 8) * 256 + (kb // 16) * 128 + (r % 8) * 16 + kb % 16] == expect
