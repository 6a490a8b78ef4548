"""The split kernels of the tcgen05 path (balanced base-256 digits, tiled digit stores) have __host__ __device__ bodies:
this test compiles tests/csrc/oz_host_probe.cu with nvcc AS HOST CODE and runs the very lines the GPU executes (digit
extraction, store addressing) on the CPU against oracle/ozaki_model.py and against the store layout the GEMM kernel's
loads and paired N = 256 MMAs assume (csrc/ozaki_gemm.cuh).  No GPU needed."""
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
S = 7


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
    L.oz_probe_digits.argtypes = [c_double_p, ctypes.c_int, ctypes.c_int, c_int8_p]
    L.oz_probe_split_a.argtypes = [c_double_p, ctypes.c_longlong] + [ctypes.c_int] * 4 + [c_int_p, c_int8_p]
    L.oz_probe_split_b.argtypes = [c_double_p, ctypes.c_longlong] + [ctypes.c_int] * 4 + [c_int_p, c_int8_p]
    assert L.oz_probe_digits_per_operand() == S
    return L


def _exps_array(exps, zero):
    return np.array([zero if e is None else e for e in exps], dtype=np.int32)


def _layout_offset(row, kbyte, t, ksteps):
    """The tiled digit store as the GEMM kernel reads it (header comment of ozaki_gemm.cuh):
    store[row tile][k step][digit][canonical K-major 4 KiB tile]; inside a tile 8-row x 16-byte core matrices,
    the two k chunks of a 32-byte step 128 B apart, 8-row groups 256 B apart."""
    tile, r, ks, kb = row // 128, row % 128, kbyte // 32, kbyte % 32
    inner = (r // 8) * 256 + (kb // 16) * 128 + (r % 8) * 16 + kb % 16
    return (((tile * ksteps + ks) * S + t) * 4096) + inner


def test_digits_match_the_model(probe, oracle):
    x = oracle.fill(1, 4096, kind=1, seed=21) * np.ldexp(1.0, np.arange(4096) % 60 - 59)[None, :]
    x[0, :6] = [0.0, 0.999999999, -0.999999999, 2.0 ** -60, -(2.0 ** -54), 0.5]
    exps = [0]  # |x| < 2^0
    want = np.stack(om.split_digits_balanced(x, exps, 1, S), axis=-1)[0]  # [n][S]
    got = np.zeros((x.shape[1], S), dtype=np.int8)
    probe.oz_probe_digits(x.ctypes.data_as(c_double_p), x.shape[1], 0, got.ctypes.data_as(c_int8_p))
    assert np.array_equal(got.astype(np.int64), want)
    # 54 bits below the scale, correctly rounded: |x - sum d_t 256^(6-t) 2^-54| <= 2^-55
    recon = sum(got[:, t].astype(object) * (256 ** (6 - t)) for t in range(S))
    for v, q in zip(x[0, :64], recon[:64]):
        assert abs(int(q) - v * 2.0 ** 54) <= 0.5
    # an all-zero / non-finite row (ZERO_EXP, NONFINITE_EXP) yields zero digits
    for e in (probe.oz_probe_zero_exp(), 2147483647):
        probe.oz_probe_digits(x.ctypes.data_as(c_double_p), 16, e, got.ctypes.data_as(c_int8_p))
        assert not got[:16].any()


def test_a_store_layout(probe, oracle):
    m, k = 200, 75
    kp, m_pad = 128, 256
    a = oracle.fill(m, k, kind=1, seed=31) * np.ldexp(1.0, (np.arange(m) % 9) * 11 - 40)[:, None]
    a[5, :] = 0.0
    exps = om.exponents(a, 1)
    digits = om.split_digits_balanced(a, exps, 1, S)
    eA = _exps_array(exps, probe.oz_probe_zero_exp())
    TA = np.full(S * m_pad * kp, 0x55, dtype=np.int8)
    probe.oz_probe_split_a(a.ctypes.data_as(c_double_p), k, m, m_pad, k, kp, eA.ctypes.data_as(c_int_p), TA.ctypes.data_as(c_int8_p))
    want = np.zeros_like(TA)
    rows, cols = np.meshgrid(np.arange(m), np.arange(k), indexing="ij")
    for t in range(S):
        off = np.vectorize(_layout_offset)(rows, cols, t, kp // 32)
        want[off] = digits[t].astype(np.int8)
    assert np.array_equal(TA, want)  # every byte written; padding rows / k padding are zero digits
    assert probe.oz_probe_store_offset(130, 70, 3, kp // 32) == _layout_offset(130, 70, 3, kp // 32)


def test_b_store_layout_and_paired_operand(probe, oracle):
    k, n = 75, 300
    kp, n_pad = 128, 384
    b = oracle.fill(k, n, kind=1, seed=32) * np.ldexp(1.0, (np.arange(n) % 7) * 13 - 30)[None, :]
    b[:, 17] = 0.0
    exps = om.exponents(b, 0)
    digits = om.split_digits_balanced(b, exps, 0, S)
    eB = _exps_array(exps, probe.oz_probe_zero_exp())
    TB = np.full(S * n_pad * kp, 0x55, dtype=np.int8)
    probe.oz_probe_split_b(b.ctypes.data_as(c_double_p), n, k, n, n_pad, kp, eB.ctypes.data_as(c_int_p), TB.ctypes.data_as(c_int8_p))
    want = np.zeros_like(TB)
    ks_, cols = np.meshgrid(np.arange(k), np.arange(n), indexing="ij")
    for t in range(S):
        off = np.vectorize(_layout_offset)(cols, ks_, t, kp // 32)  # B^T: the store row is the output column
        want[off] = digits[t].astype(np.int8)
    assert np.array_equal(TB, want)
    # The paired MMA reads digits u and u+1 of one k step as ONE K-major operand of 256 rows: row 128 + r of that operand
    # (8-row groups 256 B apart) must be row r of digit u+1.
    tile, ks, u = 1, 2, 3
    base = ((tile * (kp // 32) + ks) * S + u) * 4096
    for r256 in (0, 127, 128, 200, 255):
        for kb in (0, 15, 16, 31):
            addr = base + (r256 // 8) * 256 + (kb // 16) * 128 + (r256 % 8) * 16 + kb % 16
            digit, r = (u, r256) if r256 < 128 else (u + 1, r256 - 128)
            col, kk = tile * 128 + r, ks * 32 + kb
            expect = digits[digit][kk, col] if (kk < k and col < n) else 0
            assert TB[addr] == expect
