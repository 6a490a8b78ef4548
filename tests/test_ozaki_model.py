"""The Ozaki (tcgen05) arithmetic restated in exact integer arithmetic on the CPU (oracle/ozaki_model.py):
the digit split is error free, the truncated digit products reach FP64 accuracy with 8 digits, and the
reference's own input comes out bit-exact.  The GPU test compares the kernel with the model."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import ozaki_model as om


def test_digit_split_is_error_free(oracle):
    a = oracle.fill(6, 40, kind=1, seed=3) * np.ldexp(1.0, np.arange(6) * 17 - 40)[:, None]
    a[2, :] = 0.0
    exps = om.exponents(a, 1)
    digits, rest = om.split_digits(a, exps, 1, 8)
    assert exps[2] is None and all(np.all(d[2] == 0) for d in digits)
    for d in digits:
        assert d.min() >= -127 and d.max() <= 127  # fits a signed int8
    for i in (0, 1, 3, 5):
        for j in (0, 7, 39):
            x = Fraction(float(a[i, j]))
            recon = sum(Fraction(int(d[i, j]), 1 << (7 * (t + 1))) for t, d in enumerate(digits)) + Fraction(float(rest[i, j])) / (1 << 56)
            assert recon * Fraction(2) ** exps[i] == x  # exact, not approximately
        assert abs(a[i]).max() < 2.0 ** exps[i] and abs(a[i]).max() >= 2.0 ** (exps[i] - 1)


@pytest.mark.parametrize("m,k,n", [(5, 7, 3), (33, 129, 20), (16, 300, 24)])
def test_model_reaches_fp64_accuracy_with_8_digits(oracle, m, k, n):
    a = oracle.fill(m, k, kind=1, seed=5)
    b = oracle.fill(k, n, kind=1, seed=6)
    c0 = oracle.fill(m, n, kind=1, seed=7)
    want = oracle.gemm_block(a, b, c0)
    got = om.gemm(a, b, c0, S=8)
    assert oracle.rel_frobenius(got, want) <= 1e-15
    bound = 4.0 * np.sqrt(k) * 2.0 ** -53 * (np.abs(a) @ np.abs(b)) + 4 * 2.0 ** -53 * np.abs(want)
    assert np.all(np.abs(got - want) <= bound)


def test_model_is_bit_exact_on_the_reference_fill(oracle):
    for n in (16, 48):
        a = oracle.fill(n, n, kind=0)
        assert np.array_equal(om.gemm(a, a, S=8), oracle.index_fill_exact(n))
        assert np.array_equal(om.gemm(a, a, S=8), oracle.gemm_iterative(a, a))


def test_digit_count_sets_the_accuracy(oracle):
    a = oracle.fill(24, 256, kind=1, seed=1)
    b = oracle.fill(256, 24, kind=1, seed=2)
    want = oracle.gemm_block(a, b)
    errs = {S: oracle.rel_frobenius(om.gemm(a, b, S=S), want) for S in (4, 6, 7, 8)}
    assert errs[4] > errs[6] > errs[7] > errs[8]
    assert errs[4] < 1e-6 and errs[6] < 1e-10 and errs[7] < 1e-12 and errs[8] < 1e-15


@pytest.mark.gpu
def test_kernel_matches_the_integer_model(gpu, capi, oracle):
    """Every operation of the kernel is exact except its two FP64 additions per K chunk, which the
    model performs in the same order: the results should agree to the last bit (reported), and
    must agree to 1e-15 (asserted)."""
    from tests.test_gpu_parity import _device_gemm_from_numpy

    for (m, k, n), slices in (((40, 70, 30), 8), ((33, 129, 65), 8), ((20, 100, 17), 6)):
        a = oracle.fill(m, k, kind=1, seed=61)
        b = oracle.fill(k, n, kind=1, seed=62)
        c0 = oracle.fill(m, n, kind=1, seed=63)
        got, _ = _device_gemm_from_numpy(capi, gpu, a, b, c0, "ozaki", slices)
        want = om.gemm(a, b, c0, S=slices)
        print(f"ozaki kernel vs integer model {m}x{k}x{n} S={slices}: bit-equal={np.array_equal(got, want)}")
        assert oracle.rel_frobenius(got, want) <= 1e-15


def test_planned_mixed_signedness_digits_trade_accuracy_for_fewer_products(oracle):
    """DESIGN.md section 8 item 1, modelled before it is written in CUDA: signed first digit, unsigned
    8-bit digits after it.  7 digits = 28 digit products (instead of 36) stay inside the 1e-14 tolerance
    but are NOT as accurate as 8 signed digits: unsigned digits are never negative, so the dropped
    low-order groups no longer cancel (measured here: ~4e-15 vs ~6e-16 with 8 mixed digits)."""
    a = oracle.fill(20, 300, kind=1, seed=11)
    b = oracle.fill(300, 18, kind=1, seed=12)
    c0 = oracle.fill(20, 18, kind=1, seed=13)
    want = oracle.gemm_block(a, b, c0)
    e7 = oracle.rel_frobenius(om.gemm_mixed(a, b, c0, S=7), want)
    e8 = oracle.rel_frobenius(om.gemm_mixed(a, b, c0, S=8), want)
    assert e8 <= 1e-15 < e7 <= 1e-14
    n = 32
    ai = oracle.fill(n, n, kind=0)
    assert np.array_equal(om.gemm_mixed(ai, ai, S=7), oracle.index_fill_exact(n))


def test_balanced_base256_digits_keep_the_accuracy_with_28_products(oracle):
    """The experimental digit scheme of PHPC_OZAKI_DIGITS=balanced (csrc/ozaki_split.cuh, balanced_digits): the value
    rounded to 54 bits below its row/column scale, written in base 256 with digits in [-128, 127].  7 digits = 28
    digit products instead of 36, and because balanced digits keep their sign the dropped low-order groups still
    cancel: the model is as accurate as 8 truncated 7-bit digits (the mixed-signedness variant above is not)."""
    for (m, k, n) in ((20, 300, 18), (9, 1000, 12)):
        a = oracle.fill(m, k, kind=1, seed=11)
        b = oracle.fill(k, n, kind=1, seed=12)
        c0 = oracle.fill(m, n, kind=1, seed=13)
        want = oracle.gemm_block(a, b, c0)
        e_bal = oracle.rel_frobenius(om.gemm_balanced(a, b, c0, S=7), want)
        e_8 = oracle.rel_frobenius(om.gemm(a, b, c0, S=8), want)
        assert e_bal <= 2e-15 and e_bal <= 1.5 * e_8
    a = oracle.fill(12, 200, kind=1, seed=14) * np.ldexp(1.0, np.arange(12) * 40 - 250)[:, None]  # rows 2^-250 .. 2^190
    b = oracle.fill(200, 10, kind=1, seed=15)
    assert oracle.rel_frobenius(om.gemm_balanced(a, b, S=7), oracle.gemm_block(a, b)) <= 2e-15
    ai = oracle.fill(32, 32, kind=0)
    assert np.array_equal(om.gemm_balanced(ai, ai, S=7), oracle.index_fill_exact(32))
    d = om.split_digits_balanced(a, om.exponents(a, 1), 1, 7)
    assert all(x.min() >= -128 and x.max() <= 127 for x in d)


def test_error_is_normwise_per_row_and_column_not_componentwise(oracle):
    """What the fixed-digit scheme guarantees and what it does not ("Error model").  Digits are cut
    below the ROW (column) maximum, so an element 2^-66 of its row's maximum has no digit left: a C element that consists
    only of such products comes out as 0.  The absolute error still obeys the normwise bound
    k * 2^-53 * max|A_i| * max|B_j| (far below FP64 rounding of anything else in that row), but the componentwise bound of
    a native FP64 dot product, ~ k * 2^-53 * sum|a||b|, does not hold for such an element.  The native-FP64 kernel
    (PHPC_GEMM=dmma) is the path for inputs with that much dynamic range inside single rows/columns."""
    k = 64
    a = np.full((2, k), 1e-20)
    a[0, 0] = 1.0
    a[1, :] = np.linspace(0.5, 1.0, k)
    b = np.ones((k, 2))
    b[0, 0] = 0.0
    exact = np.array([[float(sum(Fraction(float(a[i, q])) * Fraction(float(b[q, j])) for q in range(k))) for j in range(2)] for i in range(2)])
    for got in (om.gemm(a, b, S=8), om.gemm_balanced(a, b, S=7)):
        assert exact[0, 0] > 0 and got[0, 0] == 0.0  # 63 products of 1e-20 * 1: below the last digit of a row whose maximum is 1
        norm_bound = k * 2.0 ** -53 * np.outer(np.abs(a).max(axis=1), np.abs(b).max(axis=0)) + 2.0 ** -51 * np.abs(exact)
        assert np.all(np.abs(got - exact) <= norm_bound)
        comp_bound = 4.0 * np.sqrt(k) * 2.0 ** -53 * (np.abs(a) @ np.abs(b))
        assert abs(got[0, 0] - exact[0, 0]) > comp_bound[0, 0]
        assert np.all(np.abs(got[1] - exact[1]) <= comp_bound[1])  # rows without that dynamic range meet the componentwise bound too
