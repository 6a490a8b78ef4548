"""The tcgen05 (Ozaki) arithmetic restated in exact integer arithmetic on the CPU (oracle/ozaki_model.py): the balanced
digit split is exact up to its one rounding at 54 bits, 28 digit products reach FP64 accuracy, and the reference's own input
comes out bit-exact.  The GPU test compares the kernel with the model bit for bit."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import ozaki_model as om


def test_balanced_digit_split_is_exact_up_to_its_rounding(oracle):
    a = oracle.fill(6, 40, kind=1, seed=3) * np.ldexp(1.0, np.arange(6) * 17 - 40)[:, None]
    a[2, :] = 0.0
    exps = om.exponents(a, 1)
    digits = om.split_digits_balanced(a, exps, 1, 7)
    assert exps[2] is None and all(np.all(d[2] == 0) for d in digits)
    for d in digits:
        assert d.min() >= -128 and d.max() <= 127  # fits a signed int8
    for i in (0, 1, 3, 5):
        for j in (0, 7, 39):
            q = sum(int(d[i, j]) * 256 ** (6 - t) for t, d in enumerate(digits))
            x = Fraction(float(a[i, j])) * Fraction(2) ** (54 - exps[i])
            assert abs(Fraction(q) - x) <= Fraction(1, 2)  # q = rint(x * 2^(54 - e))
        assert abs(a[i]).max() < 2.0 ** exps[i] and abs(a[i]).max() >= 2.0 ** (exps[i] - 1)


@pytest.mark.parametrize("m,k,n", [(5, 7, 3), (33, 129, 20), (16, 300, 24)])
def test_model_reaches_fp64_accuracy(oracle, m, k, n):
    a = oracle.fill(m, k, kind=1, seed=5)
    b = oracle.fill(k, n, kind=1, seed=6)
    c0 = oracle.fill(m, n, kind=1, seed=7)
    want = oracle.gemm_block(a, b, c0)
    for got in (om.gemm_kernel(a, b, c0), om.gemm_balanced(a, b, c0)):
        assert oracle.rel_frobenius(got, want) <= 2e-15
        bound = 4.0 * np.sqrt(k) * 2.0 ** -53 * (np.abs(a) @ np.abs(b)) + 4 * 2.0 ** -53 * np.abs(want)
        assert np.all(np.abs(got - want) <= bound)


def test_model_is_bit_exact_on_the_reference_fill(oracle):
    for n in (16, 48):
        a = oracle.fill(n, n, kind=0)
        assert np.array_equal(om.gemm_kernel(a, a), oracle.index_fill_exact(n))
        assert np.array_equal(om.gemm_kernel(a, a), oracle.gemm_iterative(a, a))


def test_rows_and_columns_of_any_magnitude(oracle):
    a = oracle.fill(12, 200, kind=1, seed=14) * np.ldexp(1.0, np.arange(12) * 40 - 250)[:, None]  # rows 2^-250 .. 2^190
    b = oracle.fill(200, 10, kind=1, seed=15)
    assert oracle.rel_frobenius(om.gemm_kernel(a, b), oracle.gemm_block(a, b)) <= 2e-15


@pytest.mark.gpu
def test_kernel_matches_the_integer_model_bit_for_bit(gpu, capi, oracle):
    """Every operation of the kernel is exact except its two FP64 additions per K chunk, which the model performs in the same
    order: the results agree to the last bit (also across K chunks: k = 9000 is two chunks)."""
    from tests.test_gpu_parity import _device_gemm_from_numpy

    for (m, k, n) in ((40, 70, 30), (33, 129, 65), (130, 100, 257), (6, 9000, 5)):
        a = oracle.fill(m, k, kind=1, seed=61)
        b = oracle.fill(k, n, kind=1, seed=62)
        c0 = oracle.fill(m, n, kind=1, seed=63)
        got, _ = _device_gemm_from_numpy(capi, gpu, a, b, c0, "ozaki")
        want = om.gemm_kernel(a, b, c0)
        assert np.array_equal(got, want), (m, k, n, oracle.rel_frobenius(got, want))


def test_error_is_normwise_per_row_and_column_not_componentwise(oracle):
    """What the digit scheme ITSELF guarantees and what it does not ("Error model").  Digits are cut below the ROW
    (column) maximum, so an element 2^-66 of its row's maximum has no digit left: a C element that consists only of such
    products comes out as 0.  The absolute error still obeys the normwise bound k * 2^-53 * max|A_i| * max|B_j|, but the
    componentwise bound of a native FP64 dot product, ~ k * 2^-53 * sum|a||b|, does not hold for such an element.  This is why
    the launcher's guard hands K chunks whose rows / columns span more than 2^40 to the native-FP64 kernel
    (tests/test_gpu_parity.py::test_tcgen05_path_keeps_fp64_semantics_for_special_values)."""
    k = 64
    a = np.full((2, k), 1e-20)
    a[0, 0] = 1.0
    a[1, :] = np.linspace(0.5, 1.0, k)
    b = np.ones((k, 2))
    b[0, 0] = 0.0
    exact = np.array([[float(sum(Fraction(float(a[i, q])) * Fraction(float(b[q, j])) for q in range(k))) for j in range(2)] for i in range(2)])
    got = om.gemm_kernel(a, b)
    assert exact[0, 0] > 0 and got[0, 0] == 0.0  # 63 products of 1e-20 * 1: below the last digit of a row whose maximum is 1
    norm_bound = k * 2.0 ** -53 * np.outer(np.abs(a).max(axis=1), np.abs(b).max(axis=0)) + 2.0 ** -51 * np.abs(exact)
    assert np.all(np.abs(got - exact) <= norm_bound)
    comp_bound = 4.0 * np.sqrt(k) * 2.0 ** -53 * (np.abs(a) @ np.abs(b))
    assert abs(got[0, 0] - exact[0, 0]) > comp_bound[0, 0]
    assert np.all(np.abs(got[1] - exact[1]) <= comp_bound[1])  # rows without that dynamic range meet the componentwise bound too
