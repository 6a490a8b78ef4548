"""The oracle against the reference's own outputs: committed fixtures (always) and the
live reference build oracle/_ref (when /root/reference was present at build time)."""
import os
import tempfile

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("n", [32, 64])
@pytest.mark.parametrize("f", [0, 1])
def test_iterative_matches_reference_fixture(oracle, golden, n, f):
    A = oracle.fill(n, n, kind=f, seed=oracle.SEED_A)
    B = oracle.fill(n, n, kind=f, seed=oracle.SEED_B)
    assert np.array_equal(oracle.gemm_iterative(A, B), golden[f"iterative_N{n}_F{f}"])  # bit-exact


@pytest.mark.parametrize("p", [1, 2, 4, 6, 8, 16])
@pytest.mark.parametrize("f", [0, 1])
def test_summa_matches_reference_fixture(oracle, golden, p, f):
    n = 48
    r, c = (int(x) for x in golden[f"summa_N{n}_P{p}_grid"])
    A = oracle.fill(n, n, kind=f, seed=oracle.SEED_A)
    B = oracle.fill(n, n, kind=f, seed=oracle.SEED_B)
    assert np.array_equal(oracle.summa(A, B, r, c), golden[f"summa_N{n}_P{p}_F{f}"])  # bit-exact


def test_index_fill_closed_form_is_what_the_reference_computes(oracle, golden):
    for n in (32, 64):
        assert np.array_equal(oracle.index_fill_exact(n), golden[f"iterative_N{n}_F0"])
    # N = 1024 (BASELINE config 1): iterative order, kernel order and closed form agree bit for bit
    n = 256
    A = oracle.fill(n, n, kind=0)
    exact = oracle.index_fill_exact(n)
    assert np.array_equal(oracle.gemm_iterative(A, A), exact)
    assert np.array_equal(oracle.gemm_block(A, A), exact)
    for r, c in ((1, 2), (2, 2), (2, 4), (4, 2)):
        assert np.array_equal(oracle.summa(A, A, r, c), exact)


def test_config1_n1024_single_rank(oracle):
    """BASELINE config 1: iterative.c vs SUMMA at N=1024, single rank, reference fill."""
    n = 1024
    A = oracle.fill(n, n, kind=0)
    exact = oracle.index_fill_exact(n)
    assert np.array_equal(oracle.summa(A, A, 1, 1), exact)
    if oracle.have_ref():
        assert np.array_equal(oracle.ref_iterative(A, A), exact)
    else:
        assert np.array_equal(oracle.gemm_iterative(A, A), exact)


def test_block_gemm_semantics(oracle):
    """C += sum (kernel rounding), ragged shapes, empty K."""
    rng = np.random.default_rng(0)
    a, b, c0 = rng.standard_normal((7, 5)), rng.standard_normal((5, 3)), rng.standard_normal((7, 3))
    want = c0 + sum(np.outer(a[:, p], b[p, :]) for p in range(5))
    assert np.allclose(oracle.gemm_block(a, b, c0), want, rtol=1e-15, atol=1e-15)
    assert np.array_equal(oracle.gemm_block(np.zeros((4, 0)), np.zeros((0, 6)), np.ones((4, 6))), np.ones((4, 6)))


def test_dot_exact(oracle):
    a = np.array([1e16, 1.0, -1e16, 3.0])
    b = np.ones(4)
    assert oracle.dot_exact(a, b) == 4.0  # naive summation gives 3.0 or 4.0 depending on order; dd is exact


def test_live_reference_when_built(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this box); fixtures cover the pin")
    n = 40
    A = oracle.fill(n, n, kind=1, seed=oracle.SEED_A)
    B = oracle.fill(n, n, kind=1, seed=oracle.SEED_B)
    assert np.array_equal(oracle.gemm_iterative(A, B), oracle.ref_iterative(A, B))
    with tempfile.TemporaryDirectory() as d:
        for p in (2, 4):
            C, (r, c) = oracle.ref_summa_cpu(n, p, 1, d)
            assert np.array_equal(C, oracle.summa(A, B, r, c))
