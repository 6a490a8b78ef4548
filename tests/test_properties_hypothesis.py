"""Property tests (hypothesis) of the pieces that are pure arithmetic: the digit extraction the split
kernels of the tcgen05 path run (host build of the same lines, tests/csrc/oz_host_probe.cu) on arbitrary doubles, and the operation list of
the band-pipelined host call on arbitrary shapes."""
import ctypes
import math
from fractions import Fraction

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from tests.test_host_plan import _edges, _reach, _regions
from tests.test_ozaki_split_host import c_double_p, c_int8_p, probe  # noqa: F401  (fixture)

finite = st.floats(allow_nan=False, allow_infinity=False, allow_subnormal=True, width=64)


def _exp_above(x):
    """Smallest e with |x| < 2^e, as exp_above() of csrc/ozaki_split.cuh (denormals share the smallest normal exponent)."""
    if x == 0.0:
        return None
    m, e = math.frexp(abs(x))  # |x| = m 2^e, 0.5 <= m < 1
    return max(e, -1022)


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(x=finite, slack=st.integers(min_value=0, max_value=40))
def test_balanced_digits_round_to_54_bits(probe, x, slack):
    e = _exp_above(x)
    if e is None or e + slack > 1023:
        return
    e += slack
    out = np.zeros(7, dtype=np.int8)
    arr = np.array([x])
    probe.oz_probe_digits(arr.ctypes.data_as(c_double_p), 1, e, out.ctypes.data_as(c_int8_p))
    q = sum(int(d) * 256 ** (6 - t) for t, d in enumerate(out))
    exact = Fraction(x) * Fraction(2) ** (54 - e)
    assert abs(q - exact) <= Fraction(1, 2)                # correctly rounded to 54 bits below the scale
    assert abs(q) <= 2 ** 54
    if abs(exact.denominator) == 1:                        # representable: no rounding at all
        assert q == exact


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(m=st.integers(1, 70000), nsteps=st.integers(1, 9), bands=st.integers(1, 12), align=st.sampled_from([1, 8, 128]))
def test_host_plan_is_race_free_for_any_shape(capi, m, nsteps, bands, align):
    ops = capi.host_plan(m, nsteps, bands, align)
    before = _reach(_edges(ops))
    touched = [_regions(op, capi) for op in ops]
    for i in range(len(ops)):
        ri, wi = touched[i]
        for j in range(i):
            rj, wj = touched[j]
            if (wi & (rj | wj)) or (wj & ri):
                assert before[i] >> j & 1
    rows = sorted({(op.row0, op.rows) for op in ops if op.kind == capi.HOP_UPLOAD_C})
    assert rows[0][0] == 0 and sum(r for _, r in rows) == m and len(rows) <= bands
    assert all(r0 % align == 0 for r0, _ in rows)
