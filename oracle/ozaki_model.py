"""
ozaki_model.py — numpy/Python-int restatement of the arithmetic of csrc/ozaki_split.cuh and
csrc/ozaki_gemm.cuh.  TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).

The tcgen05 kernel computes C += A*B from exact integer pieces, so its arithmetic can be restated
without any tensor core:
  split    e = 1 + floor(log2(max |x|)) per row of A / column of B;  r = x * 2^-e;
           digit_t = trunc(r * 128), r = r * 128 - digit_t          (S times, all exact in FP64)
  products P_g = sum_{t+u=g} A_t @ B_u  in exact integers (int32 on the GPU, K <= 8192 per chunk)
  combine  per K chunk and per pass of four groups  g_hi .. g_lo  (least significant pass first):
           v = sum_g P_g * 2^(7*(g_hi-g))   (exact, < 2^53)
           C = fl( C + v * 2^(eA[i] + eB[j] - 7*g_hi) )              (one rounding per pass)
This model does exactly that with Python integers, so it reproduces the kernel's result including
the order of its (two per chunk) floating-point roundings.  There is no reference counterpart: the
reference computes in native FP64 (src/phpc_gemm.cu:50-55); the model exists to pin the emulation
algorithm itself, next to the oracle that pins the result.
"""
import math

import numpy as np

DIGIT_BITS = 7
KC_MAX = 8192          # K chunk of phpc_launch_ozaki
GROUPS_PER_PASS = 4


def exponents(x, axis):
    """e with |x| < 2^e along `axis` (per row of A: axis=1; per column of B: axis=0); None for all-zero."""
    mx = np.max(np.abs(x), axis=axis)
    out = []
    for v in mx:
        if v == 0.0:
            out.append(None)
        else:
            m, e = math.frexp(float(v))  # v = m * 2^e, 0.5 <= m < 1  ->  |x| < 2^e
            out.append(e)
    return out


def split_digits(x, exps, axis, S):
    """Digit tensors d[t] (int64, same shape as x) with x = 2^e * (sum_t d[t] 2^(-7(t+1)) + rest)."""
    r = np.array(x, dtype=np.float64, copy=True)
    scale = np.array([0.0 if e is None else math.ldexp(1.0, -e) for e in exps])
    r = r * (scale[:, None] if axis == 1 else scale[None, :])  # exact: power-of-two scaling
    digits = []
    for _ in range(S):
        s = r * 128.0
        d = np.trunc(s)
        r = s - d
        digits.append(d.astype(np.int64))
    return digits, r


def gemm(a, b, c0=None, S=8):
    """C = c0 + a @ b exactly as the tcgen05 kernel computes it (returns float64)."""
    m, k = a.shape
    n = b.shape[1]
    c = np.zeros((m, n)) if c0 is None else np.array(c0, dtype=np.float64, copy=True)
    for k0 in range(0, k, KC_MAX):
        ac, bc = a[:, k0:k0 + KC_MAX], b[k0:k0 + KC_MAX, :]
        ea, eb = exponents(ac, 1), exponents(bc, 0)
        da, _ = split_digits(ac, ea, 1, S)
        db, _ = split_digits(bc, eb, 0, S)
        groups = {}
        for g in range(2, S + 2):
            acc = np.zeros((m, n), dtype=object)
            for t in range(max(1, g - S), min(S, g - 1) + 1):
                acc = acc + (da[t - 1].astype(object) @ db[g - t - 1].astype(object))
            groups[g] = acc
        npass = (S + GROUPS_PER_PASS - 1) // GROUPS_PER_PASS
        for ps in range(npass):
            g_hi = S + 1 - GROUPS_PER_PASS * ps
            g_lo = max(2, g_hi - GROUPS_PER_PASS + 1)
            v = np.zeros((m, n), dtype=object)
            for g in range(g_lo, g_hi + 1):
                v = v + groups[g] * (1 << (DIGIT_BITS * (g_hi - g)))
            for i in range(m):
                if ea[i] is None:
                    continue
                for j in range(n):
                    if eb[j] is None or v[i, j] == 0:
                        continue
                    assert abs(v[i, j]) < (1 << 53)
                    c[i, j] = c[i, j] + math.ldexp(float(v[i, j]), ea[i] + eb[j] - DIGIT_BITS * g_hi)
    return c


# ----------------------------------------------------------------------------------------------
# Planned digit scheme — modelled here before it is written in CUDA:
# first digit signed (floor; 7 bits + sign), further digits UNSIGNED 8 bits (the remainder after a
# floor is non-negative).  kind::i8 accepts s8 / u8 per operand, so only the split kernels and the
# instruction descriptors change.  7 digits then carry 7 + 6*8 = 55 bits with 28 digit products, but the
# truncation error becomes biased (tests/test_ozaki_model.py: ~4e-15 instead of ~1e-15).
# ----------------------------------------------------------------------------------------------
def split_digits_mixed(x, exps, axis, S):
    r = np.array(x, dtype=np.float64, copy=True)
    scale = np.array([0.0 if e is None else math.ldexp(1.0, -e) for e in exps])
    r = r * (scale[:, None] if axis == 1 else scale[None, :])
    digits = []
    for t in range(S):
        s = r * (128.0 if t == 0 else 256.0)
        d = np.floor(s)
        r = s - d  # in [0, 1)
        digits.append(d.astype(np.int64))
    return digits


def gemm_mixed(a, b, c0=None, S=7, kc_max=4096):
    """C = c0 + a @ b with the signed-first / unsigned-rest digits; pair (t,u) weighs 2^-(8(t+u)-2)."""
    m, k = a.shape
    n = b.shape[1]
    c = np.zeros((m, n)) if c0 is None else np.array(c0, dtype=np.float64, copy=True)
    for k0 in range(0, k, kc_max):
        ac, bc = a[:, k0:k0 + kc_max], b[k0:k0 + kc_max, :]
        ea, eb = exponents(ac, 1), exponents(bc, 0)
        da, db = split_digits_mixed(ac, ea, 1, S), split_digits_mixed(bc, eb, 0, S)
        assert da[0].min() >= -128 and da[0].max() <= 127 and all(0 <= d.min() and d.max() <= 255 for d in da[1:])
        total = np.zeros((m, n), dtype=object)  # exact sum of all kept groups, in units of 2^-(8(S+1)-2)
        for g in range(2, S + 2):
            acc = np.zeros((m, n), dtype=object)
            for t in range(max(1, g - S), min(S, g - 1) + 1):
                acc = acc + (da[t - 1].astype(object) @ db[g - t - 1].astype(object))
            assert max(abs(int(v)) for v in acc.ravel()) < (1 << 31)  # fits the int32 TMEM accumulator
            total = total + acc * (1 << (8 * (S + 1 - g)))
        for i in range(m):
            for j in range(n):
                if ea[i] is None or eb[j] is None or total[i, j] == 0:
                    continue
                c[i, j] = c[i, j] + math.ldexp(float(total[i, j]), ea[i] + eb[j] - (8 * (S + 1) - 2))  # float(int) rounds once
    return c


# Balanced base-256 digits (the candidate that keeps the cancellation): the scaled value is rounded to a
# 54-bit integer Q (|x| * 2^-e * 2^54), Q is written in base 256 with digits in [-128, 127] (carry from
# the least significant end), all digits signed 8-bit.  7 digits, 28 digit products, weights 256^-(t+u).
def split_digits_balanced(x, exps, axis, S=7, bits=54):
    r = np.array(x, dtype=np.float64, copy=True)
    scale = np.array([0.0 if e is None else math.ldexp(1.0, bits - e) for e in exps])
    q = np.rint(r * (scale[:, None] if axis == 1 else scale[None, :])).astype(np.int64)  # |q| <= 2^54, exact scaling + one rounding
    digits = []
    for _ in range(S):
        d = ((q + 128) % 256) - 128
        q = (q - d) // 256
        digits.append(d)
    assert np.all(q == 0)
    return digits[::-1]  # most significant first: x ~ 2^(e-bits) * sum_t d_t 256^(S-1-t)


def gemm_balanced(a, b, c0=None, S=7, kc_max=8192, bits=54):
    m, k = a.shape
    n = b.shape[1]
    c = np.zeros((m, n)) if c0 is None else np.array(c0, dtype=np.float64, copy=True)
    for k0 in range(0, k, kc_max):
        ac, bc = a[:, k0:k0 + kc_max], b[k0:k0 + kc_max, :]
        ea, eb = exponents(ac, 1), exponents(bc, 0)
        da, db = split_digits_balanced(ac, ea, 1, S, bits), split_digits_balanced(bc, eb, 0, S, bits)
        total = np.zeros((m, n), dtype=object)
        for g in range(2, S + 2):  # keep pairs with t+u <= S+1 (1-based, most significant first)
            acc = np.zeros((m, n), dtype=object)
            for t in range(max(1, g - S), min(S, g - 1) + 1):
                acc = acc + (da[t - 1].astype(object) @ db[g - t - 1].astype(object))
            assert max(abs(int(v)) for v in acc.ravel()) < (1 << 31)
            total = total + acc * (1 << (8 * (S + 1 - g)))
        # pair (t,u) weighs 256^(2S-t-u); `total` is in units of 256^(S-1) (the dropped groups are below it)
        for i in range(m):
            for j in range(n):
                if ea[i] is None or eb[j] is None or total[i, j] == 0:
                    continue
                c[i, j] = c[i, j] + math.ldexp(float(total[i, j]), ea[i] + eb[j] - 2 * bits + 8 * (S - 1))
    return c
