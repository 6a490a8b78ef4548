"""
ozaki_model.py — numpy/Python-int restatement of the arithmetic of csrc/ozaki_split.cuh and
csrc/ozaki_gemm.cuh.  TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).

The tcgen05 kernel computes C += A*B from exact integer pieces, so its arithmetic can be restated
without any tensor core:
  split    e = 1 + floor(log2(max |x|)) per row of A / column of B of a K chunk (8192);
           q = rint(x * 2^(54-e)); q = sum_t d_t 256^(7-t) with balanced digits d_t in [-128, 127]
  products P_g = sum_{t+u=g} A_t @ B_u  in exact integers (int32 on the GPU), groups g = 2 .. 8
  combine  per K chunk, pass 1 (groups 8..5) and pass 2 (groups 4..2):
           v_p = sum_g P_g * 256^(g_hi-g)        (exact, < 2^53)
           x   = fl( v_2 * 2^32 + v_1 )          (first rounding: joining the two passes)
           C   = fl( C + x * 2^(eA[i] + eB[j] - 60) )                     (second rounding)
gemm_kernel() does exactly that with Python integers, so it reproduces the kernel's result including
its two floating-point roundings per chunk; gemm_balanced() is the same sum written as one expression.  There is no reference
counterpart: the reference computes in native FP64 (src/phpc_gemm.cu:50-55); the model exists to pin
the emulation algorithm itself, next to the oracle that pins the result.
"""
import math

import numpy as np

DIGIT_BITS = 8
BAL_BITS = 54
S_DIGITS = 7
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


# Balanced base-256 digits: the scaled value is rounded to a 54-bit integer Q (|x| * 2^-e * 2^54), Q is written in
# base 256 with digits in [-128, 127] (carry from the least significant end), all digits signed 8-bit.
# 7 digits, 28 digit products, weights 256^-(t+u).
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


def gemm_kernel(a, b, c0=None):
    """C = c0 + a @ b exactly as csrc/ozaki_gemm.cuh computes it: per K chunk two passes of up to four groups (pass 1 = groups
    8..5, pass 2 = groups 4..2), joined with one rounding, then one FP64 addition into C."""
    S, bits = S_DIGITS, BAL_BITS
    m, k = a.shape
    n = b.shape[1]
    c = np.zeros((m, n)) if c0 is None else np.array(c0, dtype=np.float64, copy=True)
    for k0 in range(0, k, KC_MAX):
        ac, bc = a[:, k0:k0 + KC_MAX], b[k0:k0 + KC_MAX, :]
        ea, eb = exponents(ac, 1), exponents(bc, 0)
        da, db = split_digits_balanced(ac, ea, 1, S, bits), split_digits_balanced(bc, eb, 0, S, bits)
        groups = {}
        for g in range(2, S + 2):
            acc = np.zeros((m, n), dtype=object)
            for t in range(max(1, g - S), min(S, g - 1) + 1):
                acc = acc + (da[t - 1].astype(object) @ db[g - t - 1].astype(object))
            assert max(abs(int(v)) for v in acc.ravel()) < (1 << 31)  # fits the int32 TMEM accumulator
            groups[g] = acc
        parts = []
        for ps in range(2):
            g_hi = S + 1 - GROUPS_PER_PASS * ps
            g_lo = max(2, g_hi - GROUPS_PER_PASS + 1)
            v = np.zeros((m, n), dtype=object)
            for g in range(g_lo, g_hi + 1):
                v = v + groups[g] * (1 << (DIGIT_BITS * (g_hi - g)))
            assert max(abs(int(x)) for x in v.ravel()) < (1 << 53)  # each pass is exact in FP64
            parts.append(v)
        # the kernel joins the passes with one rounding, fl(part2 * 2^32 + part1), then adds fl(C + x * 2^(eA + eB - 60))
        total = parts[1] * (1 << (DIGIT_BITS * GROUPS_PER_PASS)) + parts[0]
        scale = -2 * bits + DIGIT_BITS * (2 * S - (S + 1))
        for i in range(m):
            if ea[i] is None:
                continue
            for j in range(n):
                if eb[j] is None or total[i, j] == 0:
                    continue
                c[i, j] = c[i, j] + math.ldexp(float(total[i, j]), ea[i] + eb[j] + scale)
    return c
