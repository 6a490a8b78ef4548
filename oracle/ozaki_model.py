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
