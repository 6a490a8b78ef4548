"""
oracle.py — ctypes/numpy front end of the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (hpc_multigpu_matrixmult_b200/) never does.

Functions restate the reference's arithmetic (citations in gemm_oracle.c):
  gemm_iterative   reference src/iterative.c:8-13
  gemm_block       reference src/phpc_gemm.cu:6-57 (sum in ascending k, then C += sum)
  summa            reference src/phpc_summa.c:24-122 on an r x c grid
  index_fill_exact closed form of the reference fill A[i]=B[i]=i (src/main.c:85-86)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
REF_DIR = os.path.join(_HERE, "_ref")

_dp = ctypes.POINTER(ctypes.c_double)


def build(with_ref=True):
    """Compile liboracle.so (and oracle/_ref from /root/reference when it is present)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)
    if with_ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(with_ref=False)
        L = ctypes.CDLL(path)
        L.oracle_gemm_iterative.argtypes = [_dp, _dp, _dp, ctypes.c_int]
        L.oracle_gemm_block.argtypes = [_dp, ctypes.c_long, _dp, ctypes.c_long, _dp, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_summa.argtypes = [_dp, _dp, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_find_lcm.argtypes = [ctypes.c_int, ctypes.c_int]
        L.oracle_find_lcm.restype = ctypes.c_int
        L.oracle_summa_owners.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.oracle_index_fill_exact.argtypes = [ctypes.c_longlong] * 3
        L.oracle_index_fill_exact.restype = ctypes.c_double
        L.oracle_index_fill_exact_block.argtypes = [_dp, ctypes.c_long, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_long, ctypes.c_long, ctypes.c_longlong]
        L.oracle_fill.argtypes = [_dp, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_ulonglong]
        L.oracle_dot_dd.argtypes = [_dp, ctypes.c_long, _dp, ctypes.c_long, ctypes.c_long]
        L.oracle_dot_dd.restype = ctypes.c_double
        _LIB = L
    return _LIB


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


FILL_INDEX, FILL_SEEDED = 0, 1
SEED_A, SEED_B = 1234, 5678


def fill(rows, cols, row0=0, col0=0, N=None, kind=FILL_INDEX, seed=SEED_A):
    """rows x cols window at global (row0, col0) of the N x N synthetic matrix."""
    N = cols if N is None else N
    out = np.empty((rows, cols), dtype=np.float64)
    lib().oracle_fill(_p(out), cols, rows, cols, row0, col0, N, kind, seed)
    return out


def gemm_iterative(A, B, C=None):
    n = A.shape[0]
    C = np.zeros((n, n)) if C is None else np.ascontiguousarray(C, dtype=np.float64).copy()
    lib().oracle_gemm_iterative(_p(np.ascontiguousarray(A)), _p(np.ascontiguousarray(B)), _p(C), n)
    return C


def gemm_block(A, B, C=None):
    m, k = A.shape
    k2, n = B.shape
    assert k == k2
    C = np.zeros((m, n)) if C is None else np.ascontiguousarray(C, dtype=np.float64).copy()
    A = np.ascontiguousarray(A)
    B = np.ascontiguousarray(B)
    lib().oracle_gemm_block(_p(A), k, _p(B), n, _p(C), n, m, k, n)
    return C


def summa(A, B, r, c, C=None):
    N = A.shape[0]
    assert N % r == 0 and N % c == 0
    C = np.zeros((N, N)) if C is None else np.ascontiguousarray(C, dtype=np.float64).copy()
    lib().oracle_summa(_p(np.ascontiguousarray(A)), _p(np.ascontiguousarray(B)), _p(C), N, r, c)
    return C


def find_lcm(a, b):
    return lib().oracle_find_lcm(a, b)


def summa_owners(r, c):
    l = find_lcm(r, c)
    oc = (ctypes.c_int * l)()
    orow = (ctypes.c_int * l)()
    lib().oracle_summa_owners(r, c, oc, orow)
    return list(oc), list(orow)


def index_fill_exact(N, row0=0, col0=0, rows=None, cols=None):
    rows = N if rows is None else rows
    cols = N if cols is None else cols
    out = np.empty((rows, cols), dtype=np.float64)
    lib().oracle_index_fill_exact_block(_p(out), cols, row0, col0, rows, cols, N)
    return out


def dot_exact(a_row, b_col):
    """Correctly rounded (double-double) dot product of two 1-D float64 arrays."""
    a = np.ascontiguousarray(a_row, dtype=np.float64)
    b = np.ascontiguousarray(b_col, dtype=np.float64)
    return lib().oracle_dot_dd(_p(a), 1, _p(b), 1, a.shape[0])


def rel_frobenius(C, Cref):
    d = np.linalg.norm((C - Cref).ravel())
    n = np.linalg.norm(Cref.ravel())
    return float(d / n) if n > 0 else float(d)


# ---- the reference itself, when oracle/_ref was built in this container ----------
def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_iterative.so"))


def ref_iterative(A, B, C=None):
    """The reference's own phpc_gemm_iterative (src/iterative.c compiled unchanged)."""
    L = ctypes.CDLL(os.path.join(REF_DIR, "libref_iterative.so"))
    L.phpc_gemm_iterative.argtypes = [_dp, _dp, _dp, ctypes.c_int]
    n = A.shape[0]
    C = np.zeros((n, n)) if C is None else np.ascontiguousarray(C, dtype=np.float64).copy()
    L.phpc_gemm_iterative(_p(np.ascontiguousarray(A)), _p(np.ascontiguousarray(B)), _p(C), n)
    return C


def ref_summa_cpu(N, nranks, fill_kind, workdir):
    """The reference's own phpc_summa.c under the MPI shim with a CPU gemm_t plugin."""
    root = os.path.dirname(_HERE)
    out = os.path.join(workdir, f"ref_summa_N{N}_P{nranks}_F{fill_kind}.bin")
    cmd = [os.path.join(root, "bin", "mpirun"), "-n", str(nranks), os.path.join(REF_DIR, "ref_summa_cpu.out"), str(N), str(fill_kind), out]
    res = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=300)
    n, size, d0, d1 = (int(x) for x in res.stdout.strip().split(",")[:4])  # a fifth field is the SUMMA wall time
    return np.fromfile(out, dtype=np.float64).reshape(N, N), (d0, d1)
