"""
capi.py — ctypes binding of the C-ABI (include/phpc_gemm.cuh, phpc_summa.h,
phpc_b200.h, utils.h and the MPI shim).  Python is NOT the product's host language
(the reference's host code is C: hpc_multigpu_matrixmult_b200/csrc/main.c is the
drop-in driver); this binding exists so tests/ and bench.py can call exactly the
symbols a C caller links against, with the reference's names and argument meaning.

There is no fallback: if the shared libraries are missing the import raises, and
every compute entry point aborts the process when no sm_100a GPU is present.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libphpc_b200.so")
MPI_LIB_PATH = os.path.join(LIB_DIR, "libphpcmpi.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)

BACKEND_DMMA, BACKEND_CUBLAS, BACKEND_OZAKI = 0, 1, 2
FILL_INDEX, FILL_SEEDED = 0, 1
SEED_A, SEED_B = 1234, 5678

MPI_COMM_WORLD = 0
MPI_BYTE, MPI_CHAR, MPI_INT, MPI_FLOAT, MPI_DOUBLE = 1, 2, 3, 4, 5
MPI_SUM, MPI_MAX, MPI_MIN = 1, 2, 3


class SummaStep(ctypes.Structure):
    _fields_ = [
        ("panel", ctypes.c_int),
        ("a_root", ctypes.c_int),
        ("b_root", ctypes.c_int),
        ("k0", ctypes.c_longlong),
        ("width", ctypes.c_int),
        ("own_a", ctypes.c_int),
        ("own_b", ctypes.c_int),
        ("a_off", ctypes.c_longlong),
        ("b_off", ctypes.c_longlong),
    ]


class HostOp(ctypes.Structure):
    """phpc_host_op of include/phpc_summa.h: one operation of the band-pipelined host-sourced run."""
    _fields_ = [
        ("kind", ctypes.c_int),
        ("stream", ctypes.c_int),
        ("band", ctypes.c_int),
        ("step", ctypes.c_int),
        ("row0", ctypes.c_int),
        ("rows", ctypes.c_int),
        ("ndeps", ctypes.c_int),
        ("deps", ctypes.c_int * 3),
    ]


HOP_UPLOAD_C, HOP_UPLOAD_A, HOP_UPLOAD_B, HOP_GEMM, HOP_DOWNLOAD_C, HOP_ZERO_C, HOP_ADD_C = range(7)


class SummaStats(ctypes.Structure):
    _fields_ = [
        ("total_ms", ctypes.c_float),
        ("gemm_ms", ctypes.c_float),
        ("exposed_ms", ctypes.c_float),
        ("steps", ctypes.c_int),
        ("launches", ctypes.c_int),
        ("broadcasts", ctypes.c_int),
        ("bytes_received", ctypes.c_longlong),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None
_mpi = None


def load():
    """dlopen the MPI shim and the CUDA library (in-tree, built by `make`).

    Python hosts that also use torch must `import torch` BEFORE the first call: libphpc_b200.so needs libnccl.so.2 and
    binds the system's (2.27) when nothing is loaded yet, after which torch's libtorch_cuda.so, which wants the newer
    libnccl.so.2 bundled in its wheel (2.28, same soname), fails to import.  With torch first, the library binds torch's
    copy, which is how bench.py runs.  C hosts (bin/main.out) are not affected."""
    global _lib, _mpi
    if _lib is not None:
        return _lib
    for p in (MPI_LIB_PATH, LIB_PATH):
        if not os.path.exists(p):
            raise ImportError(
                f"{p} is missing: build it with `make -C {_HERE}` (or __graft_entry__.build()); " "there is no Python/CPU fallback for the CUDA path"
            )
    _mpi = ctypes.CDLL(MPI_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)

    gemm_args = [c_double_p, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p, ctypes.c_int] + [ctypes.c_int] * 7 + [c_float_p]
    for name in ("phpc_gemm_cuda", "phpc_gemm_cublas"):
        getattr(L, name).argtypes = gemm_args
        getattr(L, name).restype = None
    L.phpc_gemm_summa_cuda.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p] + [ctypes.c_int] * 5 + [c_float_p]
    L.phpc_gemm_summa_cuda.restype = None
    L.phpc_gemm_summa_cublas.argtypes = [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int, c_float_p]
    L.phpc_gemm_summa_cublas.restype = None

    L.phpc_b200_version.restype = ctypes.c_int
    L.phpc_b200_device_count.restype = ctypes.c_int
    L.phpc_b200_set_device.argtypes = [ctypes.c_int]
    L.phpc_b200_sm_count.restype = ctypes.c_int
    L.phpc_device_malloc.argtypes = [ctypes.c_size_t]
    L.phpc_device_malloc.restype = ctypes.c_void_p
    L.phpc_device_free.argtypes = [ctypes.c_void_p]
    L.phpc_host_malloc_pinned.argtypes = [ctypes.c_size_t]
    L.phpc_host_malloc_pinned.restype = ctypes.c_void_p
    L.phpc_host_free_pinned.argtypes = [ctypes.c_void_p]
    L.phpc_host_malloc_shared.argtypes = [ctypes.c_size_t]
    L.phpc_host_malloc_shared.restype = ctypes.c_void_p
    L.phpc_host_free_shared.argtypes = [ctypes.c_void_p]
    L.phpc_host_register.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.phpc_host_unregister.argtypes = [ctypes.c_void_p]
    L.phpc_device_memset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
    L.phpc_copy2d_to_host.argtypes = [c_double_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong]
    L.phpc_copy2d_to_device.argtypes = [ctypes.c_void_p, ctypes.c_longlong, c_double_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong]
    dev_gemm = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_int] * 3
    L.phpc_gemm_device.argtypes = dev_gemm + [ctypes.c_int, ctypes.c_void_p]
    L.phpc_gemm_device.restype = ctypes.c_int
    L.phpc_gemm_device_ozaki.argtypes = dev_gemm + [ctypes.c_void_p]
    L.phpc_gemm_device_ozaki.restype = ctypes.c_int
    L.phpc_ozaki_fallback_chunks.restype = ctypes.c_longlong
    L.phpc_default_backend.restype = ctypes.c_int
    L.phpc_gemm_device_cublas.argtypes = dev_gemm + [ctypes.c_void_p]
    L.phpc_gemm_device_cublas.restype = None
    L.phpc_gemm_device_timed.argtypes = dev_gemm + [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.phpc_gemm_device_timed.restype = ctypes.c_float
    fill_args = [ctypes.c_longlong] * 6 + [ctypes.c_int, ctypes.c_ulonglong]
    L.phpc_fill_device.argtypes = [ctypes.c_void_p] + fill_args + [ctypes.c_void_p]
    L.phpc_fill_host.argtypes = [c_double_p] + fill_args

    L.phpc_summa_schedule.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(SummaStep), ctypes.c_int, c_int_p, c_int_p]
    L.phpc_summa_schedule.restype = ctypes.c_int
    L.phpc_host_plan.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(HostOp), ctypes.c_int]
    L.phpc_host_plan.restype = ctypes.c_int
    L.phpc_summa_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.phpc_summa_create.restype = ctypes.c_void_p
    L.phpc_summa_create_mkn.argtypes = [ctypes.c_int] * 5
    L.phpc_summa_create_mkn.restype = ctypes.c_void_p
    L.phpc_summa_schedule_mkn.argtypes = [ctypes.c_int] * 8 + [ctypes.POINTER(SummaStep), ctypes.c_int, c_int_p, c_int_p]
    L.phpc_summa_schedule_mkn.restype = ctypes.c_int
    L.phpc_summa_global.argtypes = [ctypes.c_void_p, c_int_p]
    L.phpc_summa_schedule_first.argtypes = [ctypes.c_int] * 9 + [ctypes.POINTER(SummaStep), ctypes.c_int, c_int_p, c_int_p]
    L.phpc_summa_schedule_first.restype = ctypes.c_int
    L.phpc_summa_chunks.argtypes = [ctypes.c_void_p, c_int_p, c_int_p, c_int_p]
    L.phpc_summa_destroy.argtypes = [ctypes.c_void_p]
    L.phpc_summa_upload.argtypes = [ctypes.c_void_p, c_double_p, c_double_p, c_double_p]
    L.phpc_summa_fill.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_ulonglong]
    L.phpc_summa_zero_c.argtypes = [ctypes.c_void_p]
    L.phpc_summa_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(SummaStats)]
    L.phpc_summa_run_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int, ctypes.POINTER(SummaStats)]
    L.phpc_summa_timeline.argtypes = [ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_int]
    L.phpc_summa_timeline.restype = ctypes.c_int
    L.phpc_summa_download_c.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int]
    L.phpc_summa_read_c_block.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_longlong] + [ctypes.c_int] * 4
    L.phpc_summa_geometry.argtypes = [ctypes.c_void_p, c_int_p, c_int_p, c_int_p]
    L.get_cur_time.restype = ctypes.c_double

    M = _mpi
    M.phpc_mpi_segment_create.argtypes = [ctypes.c_char_p, ctypes.c_int]
    M.phpc_mpi_segment_unlink.argtypes = [ctypes.c_char_p]
    M.phpc_mpi_init_explicit.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    M.MPI_Comm_size.argtypes = [ctypes.c_int, c_int_p]
    M.MPI_Comm_rank.argtypes = [ctypes.c_int, c_int_p]
    M.MPI_Comm_free.argtypes = [c_int_p]
    M.MPI_Dims_create.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p]
    M.MPI_Cart_create.argtypes = [ctypes.c_int, ctypes.c_int, c_int_p, c_int_p, ctypes.c_int, c_int_p]
    M.MPI_Cart_coords.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p]
    M.MPI_Cart_sub.argtypes = [ctypes.c_int, c_int_p, c_int_p]
    M.MPI_Barrier.argtypes = [ctypes.c_int]
    M.MPI_Bcast.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    M.MPI_Reduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    M.MPI_Allreduce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    M.MPI_Send.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    M.MPI_Recv.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    M.MPI_Type_vector.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p]
    M.MPI_Type_commit.argtypes = [c_int_p]
    M.MPI_Type_free.argtypes = [c_int_p]
    _lib = L
    return L


def mpi():
    load()
    return _mpi


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ld(a):
    """Leading dimension (elements) of a 2-D row-major float64 view."""
    assert a.dtype == np.float64 and a.ndim == 2
    if a.size == 0 or a.shape[0] <= 1:
        return max(a.shape[1], 1)
    assert a.shape[1] <= 1 or a.strides[1] == 8, "rows must be contiguous"
    assert a.strides[0] % 8 == 0
    return a.strides[0] // 8


# ----------------------------------------------------------------------------
# reference-named entry points on numpy views (row-major, ld from the strides)
# ----------------------------------------------------------------------------
def _host_gemm(fn, a, b, c, gpu_count, grid_width, grid_height, block_width):
    m, k = a.shape
    k2, n = b.shape
    assert k == k2 and c.shape == (m, n)
    t = ctypes.c_float(-1.0)
    fn(_dp(a), _ld(a), _dp(b), _ld(b), _dp(c), _ld(c), m, k, n, gpu_count, grid_width, grid_height, block_width, ctypes.byref(t))
    return t.value


def phpc_gemm_cuda(a, b, c, gpu_count=1, grid_width=1, grid_height=1, block_width=32):
    """c += a @ b in place through the C-ABI phpc_gemm_cuda; returns compute_time (s)."""
    return _host_gemm(load().phpc_gemm_cuda, a, b, c, gpu_count, grid_width, grid_height, block_width)


def phpc_gemm_cublas(a, b, c, gpu_count=1):
    return _host_gemm(load().phpc_gemm_cublas, a, b, c, gpu_count, 0, 0, 0)


def phpc_gemm_summa_cuda(grid_comm, A, B, C, gpu_count=1, grid_width=1, grid_height=1, block_width=32):
    """C += A @ B (full N x N host matrices on every rank) through the C-ABI SUMMA."""
    n = A.shape[0]
    assert A.shape == B.shape == C.shape == (n, n) and all(x.flags["C_CONTIGUOUS"] for x in (A, B, C))
    t = ctypes.c_float(-1.0)
    load().phpc_gemm_summa_cuda(grid_comm, _dp(A), _dp(B), _dp(C), n, gpu_count, grid_width, grid_height, block_width, ctypes.byref(t))
    return t.value


def phpc_gemm_summa_cublas(grid_comm, A, B, C, gpu_count=1):
    n = A.shape[0]
    t = ctypes.c_float(-1.0)
    load().phpc_gemm_summa_cublas(grid_comm, _dp(A), _dp(B), _dp(C), n, gpu_count, ctypes.byref(t))
    return t.value


def host_array_shared(rows, cols):
    """rows x cols float64 matrix in page-locked host memory that every rank of the node can map (phpc_host_malloc_shared):
    as rank 0's C it makes the SUMMA gather parallel over all PCIe links.  Returns (array, pointer) or None when the
    shared-memory file system cannot hold it; free with load().phpc_host_free_shared(pointer)."""
    nbytes = rows * cols * 8
    ptr = load().phpc_host_malloc_shared(nbytes)
    if not ptr:
        return None
    buf = (ctypes.c_double * (rows * cols)).from_address(ptr)
    return np.frombuffer(buf, dtype=np.float64).reshape(rows, cols), ptr


def device_window(dev_ptr, ld, row0, col0, rows, cols):
    """rows x cols window of a device matrix (base pointer dev_ptr, leading dimension ld) as numpy."""
    out = np.empty((rows, cols), dtype=np.float64)
    load().phpc_copy2d_to_host(_dp(out), cols, dev_ptr + (row0 * ld + col0) * 8, ld, rows, cols)
    return out


def summa_schedule(N, r, c, pi, pj, kc=0):
    L = load()
    count = L.phpc_summa_schedule(N, r, c, pi, pj, kc, None, 0, None, None)
    if count < 0:
        raise ValueError("N must be divisible by the process grid dimensions")
    steps = (SummaStep * count)()
    m, n = ctypes.c_int(), ctypes.c_int()
    L.phpc_summa_schedule(N, r, c, pi, pj, kc, steps, count, ctypes.byref(m), ctypes.byref(n))
    return list(steps), m.value, n.value


def summa_schedule_mkn(M, K, N, r, c, pi, pj, kc=0):
    L = load()
    count = L.phpc_summa_schedule_mkn(M, K, N, r, c, pi, pj, kc, None, 0, None, None)
    if count < 0:
        raise ValueError("M, N must be divisible by the grid dimensions and K by their least common multiple")
    steps = (SummaStep * count)()
    m, n = ctypes.c_int(), ctypes.c_int()
    L.phpc_summa_schedule_mkn(M, K, N, r, c, pi, pj, kc, steps, count, ctypes.byref(m), ctypes.byref(n))
    return list(steps), m.value, n.value


def summa_schedule_first(M, K, N, r, c, pi, pj, kc, kc_first):
    L = load()
    count = L.phpc_summa_schedule_first(M, K, N, r, c, pi, pj, kc, kc_first, None, 0, None, None)
    if count < 0:
        raise ValueError("M, N must be divisible by the grid dimensions and K by their least common multiple")
    steps = (SummaStep * count)()
    L.phpc_summa_schedule_first(M, K, N, r, c, pi, pj, kc, kc_first, steps, count, None, None)
    return list(steps)


def ozaki_config():
    """The fixed arithmetic of the tcgen05 (Ozaki) path: digits per operand, int8 products per FP64 product, K chunk, guard spread."""
    v = [ctypes.c_int() for _ in range(4)]
    load().phpc_ozaki_config(*[ctypes.byref(x) for x in v])
    return {"digits": v[0].value, "products": v[1].value, "k_chunk": v[2].value, "max_spread": v[3].value}


def default_backend():
    """The local GEMM the reference-named entry points run in this process (tcgen05 unless PHPC_GEMM=dmma)."""
    return load().phpc_default_backend()


def host_plan(m, nsteps, bands, align=128):
    """Operation list of the band-pipelined host-sourced run (pure host arithmetic, no GPU)."""
    L = load()
    count = L.phpc_host_plan(m, nsteps, bands, align, None, 0)
    ops = (HostOp * max(count, 1))()
    L.phpc_host_plan(m, nsteps, bands, align, ops, count)
    return list(ops)[:count]


# ----------------------------------------------------------------------------
# process model for Python hosts (tests, bench.py under torchrun)
# ----------------------------------------------------------------------------
def mpi_init(rank=0, size=1, segment_path=None):
    """MPI_Init of the shim for a process that was not started by bin/mpirun."""
    M = mpi()
    path = segment_path.encode() if (segment_path and size > 1) else None
    M.phpc_mpi_init_explicit(path, rank, size)


def mpi_segment_create(path, size):
    if mpi().phpc_mpi_segment_create(path.encode(), size) != 0:
        raise OSError(f"cannot create MPI shim segment {path}")


def cart_create(dims, comm=MPI_COMM_WORLD):
    """MPI_Cart_create(comm, 2, dims, periods={1,1}, reorder=0) as reference src/main.c:58-61."""
    M = mpi()
    d = (ctypes.c_int * 2)(*dims)
    p = (ctypes.c_int * 2)(1, 1)
    out = ctypes.c_int(-1)
    M.MPI_Cart_create(comm, 2, d, p, 0, ctypes.byref(out))
    return out.value


def dims_create(size):
    """Process grid as reference src/main.c:38-45."""
    if size == 1:
        return (1, 1)
    d = (ctypes.c_int * 2)(0, 0)
    mpi().MPI_Dims_create(size, 2, d)
    return (d[0], d[1])


class Summa:
    """Device-resident SUMMA object (phpc_summa_* additions of include/phpc_summa.h)."""

    def __init__(self, grid_comm, n, kc=0, m=None, k=None):
        """n x n problem, or (m=, k=) given: C[m x n] += A[m x k] * B[k x n] (phpc_summa_create_mkn)."""
        self.L = load()
        self.n = n
        if m is None and k is None:
            self.h = self.L.phpc_summa_create(grid_comm, n, kc)
        else:
            self.h = self.L.phpc_summa_create_mkn(grid_comm, n if m is None else m, n if k is None else k, n, kc)
        g = (ctypes.c_int * 3)()
        self.L.phpc_summa_global(self.h, g)
        self.mkn = (g[0], g[1], g[2])
        kc_, kf_, ns_ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.L.phpc_summa_chunks(self.h, ctypes.byref(kc_), ctypes.byref(kf_), ctypes.byref(ns_))
        self.kc, self.kc_first, self.nsteps = kc_.value, kf_.value, ns_.value
        d, co, bl = (ctypes.c_int * 2)(), (ctypes.c_int * 2)(), (ctypes.c_int * 2)()
        self.L.phpc_summa_geometry(self.h, d, co, bl)
        self.dims, self.coords, self.block = (d[0], d[1]), (co[0], co[1]), (bl[0], bl[1])

    def upload(self, A, B, C=None):
        self.L.phpc_summa_upload(self.h, _dp(A), _dp(B), _dp(C) if C is not None else None)

    def fill(self, kind=FILL_INDEX, seed_a=SEED_A, seed_b=SEED_B):
        self.L.phpc_summa_fill(self.h, kind, seed_a, seed_b)

    def zero_c(self):
        self.L.phpc_summa_zero_c(self.h)

    def run(self, backend=None, ctas=0, stream=None, stats=True):
        backend = default_backend() if backend is None else backend
        st = SummaStats() if stats else None
        self.L.phpc_summa_run(self.h, backend, ctas, stream, ctypes.byref(st) if stats else None)
        return st

    def timeline(self):
        n = 4096
        a, b = (ctypes.c_float * n)(), (ctypes.c_float * n)()
        k = self.L.phpc_summa_timeline(self.h, a, b, n)
        return [(a[i], b[i]) for i in range(k)]

    def run_host(self, A, B, C, backend=None, ctas=0, gather=True):
        backend = default_backend() if backend is None else backend
        st = SummaStats()
        self.L.phpc_summa_run_host(self.h, backend, ctas, _dp(A), _dp(B), _dp(C), 1 if gather else 0, ctypes.byref(st))
        return st

    def download_c(self, C, gather=True):
        self.L.phpc_summa_download_c(self.h, _dp(C), 1 if gather else 0)

    def read_c_block(self, row0=0, col0=0, rows=None, cols=None):
        rows = self.block[0] - row0 if rows is None else rows
        cols = self.block[1] - col0 if cols is None else cols
        out = np.empty((rows, cols), dtype=np.float64)
        self.L.phpc_summa_read_c_block(self.h, _dp(out), cols, row0, col0, rows, cols)
        return out

    def destroy(self):
        if self.h:
            self.L.phpc_summa_destroy(self.h)
            self.h = None
