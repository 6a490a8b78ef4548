"""tools/ozaki_variants.py is the one GPU call that validates the experimental Ozaki kernels; a Python slip in it would
waste that call.  Its worker is run here against a stand-in for the device side of the C-ABI (device memory = host
buffers, both GEMM entry points = numpy), so every line of the tool executes on the CPU: fills, uploads, the per-row
comparison, the bit-exact rule for the reference's fill, the JSON it prints and the hang report."""
import ctypes
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_tool():
    spec = importlib.util.spec_from_file_location("ozaki_variants", os.path.join(ROOT, "tools", "ozaki_variants.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _FakeDevice:
    """Device memory as host buffers; pointers are plain addresses, as the real C-ABI hands them to Python."""

    def __init__(self, real_lib, ozaki_error=0.0, idle=True):
        self.real, self.bufs, self.ozaki_error, self.idle = real_lib, {}, ozaki_error, idle

    def _view(self, ptr, rows, ld):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(rows, ld))

    def phpc_b200_set_device(self, d):
        return 0

    def phpc_fill_host(self, *a):
        return self.real.phpc_fill_host(*a)  # host code of the real library

    def phpc_device_malloc(self, nbytes):
        buf = np.zeros(max(nbytes // 8, 1))
        self.bufs[buf.ctypes.data] = buf
        return buf.ctypes.data

    def phpc_device_free(self, ptr):
        self.bufs.pop(ptr, None)

    def phpc_device_memset(self, ptr, value, nbytes):
        ctypes.memset(ptr, value, nbytes)

    def phpc_copy2d_to_device(self, dst, ld_dst, src, ld_src, rows, cols):
        s = np.ctypeslib.as_array(src, shape=(rows, ld_src))
        self._view(dst, rows, ld_dst)[:, :cols] = s[:, :cols]

    def phpc_fill_device(self, ptr, ld, rows, cols, row0, col0, N, kind, seed, stream):
        self.real.phpc_fill_host(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), ld, rows, cols, row0, col0, N, kind, seed)

    def _gemm(self, dA, lda, dB, ldb, dC, ldc, m, k, n, err):
        c = self._view(dC, m, ldc)
        c[:, :n] += self._view(dA, m, lda)[:, :k] @ self._view(dB, k, ldb)[:, :n]
        if err:
            c[0, 0] *= 1.0 + err

    def phpc_gemm_device(self, dA, lda, dB, ldb, dC, ldc, m, k, n, ctas, stream):
        self._gemm(dA, lda, dB, ldb, dC, ldc, m, k, n, 0.0)
        return 1

    def phpc_gemm_device_ozaki(self, dA, lda, dB, ldb, dC, ldc, m, k, n, slices, stream):
        self._gemm(dA, lda, dB, ldb, dC, ldc, m, k, n, self.ozaki_error)
        return 6

    def phpc_gemm_device_timed(self, dA, lda, dB, ldb, dC, ldc, m, k, n, ctas, reps, backend):
        return 1.5

    def phpc_device_synchronize(self):
        return 0

    def phpc_compute_stream_idle(self):
        return 1 if self.idle else 0

    class _Progress:
        argtypes = None

        def __call__(self, words, n):
            for c in range(4):  # two CTA pairs: producer waiting on its 7th load, MMA issuer waiting for peer_full, ...
                words[c * 8 + 0] = (1 << 28) | 7
                words[c * 8 + 1] = (5 << 28) | 3
                words[c * 8 + 6] = (2 << 28)
            return 32

    phpc_oz_progress_read = _Progress()


@pytest.fixture
def tool(capi, monkeypatch):
    real = capi.load()
    mod = _load_tool()

    def install(**kw):
        dev = _FakeDevice(real, **kw)
        fake = types.ModuleType("capi")
        fake.c_double_p = capi.c_double_p
        fake.load = lambda: dev

        def device_window(ptr, ld, row0, col0, rows, cols):
            return dev._view(ptr, row0 + rows, ld)[row0:row0 + rows, col0:col0 + cols].copy()

        fake.device_window = device_window
        import hpc_multigpu_matrixmult_b200 as pkg

        monkeypatch.setattr(pkg, "capi", fake, raising=False)
        monkeypatch.setitem(sys.modules, "hpc_multigpu_matrixmult_b200.capi", fake)
        return mod

    return install


def test_worker_passes_every_check_when_the_kernels_agree(tool, monkeypatch):
    mod = tool()
    monkeypatch.setattr(mod, "SHAPES", [(128, 128, 128), (100, 77, 50), (300, 520, 200)])
    out = io.StringIO()
    with redirect_stdout(out):
        rc = mod.worker([256])
    lines = [json.loads(l) for l in out.getvalue().splitlines()]
    checks = [l for l in lines if "check" in l]
    assert rc == 0 and len(checks) == 9 and all(c["ok"] for c in checks)
    assert any(c["fill"] == "index" and c["bit_equal"] for c in checks) and any(c["tweak"] == "scaled" for c in checks)
    timing = [l for l in lines if "time_n" in l]
    assert timing and abs(timing[0]["fp64_equivalent_tflops"] - 2.0 * 256 ** 3 / 1.5 / 1e9) < 1e-9


def test_worker_fails_the_variant_and_skips_timing_when_a_kernel_is_wrong(tool, monkeypatch):
    mod = tool(ozaki_error=1e-9)
    monkeypatch.setattr(mod, "SHAPES", [(128, 128, 128)])
    out = io.StringIO()
    with redirect_stdout(out):
        rc = mod.worker([256])
    lines = [json.loads(l) for l in out.getvalue().splitlines()]
    assert rc == 1 and any(not l["ok"] for l in lines if "check" in l) and not any("time_n" in l for l in lines)


def test_hang_report_decodes_the_progress_words(tool, monkeypatch):
    mod = tool(idle=False)
    monkeypatch.setattr(mod.os, "_exit", lambda code: (_ for _ in ()).throw(SystemExit(code)))
    out = io.StringIO()
    with redirect_stdout(out), pytest.raises(SystemExit) as e:
        mod.wait_or_report_hang(mod.__dict__["worker"].__globals__["sys"].modules["hpc_multigpu_matrixmult_b200.capi"].load(), (256, 128, 128), seconds=0.05)
    assert e.value.code == 3
    rep = json.loads(out.getvalue().strip())
    assert rep["hang"] == [256, 128, 128] and rep["ctas_total"] == 4 and rep["ctas_not_finished"] == 4
    assert rep["first_ctas"][0]["producer"] == "waiting@7" and rep["first_ctas"][0]["mma/relay"].startswith("own full ok")
