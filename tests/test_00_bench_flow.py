"""bench.py's control flow and JSON contract exercised on the CPU with the CUDA library and torch.cuda replaced by
stand-ins (the numbers are meaningless; the point is that every key the driver reads is produced, that the e2e leg
verifies its result and refuses to report a wrong one, and that no code path of bench.py raises).  The real
measurement needs a B200: `python bench.py`.

The file sorts first on purpose: torch has to be imported before libphpc_b200.so is loaded into the process (see the note
on libnccl.so.2 in capi.load), and it is imported lazily so that a `-m gpu` run, where these tests are deselected, never
pulls torch into the test process at all."""
import argparse
import ctypes
import io
import json
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

import bench


class _Stats:
    total_ms, gemm_ms, exposed_ms, steps, launches, broadcasts, bytes_received = 10.0, 9.5, 0.5, 1, 6, 0, 0


def _fake_value(seed, rows, cols, N):
    """Position-based stand-in for the library's counter-based fill: any window regenerates the same values."""
    flat = (np.asarray(rows, dtype=np.uint64)[:, None] * np.uint64(N) + np.asarray(cols, dtype=np.uint64)[None, :])
    h = (flat * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503)) % np.uint64(1 << 32)
    return h.astype(np.float64) / float(1 << 31) - 1.0


class _FakeSumma:
    wrong = False

    def __init__(self, comm, n, kc=0):
        self.n = n
        self.block = (n, n)
        self.coords = (0, 0)
        self.dims = (1, 1)

    def fill(self, kind):
        idx = np.arange(self.n)
        self.A = _fake_value(1234, idx, idx, self.n)
        self.B = _fake_value(5678, idx, idx, self.n)
        self.C = np.zeros((self.n, self.n))

    def run(self, backend, ctas, stream, stats=True):
        self.C += self.A @ self.B
        if self.wrong:
            self.C[0, 0] += 1.0
        return _Stats() if stats else None

    def read_c_block(self, row0=0, col0=0, rows=None, cols=None):
        return self.C[row0:row0 + rows, col0:col0 + cols].copy()

    def destroy(self):
        pass


class _FakeLib:
    def __getattr__(self, name):  # every C entry point: accept anything, return 1 (device count etc.)
        return lambda *a, **k: 1


def _fake_capi(wrong_result=False):
    m = types.ModuleType("capi")
    m.BACKEND_DMMA, m.BACKEND_CUBLAS, m.BACKEND_OZAKI = 0, 1, 2
    m.FILL_INDEX, m.FILL_SEEDED, m.SEED_A, m.SEED_B = 0, 1, 1234, 5678
    m.c_double_p = ctypes.POINTER(ctypes.c_double)
    lib = _FakeLib()

    def fill_host(ptr, ld, rows, cols, row0, col0, N, kind, seed):
        a = np.ctypeslib.as_array(ptr, shape=((rows - 1) * ld + cols,))
        v = _fake_value(seed, np.arange(row0, row0 + rows), np.arange(col0, col0 + cols), N)
        for r in range(rows):
            a[r * ld:r * ld + cols] = v[r]
        return 0

    lib.phpc_fill_host = fill_host
    m.load = lambda: lib

    def ozaki_config():
        return {"digits": 7, "products": 28, "k_chunk": 8192, "max_spread": 40}

    m.ozaki_config = ozaki_config
    m.mpi_init = lambda *a: None
    m.cart_create = lambda dims: 0
    m.Summa = _FakeSumma

    def summa_cuda(comm, A, B, C):
        C += A @ B
        if wrong_result:
            C[0, 0] += 1.0
        return 0.001

    m.phpc_gemm_summa_cuda = summa_cuda
    return m


class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 30.0


@pytest.fixture
def fake_cuda(monkeypatch):
    try:
        import torch
    except ImportError as e:  # libphpc_b200.so came first in this process and bound the system libnccl.so.2
        pytest.skip(f"torch cannot be imported after libphpc_b200.so in the same process: {e}")

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("LOCAL_RANK", raising=False)


def _args(**kw):
    d = dict(gpus=1, steps=2, warmup=1, impl="b200", n=64, kc=0, e2e_steps=1, no_e2e=False, no_cpu=True, no_refcuda=True, no_secondary=False,
             e2e_child=False)
    d.update(kw)
    return argparse.Namespace(**d)


def test_product_arm_prints_the_contract_line(monkeypatch, fake_cuda):
    import hpc_multigpu_matrixmult_b200 as pkg

    monkeypatch.setattr(pkg, "capi", _fake_capi(), raising=False)
    monkeypatch.setitem(sys.modules, "hpc_multigpu_matrixmult_b200.capi", pkg.capi)
    # the e2e leg normally runs in a child process; here it runs in-process against the stand-in library
    monkeypatch.setattr(bench, "e2e_in_child",
                        lambda args: bench.e2e_measure(args, pkg.capi, pkg.capi.load(), 0, (1, 1), 0, 1, lambda: None, lambda x: x))
    out = io.StringIO()
    with redirect_stdout(out):
        bench.product_arm(_args())
    lines = [l for l in out.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1  # ONE JSON line
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["metric"] == "summa_gemm_tflops" and line["unit"] == "TFLOP/s" and line["dtype"] == "f64" and line["n_gpus"] == 1
    assert "workload" in line["config"] and "model" not in line["config"]
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roof, key
    assert roof["bound"] == "tensor" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    e2e = line["e2e"]
    assert e2e["verified"] is True and e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 3 * 8 * 64 * 64 and e2e["d2h_bytes_per_step"] == 8 * 64 * 64
    assert line["gpu_launches"] == _Stats.launches * 2
    assert line["value_verified"] is True and line["value_verify"]["elements_checked_per_rank"] >= 16
    assert line["value_verify"]["passes_accumulated"] == 3  # 1 warm-up + 2 timed steps
    assert e2e["verify"]["worst_error_over_bound"] < 1.0
    assert line["native_fp64_dmma"]["roofline"]["unit"] == "TFLOP/s"  # the second kernel is reported beside the first


def test_e2e_leg_refuses_to_report_a_wrong_result(monkeypatch, fake_cuda):
    capi = _fake_capi(wrong_result=True)
    e2e = bench.e2e_measure(_args(), capi, capi.load(), 0, (1, 1), 0, 1, lambda: None, lambda x: x)
    assert e2e["verified"] is False and e2e["value"] is None


def test_value_verification_catches_a_wrong_device_result(monkeypatch, fake_cuda):
    import hpc_multigpu_matrixmult_b200 as pkg

    capi = _fake_capi()

    class Wrong(_FakeSumma):
        wrong = True

    capi.Summa = Wrong
    monkeypatch.setattr(pkg, "capi", capi, raising=False)
    monkeypatch.setitem(sys.modules, "hpc_multigpu_matrixmult_b200.capi", pkg.capi)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.product_arm(_args(no_e2e=True, no_secondary=True))
    line = json.loads([l for l in out.getvalue().splitlines() if l.startswith("{")][0])
    assert line["value_verified"] is False


def test_roofline_counts_the_int8_products_of_the_fixed_arithmetic(monkeypatch, fake_cuda):
    import hpc_multigpu_matrixmult_b200 as pkg

    monkeypatch.setattr(pkg, "capi", _fake_capi(), raising=False)
    monkeypatch.setitem(sys.modules, "hpc_multigpu_matrixmult_b200.capi", pkg.capi)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.product_arm(_args(no_e2e=True, no_secondary=True))
    line = json.loads([l for l in out.getvalue().splitlines() if l.startswith("{")][0])
    assert "28 int8 MMAs" in line["roofline"]["note"] and "tcgen05" in line["roofline"]["kernel"]
    assert line["roofline"]["ops_per_launch"] == 2.0 * 64 ** 3 * 28
    assert "emulated FP64" in line["config"]["arithmetic"]


def test_reference_arm_line(monkeypatch):
    monkeypatch.setattr(bench, "run_reference_cpu", lambda n, opt="O0": (1e-3, 2.0, "reference"))
    monkeypatch.setattr(bench, "run_reference_summa_all_cores", lambda n=4096: {"value": 1e-2, "unit": "TFLOP/s", "cores": 8, "kind": "port"})
    monkeypatch.delenv("RANK", raising=False)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.reference_arm(_args(impl="reference"))
    line = json.loads(out.getvalue().strip())
    assert line["impl"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "reference"
    assert line["metric"] == "summa_gemm_tflops" and line["value"] == line["e2e"]["value"] == line["cpu_baseline"]["value"]
    # ranks other than 0 print nothing
    monkeypatch.setenv("RANK", "1")
    out = io.StringIO()
    with redirect_stdout(out):
        bench.reference_arm(_args(impl="reference"))
    assert out.getvalue() == ""


def test_e2e_child_prints_one_tagged_json_object(monkeypatch, fake_cuda):
    import hpc_multigpu_matrixmult_b200 as pkg

    monkeypatch.setattr(pkg, "capi", _fake_capi(), raising=False)
    monkeypatch.setitem(sys.modules, "hpc_multigpu_matrixmult_b200.capi", pkg.capi)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.e2e_child(_args())
    tagged = [l for l in out.getvalue().splitlines() if l.startswith("E2E_JSON ")]
    assert len(tagged) == 1
    e2e = json.loads(tagged[0][len("E2E_JSON "):])
    assert e2e["verified"] is True and e2e["unit"] == "TFLOP/s" and "host_row_bands" in e2e


def test_all_cores_reference_summa_baseline_runs_here(built):
    """The reference's phpc_summa.c (unchanged) over the MPI shim with the CPU plugin, one rank per core, timed."""
    out = bench.run_reference_summa_all_cores(n=512)
    assert out is not None and out["value"] > 0 and out["cores"] >= 1 and out["kind"] == "port"


def test_a_throttled_measurement_is_taken_again(monkeypatch, fake_cuda):
    import hpc_multigpu_matrixmult_b200 as pkg

    monkeypatch.setattr(pkg, "capi", _fake_capi(), raising=False)
    monkeypatch.setitem(sys.modules, "hpc_multigpu_matrixmult_b200.capi", pkg.capi)
    calls = []

    def stop(self, t0, t1):
        calls.append(1)
        reasons = ["hw_thermal_slowdown"] if len(calls) == 1 else ["sw_power_cap"]
        return {"sm_mhz": 1400.0, "sm_max_mhz": 1965.0, "power_w_max": 990.0, "samples": 5, "reasons": reasons}

    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", stop)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.product_arm(_args(no_e2e=True, no_secondary=True))
    line = json.loads([l for l in out.getvalue().splitlines() if l.startswith("{")][0])
    assert line["clocks"]["reasons"] == ["sw_power_cap"]  # the kept measurement; power capping is normal and only noted
    assert line["clocks"]["remeasured_after"]["reasons"] == ["hw_thermal_slowdown"]
