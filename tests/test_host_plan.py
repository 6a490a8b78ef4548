"""The band pipeline of the single-GPU host-sourced run (phpc_host_plan, include/phpc_summa.h) checked
on the CPU (bands of 1/2, 1/4, ... of the block; each band runs on a zeroed block and the caller's C rows, uploaded
meanwhile into a side buffer, are added at its end): the operation list the CUDA executor walks (csrc/phpc_summa.cu, summa_run_host_banded) is
pure host arithmetic, so its ordering is proven here without a GPU:
  * every pair of operations that touch the same memory (one of them writing) is ordered by the
    stream order + the listed dependencies (no read-after-write / write-after-read race);
  * every execution order those constraints allow produces exactly C + A*B on integer inputs, with
    each element's K chunks added in ascending order (the reference's order, src/phpc_gemm.cu:33-52).
The GPU side of the same path is tests/test_gpu_parity.py::test_host_bands_*."""
import random

import numpy as np
import pytest


def _regions(op, capi):
    """(reads, writes): sets of (buffer, band-or-step) names an operation touches."""
    k = op.kind
    if k == capi.HOP_UPLOAD_C:
        return {("hC", op.band)}, {("dC0", op.band)}
    if k == capi.HOP_ZERO_C:
        return set(), {("dC", op.band)}
    if k == capi.HOP_ADD_C:
        return {("dC0", op.band), ("dC", op.band)}, {("dC", op.band)}
    if k == capi.HOP_UPLOAD_A:
        return {("hA", op.band, op.step)}, {("dA", op.band, op.step)}
    if k == capi.HOP_UPLOAD_B:
        return {("hB", op.step)}, {("dB", op.step)}
    if k == capi.HOP_GEMM:
        return {("dA", op.band, op.step), ("dB", op.step), ("dC", op.band)}, {("dC", op.band)}
    if k == capi.HOP_DOWNLOAD_C:
        return {("dC", op.band)}, {("hC", op.band)}
    raise AssertionError(k)


def _edges(ops):
    """Ordering constraints the executor enforces: in-order streams + explicit dependencies."""
    last = {}
    edges = [[] for _ in ops]
    for i, op in enumerate(ops):
        if op.stream in last:
            edges[i].append(last[op.stream])
        last[op.stream] = i
        for d in range(op.ndeps):
            assert 0 <= op.deps[d] < i, "dependencies must point backwards"
            assert ops[op.deps[d]].stream != op.stream
            edges[i].append(op.deps[d])
    return edges


def _reach(edges):
    n = len(edges)
    before = [0] * n  # bitset of operations ordered before i
    for i in range(n):
        b = 0
        for p in edges[i]:
            b |= before[p] | (1 << p)
        before[i] = b
    return before


@pytest.mark.parametrize("m,nsteps,bands,align", [(1000, 3, 4, 128), (32768, 8, 8, 128), (40, 5, 3, 1), (7, 1, 8, 1), (300, 2, 1, 128)])
def test_conflicting_operations_are_ordered(capi, m, nsteps, bands, align):
    ops = capi.host_plan(m, nsteps, bands, align)
    edges = _edges(ops)
    before = _reach(edges)
    touched = [_regions(op, capi) for op in ops]
    for i in range(len(ops)):
        ri, wi = touched[i]
        for j in range(i):
            rj, wj = touched[j]
            if (wi & (rj | wj)) or (wj & ri):
                assert before[i] >> j & 1, f"operations {j} and {i} race on {(wi & (rj | wj)) | (wj & ri)}"
    # the bands tile the block exactly, in multiples of `align`
    rows = sorted({(op.row0, op.rows) for op in ops if op.kind == capi.HOP_UPLOAD_C})
    assert rows[0][0] == 0 and sum(r for _, r in rows) == m
    for (r0, n0), (r1, _) in zip(rows, rows[1:]):
        assert r0 + n0 == r1 and n0 % align == 0
    sizes = [n0 for _, n0 in rows]
    assert len(sizes) <= bands and sizes == sorted(sizes, reverse=True)  # 1/2, 1/4, ...: the first band is the tallest
    if len(sizes) >= 3 and m >= 8 * align * bands:
        assert sizes[0] >= m // 2 and sizes[1] <= sizes[0] // 2 + align
    kinds = [op.kind for op in ops]
    nb = len(rows)
    assert kinds.count(capi.HOP_GEMM) == nb * nsteps and kinds.count(capi.HOP_UPLOAD_B) == nsteps
    assert kinds.count(capi.HOP_UPLOAD_A) == nb * nsteps and kinds.count(capi.HOP_DOWNLOAD_C) == nb
    assert kinds.count(capi.HOP_ZERO_C) == nb and kinds.count(capi.HOP_ADD_C) == nb
    # the caller's C rows are not needed before the first GEMM: within a band they are uploaded after every A window
    for b in range(nb):
        idx = {k: [i for i, op in enumerate(ops) if op.kind == k and op.band == b] for k in set(kinds)}
        assert idx[capi.HOP_UPLOAD_C][0] > max(idx[capi.HOP_UPLOAD_A])
        assert idx[capi.HOP_ZERO_C][0] < min(idx[capi.HOP_GEMM]) and idx[capi.HOP_ADD_C][0] > max(idx[capi.HOP_GEMM])


@pytest.mark.parametrize("m,n,widths,bands,align", [(23, 9, (4, 4, 3), 3, 1), (64, 10, (8, 8), 4, 8), (5, 5, (5,), 2, 1)])
def test_every_legal_order_gives_c_plus_ab(capi, m, n, widths, bands, align):
    rng = np.random.default_rng(7)
    K = sum(widths)
    A = rng.integers(-9, 10, (m, K)).astype(np.float64)
    B = rng.integers(-9, 10, (K, n)).astype(np.float64)
    C0 = rng.integers(-9, 10, (m, n)).astype(np.float64)
    k0 = np.concatenate([[0], np.cumsum(widths)])
    ops = capi.host_plan(m, len(widths), bands, align)
    edges = _edges(ops)
    for trial in range(6):
        hC = C0.copy()
        dA = np.full((m, K), np.nan)
        dB = np.full((K, n), np.nan)
        dC = np.full((m, n), np.nan)
        dC0 = np.full((m, n), np.nan)
        chunk_order = {}
        done, pending = set(), list(range(len(ops)))
        r = random.Random(trial)
        while pending:
            ready = [i for i in pending if all(p in done for p in edges[i])]
            i = ready[0] if trial == 0 else r.choice(ready)  # trial 0 = issue order
            op = ops[i]
            rows = slice(op.row0, op.row0 + op.rows)
            if op.kind == capi.HOP_UPLOAD_C:
                dC0[rows] = hC[rows]
            elif op.kind == capi.HOP_ZERO_C:
                dC[rows] = 0.0
            elif op.kind == capi.HOP_ADD_C:
                dC[rows] += dC0[rows]
            elif op.kind == capi.HOP_UPLOAD_A:
                dA[rows, k0[op.step]:k0[op.step + 1]] = A[rows, k0[op.step]:k0[op.step + 1]]
            elif op.kind == capi.HOP_UPLOAD_B:
                dB[k0[op.step]:k0[op.step + 1]] = B[k0[op.step]:k0[op.step + 1]]
            elif op.kind == capi.HOP_GEMM:
                ks = slice(k0[op.step], k0[op.step + 1])
                dC[rows] += dA[rows, ks] @ dB[ks]
                chunk_order.setdefault(op.band, []).append(op.step)
            elif op.kind == capi.HOP_DOWNLOAD_C:
                hC[rows] = dC[rows]
            done.add(i)
            pending.remove(i)
        assert np.array_equal(hC, C0 + A @ B)  # NaN anywhere = something was used before it arrived
        assert all(steps == sorted(steps) for steps in chunk_order.values())
