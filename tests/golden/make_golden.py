"""
make_golden.py — regenerates tests/golden/reference_outputs.npz from the reference
itself (run in the build container, where /root/reference exists):

  iterative_N{n}_F{f}   C of the reference's own phpc_gemm_iterative (src/iterative.c,
                        compiled unchanged into oracle/_ref/libref_iterative.so)
  summa_N{n}_P{p}_F{f}  rank 0's gathered C of the reference's own phpc_gemm_summa_cuda
                        (src/phpc_summa.c compiled unchanged, run under bin/mpirun with
                        the CPU gemm_t plugin oracle/ref_cpu_plugin.c), plus its grid

Inputs are regenerable: F0 = the reference fill A[i]=B[i]=i (src/main.c:85-86),
F1 = splitmix64 seeded uniform(-1,1) with seeds 1234 (A) / 5678 (B).
The reference ships no golden vectors of its own; these
files are what pins the oracle when /root/reference is absent (GPU box).

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402


def main():
    o.build(with_ref=True)
    assert o.have_ref(), "oracle/_ref missing: /root/reference must be present to regenerate"
    out = {}
    for n in (32, 64):
        for f in (0, 1):
            A = o.fill(n, n, kind=f, seed=o.SEED_A)
            B = o.fill(n, n, kind=f, seed=o.SEED_B)
            out[f"iterative_N{n}_F{f}"] = o.ref_iterative(A, B)
    with tempfile.TemporaryDirectory() as d:
        for n, ranks in ((48, (1, 2, 4, 6, 8, 16)),):
            for p in ranks:
                for f in (0, 1):
                    C, grid = o.ref_summa_cpu(n, p, f, d)
                    out[f"summa_N{n}_P{p}_F{f}"] = C
                    out[f"summa_N{n}_P{p}_grid"] = np.array(grid, dtype=np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
