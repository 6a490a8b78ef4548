"""sass_diff.py — are the kernels of the default path byte-for-byte the same program as in an earlier commit?

    python tools/sass_diff.py <commit>        # e.g. the last commit whose kernels ran on a B200

Builds libphpc_b200.so of <commit> in a temporary git worktree, dumps the SASS of both libraries (cuobjdump, instruction
text without addresses and encodings) and compares every kernel of the old library with the kernel of the same (or the
correspondingly re-templated) name in the working tree.  Used at the end of round 1, when host-side work and opt-in
experimental kernels were added without GPU time left: it shows the validated device code did not change."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RENAMES = [("ozaki_gemm_kernelILi8EE", "ozaki_gemm_kernelILi8ELb0EE"), ("ozaki_gemm_kernelILi0EE", "ozaki_gemm_kernelILi0ELb0EE")]


def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        line = re.sub(r"/\*[0-9a-f]{4,}\*/|/\* 0x[0-9a-f]+ \*/", "", line).strip()
        if cur and line and not line.startswith("."):
            funcs[cur].append(line)
    return funcs


def main():
    commit = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        tree = os.path.join(tmp, "old")
        subprocess.run(["git", "-C", ROOT, "worktree", "add", "-f", tree, commit], check=True, capture_output=True)
        try:
            subprocess.run(["make", "-C", os.path.join(tree, "hpc_multigpu_matrixmult_b200"), "lib/libphpc_b200.so"], check=True, capture_output=True)
            old = sass(os.path.join(tree, "hpc_multigpu_matrixmult_b200", "lib", "libphpc_b200.so"))
        finally:
            subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", tree], capture_output=True)
    new = sass(os.path.join(ROOT, "hpc_multigpu_matrixmult_b200", "lib", "libphpc_b200.so"))
    bad = 0
    for name, body in old.items():
        other = name
        for a, b in RENAMES:
            other = other.replace(a, b)
        status = "MISSING" if other not in new else ("same" if new[other] == body else "DIFFERENT")
        bad += status != "same"
        print(f"{status:9s} {len(body):5d} instructions  {name}")
    print(f"{len(new) - len(old)} kernels only in the working tree:", ", ".join(sorted(set(new) - {n.replace(a, b) for n in old for a, b in RENAMES} - set(old)))[:600])
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
