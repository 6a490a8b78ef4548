"""ncu_summary.py <report.ncu-rep> [regex ...] — dump the metrics quoted in DESIGN.md / bench.py
from an `ncu --set full` report as 'metric unit value' lines (one block per profiled launch)."""
import csv
import re
import subprocess
import sys

DEFAULT = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^gpu__dram_throughput\.avg\.pct", r"^lts__throughput\.avg\.pct",
           r"^sm__throughput\.avg\.pct", r"^sm__pipe_tensor_cycles_active\.avg\.pct", r"^sm__pipe_tensor_subpipe_(dmma|imma|hmma)_cycles_active\.avg\.pct",
           r"^sm__inst_executed_pipe_tensor", r"^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active", r"^sm__warps_active\.avg\.pct",
           r"^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit)", r"^sm__issue_active\.avg\.pct",
           r"^smsp__average_warps_issue_stalled_.*_per_issue_active", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
           r"^sm__cycles_active\.avg$", r"^smsp__inst_executed\.sum$", r"^sm__inst_executed_pipe_uniform", r"tmem", r"^sm__pipe_tc_"]


def main():
    rep = sys.argv[1]
    pats = [re.compile(p) for p in (sys.argv[2:] or DEFAULT)]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        for h, u, v in zip(hdr, units, vals):
            if any(p.search(h) for p in pats) and v not in ("", "0"):
                print(f"{h} [{u}] {v}")


if __name__ == "__main__":
    main()
