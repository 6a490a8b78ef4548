"""ncu_summary.py <report.ncu-rep> [--traffic-json out.json --shape m,k,n] [regex ...] — dump the metrics quoted in DESIGN.md / bench.py
from an `ncu --set full` report as 'metric unit value' lines (one block per profiled launch).
--traffic-json also writes the DRAM traffic of the LAST profiled launch as JSON ({"kernel", "dram_bytes", "duration_ms", ...});
bench.py fills roofline.traffic from profiles/ozaki_traffic*.json when the launch shape matches."""
import csv
import json
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

DEFAULT = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^gpu__dram_throughput\.avg\.pct", r"^lts__throughput\.avg\.pct",
           r"^sm__throughput\.avg\.pct", r"^sm__pipe_tensor_cycles_active\.avg\.pct", r"^sm__pipe_tensor_subpipe_(dmma|imma|hmma)_cycles_active\.avg\.pct",
           r"^sm__inst_executed_pipe_tensor", r"^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active", r"^sm__warps_active\.avg\.pct",
           r"^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit)", r"^sm__issue_active\.avg\.pct",
           r"^smsp__average_warps_issue_stalled_.*_per_issue_active", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
           r"^sm__cycles_active\.avg$", r"^smsp__inst_executed\.sum$", r"^sm__inst_executed_pipe_uniform", r"tmem", r"^sm__pipe_tc_"]


def main():
    rep = sys.argv[1]
    rest = sys.argv[2:]
    traffic_out = None
    if "--traffic-json" in rest:
        i = rest.index("--traffic-json")
        traffic_out = rest[i + 1]
        rest = rest[:i] + rest[i + 2:]
    shape = None
    if "--shape" in rest:  # m,k,n of the profiled launch (ncu cannot know it); stored in the traffic JSON
        i = rest.index("--shape")
        shape = [int(x) for x in rest[i + 1].split(",")]
        rest = rest[:i] + rest[i + 2:]
    pats = [re.compile(p) for p in (rest or DEFAULT)]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        for h, u, v in zip(hdr, units, vals):
            if any(p.search(h) for p in pats) and v not in ("", "0"):
                print(f"{h} [{u}] {v}")
        if traffic_out:
            def val(metric):
                i = hdr.index(metric)
                return float(vals[i].replace(",", "")) * UNIT[units[i]]

            traffic = {"kernel": name, "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                       "duration_ms": val("gpu__time_duration.sum"), "grid": vals[hdr.index("launch__grid_size")], "report": rep, "shape_mkn": shape}
            traffic["dram_bytes"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
    if traffic_out:
        with open(traffic_out, "w") as f:
            json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main()
