#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small CSVs committed under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/rXX_launch_summary.csv
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rXX_ncu_summary.csv

`launches`: per-kernel count, mean device time and share of the profiled steps (cold-cache,
serialised -- compare shares, not absolutes).  `full`: one row per captured launch with the
metrics DESIGN.md quotes (duration, DRAM bytes, pipe utilisation, issue activity, stall reasons).
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "not_selected",
          "branch_resolving", "no_instruction", "mio_throttle", "lg_throttle", "dispatch_stall", "membar",
          "sleeping"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    d = collections.OrderedDict()
    for r in rows:
        k = r[ki].split("(")[0].replace("void ", "")
        d.setdefault(k, {"t": [], "grid": r[gi], "block": r[bi]})["t"].append(float(r[vi]))
    tot = sum(sum(v["t"]) for v in d.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "grid", "block", "mean_us", "share_pct"])
        for k, v in d.items():
            w.writerow([k, len(v["t"]), v["grid"], v["block"], "%.3f" % (sum(v["t"]) / len(v["t"]) / 1e3),
                        "%.2f" % (100 * sum(v["t"]) / tot)])
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = ["Kernel Name", "Block Size", "Grid Size"] + [m for m in FULL_METRICS if m in hdr]
    cols += ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s for s in STALLS
             if "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for r in data])
        for c in cols[1:]:
            i = hdr.index(c)
            w.writerow([c, units[i]] + [r[i] for r in data])
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
