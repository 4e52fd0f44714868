#!/usr/bin/env python
"""Prints the metrics we track from an `ncu --set full` report (run where there is no GPU).
usage: summarize.py <report.ncu-rep>"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:100])
    for k in KEYS:
        if k in hdr:
            print("  %-72s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    print("  stall reasons (warps stalled per issue-active cycle):")
    st = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:8]:
        print("    %-40s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


def write_profile_json(report, workload, walks_per_launch, capture_name, path="profiles/walk_kernel_profile.json"):
    """Adds the per-build figures bench.py quotes (roofline.traffic, roofline_issue) for the walk kernel in `report`."""
    import json
    import os
    import re
    m = re.search(r"fastKernel<(\d)", rows[2][hdr.index("Kernel Name")])
    r = rows[2]
    val = lambda k: float(r[hdr.index(k)].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum(val(k)*scale[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    prof = json.load(open(path)) if os.path.exists(path) else {}
    prof["fastKernel<%s>" % m.group(1)] = {
        "capture": capture_name, "workload": workload, "walks_per_launch": walks_per_launch,
        "dram_bytes_per_launch": dram, "warp_inst_per_walk": val("smsp__inst_executed.sum")/walks_per_launch,
        "active_lanes_per_inst": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "kernel_ms_under_ncu": val("gpu__time_duration.sum")*{"ms": 1.0, "us": 1e-3, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]}
    json.dump(prof, open(path, "w"), indent=1, sort_keys=True)


if len(sys.argv) >= 5:  # summarize.py report.ncu-rep WORKLOAD WALKS_PER_LAUNCH CAPTURE_NAME  -> updates walk_kernel_profile.json
    write_profile_json(sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4])
