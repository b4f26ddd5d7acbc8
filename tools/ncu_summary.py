#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) or a launch-list CSV into the text files
committed under profiles/.

    python tools/ncu_summary.py report gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt
    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import csv
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_evict_last_lookup_hit.sum", "lts__t_sectors_evict_last_lookup_miss.sum",
    "lts__t_sectors_evict_first_lookup_hit.sum", "lts__t_sectors_evict_first_lookup_miss.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[h.index("Kernel Name")])
        for n in WANT:
            if n in h:
                i = h.index(n)
                print("  %-82s %-16s %s" % (n, u[i], r[i]))


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    agg, order = {}, []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][-70:]
        v = float(r[vi].replace(",", ""))
        if name not in agg:
            agg[name] = []
            order.append(name)
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%-72s %6s %12s %12s %7s" % ("kernel", "count", "avg_us", "total_ms", "share"))
    for n in order:
        v = agg[n]
        print("%-72s %6d %12.1f %12.3f %6.1f%%" % (n, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e6, 100 * sum(v) / tot))
    # the launches of this library one by one, in order (a bench step = one full-batch lookup_kernel
    # launch; the short lookup_kernel launches are the 4 Mi-query chunks of the host-buffer pipeline)
    print()
    print("launches of sshash_b200 kernels >= 1 ms, in order (ms):")
    line = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi or "sshash_b200" not in r[ki]:
            continue
        v = float(r[vi].replace(",", "")) / 1e6
        if v >= 1.0:
            line.append("%s [grid %s] %.3f" % (re.search(r"(\w+_kernel)", r[ki]).group(1), r[gi].strip("()").split(",")[0], v))
    print("  " + "\n  ".join(line))


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2])
