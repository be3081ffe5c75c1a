#!/usr/bin/env python
"""Summarises ncu outputs into the small text files kept under profiles/.
  launches:  python tools/ncu_summary.py launches <launches.csv>
  kernel:    python tools/ncu_summary.py kernel <report.ncu-rep>
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        n = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>6s}")
    for n, a in agg.items():
        print(f"{n:44s} {a[0]:8d} {a[1]:10.3f} {a[1] / a[0]:9.3f} {a[1] / tot:6.3f}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for d in rows[2:]:
        print("kernel:", d[h.index("Kernel Name")])
        for w in WANT:
            if w in h:
                print(f"  {w:86s} {d[h.index(w)]:>18s} {u[h.index(w)]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
