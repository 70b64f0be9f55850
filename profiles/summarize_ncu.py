#!/usr/bin/env python
"""Turns the ncu outputs a gpurun call brings back (gpurun_out/) into the small text summaries kept under profiles/.
  python profiles/summarize_ncu.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.txt
  python profiles/summarize_ncu.py full gpurun_out/prof_r1.ncu-rep  > profiles/r1_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "dram__sectors_read.sum", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            agg[d["Kernel Name"]].append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)  source: {path}")
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:70]:70s} {len(v):8d} {sum(v) / 1e6:10.3f} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / tot:7.1%}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on   source: {path}")
    for r in rows[2:]:
        print("---", r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:82s} {r[i]:>22s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
