#!/usr/bin/env python
"""Summarise ncu outputs into text for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/launches.csv
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        name = r[ki].split('(')[0]
        v = float(r[vi].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 's': v * 1e6}.get(r[ui], v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-70s %6s %12s %7s %9s" % ("kernel", "n", "total_us", "share", "avg_us"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-70s %6d %12.1f %7.3f %9.1f" % (n[-70:], c, t, t / tot, t / c))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    ki = H.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[ki][:100])
    for m in METRICS:
        if m in H:
            i = H.index(m)
            print("%-75s %-12s %s" % (m, rows[1][i], "  ".join(r[i] for r in rows[2:])))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
