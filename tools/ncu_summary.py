#!/usr/bin/env python
"""Summarise ncu outputs into text for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/launches.csv
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
           "smsp__inst_executed_op_shared_atom.sum",
           "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"')) if len(r) > 5]
    H, data = rows[0], rows[1:]
    ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        name = r[ki].split('(')[0].replace('tb::<unnamed>::', '').replace('void ', '')
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 's': v * 1e6}.get(r[ui], v)
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    print("%-58s %7s %12s %12s %7s" % ("kernel", "n", "mean us", "total ms", "share"))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %7d %12.1f %12.2f %7.3f" % (name[:58], n, us / n, us / 1e3, us / total))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[H.index("Kernel Name")][:160])
        for m in METRICS:
            if m in H:
                print("  %-72s %18s %s" % (m, r[H.index(m)], U[H.index(m)]))
        stalls = []
        for i, h in enumerate(H):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h))
                except ValueError:
                    pass
        print("  warp stall reasons (warps stalled per issue-active cycle, top 8):")
        for v, h in sorted(stalls, reverse=True)[:8]:
            print("    %6.2f  %s" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace(
                "smsp__warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
