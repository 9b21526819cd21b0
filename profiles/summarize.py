#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launches*.csv into the small text summaries committed under profiles/.
usage: python profiles/summarize.py launches <csv> | kernel <ncu-rep>"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki][:90], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    print("%-92s %5s %10s %10s %6s" % ("kernel", "n", "avg_us", "total_ms", "share"))
    for k, v in agg.items():
        print("%-92s %5d %10.1f %10.2f %5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e6, 100 * sum(v) / tot))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")][:110])
        for i, h in enumerate(hdr):
            if h in WANT:
                print("  %-86s %-10s %s" % (h, rows[1][i], vals[i]))
        pipes = []
        for i, h in enumerate(hdr):
            if (h.startswith("sm__inst_executed_pipe") and h.endswith(".avg.pct_of_peak_sustained_active")) or \
               ("issue_stalled" in h and h.endswith("_per_issue_active.ratio")):
                try:
                    pipes.append((float(vals[i].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(pipes, reverse=True)[:10]:
            print("    %10.3f  %s" % (v, h))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
