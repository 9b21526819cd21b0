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


if __name__ == "__main__" and sys.argv[1] in ("launches", "kernel"):
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])


def source(path, top=45):
    """per source line: warp instructions executed and stall samples (ncu --page source with --import-source on),
    first kernel of the report only"""
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur_file, hdr, per = None, None, collections.OrderedDict()
    nk = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Kernel Name":
            nk += 1
            if nk > 1:
                break
            continue
        if r[0] == "Line No":
            hdr = r
            ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[0].isdigit():           # a source line row: totals of its SASS
            key = (cur_file, int(r[0]))
            try:
                ent = per.setdefault(key, [0, 0, r[1].strip()[:110]])
                ent[0] += int(r[ii]); ent[1] += int(r[si])
            except ValueError:
                pass
    tot_i = sum(v[0] for v in per.values()) or 1
    tot_s = sum(v[1] for v in per.values()) or 1
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    print("%-22s %6s %6s  %s" % ("file:line", "inst%", "smpl%", "source"))
    for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %6.2f %6.2f  %s" % ("%s:%d" % (f, ln), 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s, v[2]))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[1] == "source":
    source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 45)
