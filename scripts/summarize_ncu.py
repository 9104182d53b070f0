#!/usr/bin/env python
"""Summarises ncu outputs brought back in gpurun_out/ into small text files for profiles/.
  launches: python scripts/summarize_ncu.py launches gpurun_out/<tag>_launches.csv > profiles/<tag>_launches.md
  full:     python scripts/summarize_ncu.py full gpurun_out/<tag>_top.ncu-rep > profiles/<tag>_top.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def launches(fn):
    rows = []
    with open(fn) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^.*::", "", name.replace("(anonymous namespace)::", ""))
        rows.append((name, val * scale))
    agg = OrderedDict()
    for n, t in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.2f | %.1f%% |" % (n, c, t, t / c, 100 * t / tot))
    print("\ntotal %d launches, %.1f us (cold-cache, serialised: compare shares, not absolutes)" % (len(rows), tot))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block", "launch__occupancy_limit", "sm__pipe_fma_cycles_active",
        "smsp__inst_executed.sum", "l1tex__t_bytes", "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fma", "smsp__issue_active.avg.pct", "smsp__warp_issue_stalled"]


def full(fn):
    out = subprocess.run(["ncu", "-i", fn, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for row in rd[2:]:
        d = dict(zip(hdr, row))
        print("## %s  grid %s block %s" % (d.get("Kernel Name", "?")[:120], d.get("Grid Size"), d.get("Block Size")))
        for h, u, v in zip(hdr, units, row):
            if any(k in h for k in KEYS):
                print("  %-70s %s %s" % (h, v, u))
        print()


def table(fn):
    """One row per kernel FAMILY of an `ncu --set full` report (the instance with the longest duration): duration,
    tensor-pipe / issue / DRAM utilisation, DRAM bytes and rate, L2 hit rate, registers, grid."""
    out = open(fn).read() if fn.endswith(".csv") else subprocess.run(["ncu", "-i", fn, "--page", "raw", "--csv"],
                                                                     capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr = rd[0]
    col = lambda name: next((i for i, h in enumerate(hdr) if h == name), None)
    want = [("us", "gpu__time_duration.sum", 1e-3), ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
            ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1), ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
            ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("dram rd MB", "dram__bytes_read.sum", None),
            ("dram wr MB", "dram__bytes_write.sum", None), ("L2 hit %", "lts__t_sector_hit_rate.pct", 1),
            ("regs", "launch__registers_per_thread", 1), ("grid", "launch__grid_size", 1)]
    units = rd[1]
    best = OrderedDict()
    for row in rd[2:]:
        name = row[col("Kernel Name")]
        fam = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("iq::", ""))
        t = float(row[col("gpu__time_duration.sum")].replace(",", "") or 0)
        if fam not in best or t > best[fam][0]:
            best[fam] = (t, row)
    print("| kernel | " + " | ".join(w[0] for w in want) + " | GB/s |")
    print("|---|" + "---:|" * (len(want) + 1))
    for fam, (t, row) in sorted(best.items(), key=lambda kv: -kv[1][0]):
        cells, byts, dur = [], 0.0, None
        for label, metric, scale in want:
            i = col(metric)
            if i is None or row[i] == "":
                cells.append("-")
                continue
            v = float(row[i].replace(",", ""))
            u = units[i]
            if metric.startswith("dram__bytes"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                byts += v
            elif metric == "gpu__time_duration.sum":
                v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
                dur = v
            cells.append("%.1f" % v)
        print("| %s | " % fam[:60] + " | ".join(cells) + " | %.0f |" % (byts / dur * 1e3 if dur else 0))


if __name__ == "__main__":
    {"launches": launches, "full": full, "table": table}[sys.argv[1]](sys.argv[2])
