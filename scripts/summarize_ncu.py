#!/usr/bin/env python
"""Summarise ncu output into small text files for profiles/.
  launches: python scripts/summarize_ncu.py launches gpurun_out/launches.csv > profiles/x.txt
  full:     python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep > profiles/y.txt"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms summed kernel time "
          f"(ncu: cold-cache, serialised -- compare shares, not absolutes)")
    print(f"{'kernel':62s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{k[:62]:62s} {v[0]:8d} {v[1] / 1e3:10.2f} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path} (ncu --set full --clock-control none)")
    for r in rows[2:]:
        print(f"\n== {r[idx['Kernel Name']][:100]}")
        for k in KEYS:
            if k in idx:
                print(f"{k:72s} {r[idx[k]]:>16s} {units[idx[k]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
