#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total / mean duration and share of the profiled time.  Usage:
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = name.replace("(anonymous namespace)::", "")
    name = re.sub(r"^void ", "", name)
    name = name.replace("<unnamed>::", "")
    m = re.match(r"([\w:]+)(<[^(]*>)?\(", name)
    if m:
        return m.group(1).split("::")[-1] + (m.group(2) or "")
    return name[:60]


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            v *= 1e3
        rows.append((r["Kernel Name"], v, r["Grid Size"], r["Block Size"]))
    agg = defaultdict(lambda: [0, 0.0, set()])
    for name, ns, grid, block in rows:
        a = agg[short(name)]
        a[0] += 1
        a[1] += ns
        a[2].add(f"{grid}x{block}")
    total = sum(a[1] for a in agg.values())
    print(f"source: {path}  ({len(rows)} launches, {total / 1e3:.1f} us of kernel time; cold-cache, serialised - compare shares)\n")
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / 1e3 / a[0]:.2f} | {100 * a[1] / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
