#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total / mean duration and share of the profiled time.  Usage:
    python tools/summarize_launches.py gpurun_out/launches.csv [--steps] > profiles/rNN_launches.md
--steps: the capture covers a whole bench run; keep the two consecutive training steps with the largest kernel time (a step starts at
pack_kernel; its length is the most frequent distance between two pack_kernel launches)."""
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


def last_two_steps(rows):
    packs = [i for i, r in enumerate(rows) if "pack_kernel" in r[0]]
    gaps = [b - a for a, b in zip(packs, packs[1:])]
    if not gaps:
        return rows
    step = max(set(gaps), key=gaps.count)
    best = None
    for k in range(len(gaps) - 1):
        if gaps[k] == step and gaps[k + 1] == step:
            win = rows[packs[k]:packs[k] + 2 * step]
            t = sum(r[1] for r in win)
            if best is None or t >= best[0]:      # the bench's main workload, not its one-window leg
                best = (t, win)
    return best[1] if best else rows


def main(path, steps=False):
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
    if steps:
        rows = last_two_steps(rows)
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
    main(sys.argv[1], "--steps" in sys.argv)
