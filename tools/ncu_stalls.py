#!/usr/bin/env python
"""Stall-reason totals of one launch in an `ncu --page source --csv` dump (SASS view), and the SASS lines with the most
shared-memory wavefronts.  Usage: python tools/ncu_stalls.py src.csv [launch_index]"""
import csv
import sys


def main(path, which=0, top=25):
    rows = list(csv.reader(open(path)))
    his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = his[which]
    hdr = rows[start]
    col = {h: i for i, h in enumerate(hdr)}
    end = his[which + 1] if which + 1 < len(his) else len(rows)
    data = [r for r in rows[start + 1:end] if len(r) == len(hdr)]

    def f(r, k):
        try:
            return float(r[col[k]] or 0)
        except ValueError:
            return 0.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(f(r, "# Samples") for r in data)
    print(f"launch {which}: {len(data)} SASS lines, {tot:.0f} samples")
    for s in sorted(stalls, key=lambda s: -sum(f(r, s) for r in data)):
        v = sum(f(r, s) for r in data)
        if v > 0.005 * tot:
            print(f"  {s[6:]:20s} {100 * v / tot:5.1f}%")
    w = sum(f(r, "L1 Wavefronts Shared") for r in data)
    print(f"shared wavefronts {w:.0f} (ideal {sum(f(r, 'L1 Wavefronts Shared Ideal') for r in data):.0f})")
    ranked = sorted(enumerate(data), key=lambda t: -f(t[1], "L1 Wavefronts Shared"))[:top]
    for i, r in sorted(ranked):
        print(f"  {i:5d} {100 * f(r, 'L1 Wavefronts Shared') / max(w, 1):5.1f}%  x{f(r, 'Instructions Executed'):9.0f}  "
              f"{r[col['Source']].strip()[:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
