#!/usr/bin/env python
"""Top stall hot spots of one kernel from `ncu --page source --csv` output (SASS view):
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv
    python tools/ncu_hotspots.py src.csv [N]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hdr_i + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break                      # only the first launch in the file
        if len(r) == len(hdr):
            data.append(r)
    total = sum(int(r[col["# Samples"]] or 0) for r in data)
    print(f"{path}: {total} samples")
    ranked = sorted(enumerate(data), key=lambda t: -int(t[1][col["# Samples"]] or 0))[:top]
    for i, r in sorted(ranked):
        n = int(r[col["# Samples"]] or 0)
        reasons = sorted(((int(r[col[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:2]
        rs = ", ".join(f"{s}:{c}" for c, s in reasons if c)
        print(f"{i:5d} {100.0 * n / total:5.1f}%  {r[col['Source']].strip()[:90]:90s} {rs}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
