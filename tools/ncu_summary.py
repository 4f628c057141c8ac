#!/usr/bin/env python
"""Markdown table of the key `ncu --set full` metrics of every launch in one or more .ncu-rep files
(read with `ncu -i rep --page raw --csv`).  Usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

COLS = [  # (header, metric name, scale, format)
    ("duration", "gpu__time_duration.sum", 1e-3, "{:.1f} us"),
    ("grid", "launch__grid_size", 1, "{:.0f}"),
    ("block", "launch__block_size", 1, "{:.0f}"),
    ("DRAM read", "dram__bytes_read.sum", 1e-6, "{:.1f} MB"),
    ("DRAM write", "dram__bytes_write.sum", 1e-6, "{:.1f} MB"),
    ("DRAM thr %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("L2 thr %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("L1/TEX thr %", "l1tex__throughput.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("issue active %", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("regs", "launch__registers_per_thread", 1, "{:.0f}"),
    ("dyn smem", "launch__shared_mem_per_block_dynamic", 1e-3, "{:.1f} KB"),
]
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr, units, data = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        vals = {}
        for h, m, sc, fmt in COLS:
            cands = [c for c in col if c == m or c.startswith(m)]
            if not cands:
                vals[h] = "-"
                continue
            i = col[cands[0]]
            try:
                v = float(r[i].replace(",", "")) * UNIT.get(units[i].split("/")[0], 1.0)
                vals[h] = fmt.format(v * sc)
            except ValueError:
                vals[h] = r[i]
        name = r[col["Kernel Name"]]
        yield name.split("(")[0].split("::")[-1], vals


def main(reps):
    print("| kernel | " + " | ".join(h for h, *_ in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for rep in reps:
        for name, vals in rows_of(rep):
            print(f"| {name} | " + " | ".join(vals[h] for h, *_ in COLS) + " |")


if __name__ == "__main__":
    main(sys.argv[1:])
