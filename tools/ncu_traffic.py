#!/usr/bin/env python
"""profiles/ncu_traffic.json from an ncu CSV of dram__bytes_{read,write}.sum per launch:
average DRAM bytes per launch of each tensor-core kernel family (bench_roofline.py reports it as `traffic`)."""
import csv
import json
import sys
from collections import defaultdict

FAM = {"mlp_fwd_kernel": "tc_mlp_fwd", "mlp_bwd_chain_kernel": "tc_mlp_bwd", "mlp_bwd_input_kernel": "tc_dw"}


def main(path, out):
    lines = [l for l in open(path, newline="") if l.startswith('"')]
    per = defaultdict(lambda: defaultdict(float))
    for r in csv.DictReader(lines):
        if r["Metric Name"] not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        fam = next((f for k, f in FAM.items() if k in r["Kernel Name"]), None)
        if fam:
            per[fam][r["ID"]] += v
    res = {f: sum(d.values()) / len(d) for f, d in per.items()}
    res["_launches"] = {f: len(d) for f, d in per.items()}
    res["_source"] = path
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
