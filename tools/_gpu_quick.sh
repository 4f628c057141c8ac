# quick GPU check: parity tests, then the default bench line (summary only)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1
python bench.py --no-shooting-leg > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -20 gpurun_out/bench_quick.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e_ms", d["e2e"]["ms_per_step"], "batch1_ms", d["batch1"]["ms_per_step"])
print({k:v for k,v in d["roofline"].items() if k!="note"})
for k in ("families","kernel_families","breakdown"):
    if k in d: print(k, json.dumps(d[k])[:1500])
PY
