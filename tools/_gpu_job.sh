( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -8
( time python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>&1 | tail -4
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_default.json"))
print(d["ms_per_step"], d["value"], d["e2e"], d["batch1"]["ms_per_step"])
print(json.dumps(d.get("multiple_shooting_sharded"))[:400])
print({k:v for k,v in d["roofline"].items() if k!="note"})
PY
( time python bench.py --impl reference --steps 5 --warmup 2 ) 2>&1 | tail -5 | cut -c1-600
