python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_tests6.log 2>&1; tail -4 gpurun_out/r2_tests6.log
for pdl in 1 0; do
MGN_PDL=$pdl python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench6_pdl$pdl.json 2> gpurun_out/r2_bench6.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench6_pdl$pdl.json"))
print("pdl=$pdl", d["ms_per_step"], d["batch1"]["ms_per_step"])
for r in d["kernel_families"]: print(r["kernel"], round(r["ms_per_step"],3), round(r["hbm_frac"],3), round(r["tensor_frac"],3))
PY
done
python tools/bench_rollout.py bf16 2>/dev/null | cut -c1-400
