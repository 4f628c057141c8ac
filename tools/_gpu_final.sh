# Final single-GPU numbers of the round -> gpurun_out/r2f_*.json
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -6
python bench.py > gpurun_out/r2f_bench_bf16.json 2> gpurun_out/r2f_bench_bf16.err
python bench.py --batch 32 --no-shooting-leg --cpu-seconds 1 > gpurun_out/r2f_bench_bf16_w32.json 2> /dev/null
python bench.py --mode fp32 --batch 8 --no-shooting-leg --cpu-seconds 1 > gpurun_out/r2f_bench_fp32.json 2> /dev/null
python tools/bench_rollout.py 2>/dev/null | tail -1 > gpurun_out/r2f_rollout.json
MGN_FWD_PERSIST=2 python tools/bench_rollout.py 2>/dev/null | tail -1 > gpurun_out/r2f_rollout_persist.json
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2f_bench_reference.json 2>/dev/null
python - <<PY
import json
for f in ("r2f_bench_bf16","r2f_bench_bf16_w32","r2f_bench_fp32"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("batch1",{}).get("ms_per_step"), d.get("gpu_launches"))
for f in ("r2f_rollout","r2f_rollout_persist","r2f_bench_reference"):
    print(f, open(f"gpurun_out/{f}.json").read()[:500])
PY
