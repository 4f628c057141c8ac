python -m pytest tests/test_gpu_tc_parity.py tests/test_gpu_bench_config_parity.py tests/test_gpu_parity.py tests/test_partition.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_tests9.log 2>&1; tail -3 gpurun_out/r2_tests9.log
python bench.py --steps 10 --warmup 3 --no-shooting-leg > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench9.json"))
print(d["ms_per_step"], d["batch1"]["ms_per_step"], d["gpu_launches_per_step"])
for r in d["kernel_families"]: print(r["kernel"], round(r["ms_per_step"],3), round(r["hbm_frac"],3), round(r["tensor_frac"],3))
PY
