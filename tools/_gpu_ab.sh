# A/B on one box: $1 = alternative library (same ABI); prints step / family times for both, twice
L=meshgraphnets.jl_b200/csrc/libmgn_b200.so
cp $L /tmp/lib_b.so
for rep in 1 2; do
  for which in A B; do
    if [ $which = A ]; then cp $1 $L; else cp /tmp/lib_b.so $L; fi
    python bench.py --no-shooting-leg --cpu-seconds 1 > /tmp/ab.json 2>/dev/null
    python - <<PY
import json
d=json.load(open("/tmp/ab.json"))
print("$which", round(d["ms_per_step"],3), {f["kernel"]: round(f["ms_per_step"],3) for f in d["kernel_families"]}, "b1", round(d["batch1"]["ms_per_step"],3))
PY
  done
done
cp /tmp/lib_b.so $L
