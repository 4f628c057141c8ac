# compare several library builds on one box: args = library files (same ABI); prints step / family times, two rounds
L=meshgraphnets.jl_b200/csrc/libmgn_b200.so
cp $L /tmp/lib_keep.so
for rep in 1 2; do
  for f in "$@"; do
    cp $f $L
    python bench.py --no-shooting-leg --cpu-seconds 1 > /tmp/ab.json 2>/dev/null
    python - <<PY
import json
d=json.load(open("/tmp/ab.json"))
print("$f", round(d["ms_per_step"],3), {f["kernel"]: round(f["ms_per_step"],3) for f in d["kernel_families"]}, "b1", round(d["batch1"]["ms_per_step"],3))
PY
  done
done
cp /tmp/lib_keep.so $L
