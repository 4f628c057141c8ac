# RHS / rollout latency of several library builds on one box (args: library files)
L=meshgraphnets.jl_b200/csrc/libmgn_b200.so
cp $L /tmp/lib_keep.so
for rep in 1 2; do
  for f in "$@"; do
    cp $f $L
    python tools/bench_rollout.py 2>/dev/null | tail -1 > /tmp/r.json
    python -c "
import json; d=json.load(open('/tmp/r.json')); print('$f', round(d['rhs_ms_plain'],4), round(d['rhs_ms_cuda_graph'],4), round(d['tsit5_50_steps_one_graph_ms'],1))"
  done
done
cp /tmp/lib_keep.so $L
