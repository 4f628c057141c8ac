# Run this where Julia + GraphNetCore.jl exist to pin the oracle against the real reference:
#   julia --project=/path/to/MeshGraphNets.jl oracle/julia/dump_reference.jl tests/golden/cyl_small_inputs.npz out_dir
# It rebuilds the model of tests/golden (same flat parameters), runs mgn.model and step!, and writes
# out.npy / loss.npy / grads.npy for tests/test_golden.py --reference-dir out_dir to diff.
# UNTESTED HERE (no Julia in this image); kept deliberately tiny.
using GraphNetCore, NPZ, Lux, ComponentArrays
inp = npzread(ARGS[1]); outdir = ARGS[2]; mkpath(outdir)
model = GraphNetCore.build_model(Int(inp["node_in"]), Int(inp["edge_in"]) - 1, Int(inp["out_dim"]),
                                 Int(inp["mps"]), Int(inp["latent"]), Int(inp["hidden_layers"]), cpu_device())
ps0, st = Lux.setup(Lux.Random.default_rng(), model)
ps = ComponentArray(ps0); @assert length(ps) == length(inp["params"]) "parameter count differs: check DESIGN.md section 5"
ps .= inp["params"]                       # flat order must match mgn_model_param_layout (DESIGN.md section 1)
graph = GraphNetCore.FeatureGraph(permutedims(inp["nf"]), permutedims(inp["ef"]), inp["senders"], inp["receivers"])
out, _ = model(graph, ps, st)
npzwrite(joinpath(outdir, "out.npy"), permutedims(out))
