# dump_reference.jl - pins the oracle against the REAL reference arithmetic (GraphNetCore.jl) on a machine that
# has Julia.  UNTESTED HERE: this image has neither Julia nor GraphNetCore.jl (SURVEY.md section 8c), which is
# why oracle/mgn_oracle.py says "PARITY UNPINNED".  Test infrastructure only, like everything under oracle/.
#
#   julia --project=/path/to/MeshGraphNets.jl oracle/julia/dump_reference.jl tests/golden out_dir
#
# It replays tests/golden/cyl_small_inputs.npz / index_golden.npz (written by tests/golden/make_golden.py) through
# GraphNetCore's own functions - the ones MeshGraphNets.jl calls at src/graph.jl:26-38,87-96, src/solve.jl:200 and
# src/strategies.jl:421 - and writes out_dir/reference_golden.npz with the same keys as
# tests/golden/cyl_small_golden.npz / index_golden.npz.  Compare with
#   python tests/golden/compare_reference.py out_dir/reference_golden.npz
# Every mismatch names the recalled semantic of DESIGN.md section 5 that has to be flipped in oracle/mgn_oracle.py
# (and, through the failing GPU parity tests, in the kernels).
#
# Layout: Julia matrices are (features, entities) column-major == the [entities][features] row-major arrays in the
# .npz files, so every matrix is permuted on load and on store.
using NPZ, GraphNetCore, Lux, ComponentArrays, Zygote, Optimisers, Random

golden_dir, out_dir = ARGS[1], ARGS[2]
inp = npzread(joinpath(golden_dir, "cyl_small_inputs.npz"))
idx = npzread(joinpath(golden_dir, "index_golden.npz"))
jl(a::AbstractMatrix) = permutedims(a)          # [entities][features] -> (features, entities)
py(a::AbstractMatrix) = permutedims(a)

out = Dict{String, Any}()

# ---- integer path (bit exact): src/graph.jl:26-38 --------------------------------------------------------------
cells = jl(idx["cells"])                                             # 3 x C, 0-based as in the dataset
senders, receivers = triangles_to_edges(cells)
if 0 in senders || 0 in receivers                                    # src/graph.jl:31-34
    senders .+= 1; receivers .+= 1
end
out["senders"] = Int32.(senders); out["receivers"] = Int32.(receivers)
out["onehot"] = py(Float32.(one_hot(vec(idx["node_type"]), 7, 1)))
chain = Int32.(hcat([[i, i + 1] for i in 1:8]...))                  # src/dataset.jl:379-382
cs, cr = parse_edges(chain)
out["chain_senders"] = Int32.(cs); out["chain_receivers"] = Int32.(cr)
pos = jl(idx["pos"])
rel = pos[:, senders] .- pos[:, receivers]                           # src/graph.jl:35-36,49-52
out["edge_features"] = py(vcat(rel, mapslices(c -> Float32(sqrt(sum(abs2, Float64.(c)))), rel; dims = 1)))

# ---- float path: the model on the oracle's parameters ---------------------------------------------------------
model, ps0, st = build_model(Int(inp["node_in"]), Int(inp["edge_in"]) - 1, Int(inp["out_dim"]), Int(inp["mps"]),
    Int(inp["latent"]), Int(inp["hidden_layers"]), cpu_device())
ps = ComponentArray(ps0)
out["param_labels"] = join(ComponentArrays.labels(ps), "\n")         # check against mgn_model_param_layout
length(ps) == length(inp["params"]) || error("parameter count differs: $(length(ps)) vs $(length(inp["params"])): " *
    "the recalled MLP depth (hidden_layers + 2 Dense layers) or the LayerNorm placement is wrong")
ps = ComponentArray(Float32.(inp["params"]), getaxes(ps))            # flat order = recalled ComponentArray order
graph = FeatureGraph(jl(inp["nf"]), jl(inp["ef"]), Int32.(inp["senders"]), Int32.(inp["receivers"]))
y, _ = model(graph, ps, st)
out["out"] = py(y)

mgn = GraphNetwork(model, ps, st, nothing, nothing, nothing)
gs, loss = step!(mgn, graph, jl(inp["target"]), Int32.(inp["mask"]), mse_reduce)
g = collect(getdata(gs[1]))
out["loss"] = Float64(loss)
out["grad_norm"] = sqrt(sum(abs2, Float64.(g)))
out["grad_sample"] = Float64.(g[1:97:end])
dnf = Zygote.gradient(nf -> sum(abs2, first(model(FeatureGraph(nf, graph.ef, graph.senders, graph.receivers), ps, st)) .-
                                 jl(inp["target"])), graph.nf)[1]
out["dnf_sumsq_loss"] = py(dnf)                                      # a second differentiable functional of the output

# ---- Adam and the online normaliser ---------------------------------------------------------------------------
on = npzread(joinpath(golden_dir, "optim_norm_golden.npz"))
p = Float32.(on["adam_traj"][1, :]); stt = Optimisers.setup(Optimisers.Adam(1.0f-4), p)
traj = [copy(p)]
for t in 1:size(on["adam_grads"], 1)
    global stt, p
    stt, p = Optimisers.update(stt, p, Float32.(on["adam_grads"][t, :]))
    push!(traj, copy(p))
end
out["adam_traj"] = permutedims(hcat(traj...))
norm = NormaliserOnline(3, cpu_device())
out["norm_y"] = cat([py(norm(jl(on["norm_x"][i, :, :]))) for i in 1:size(on["norm_x"], 1)]...; dims = 3)

mkpath(out_dir)
npzwrite(joinpath(out_dir, "reference_golden.npz"), out)
println("wrote ", joinpath(out_dir, "reference_golden.npz"))
