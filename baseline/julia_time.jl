# baseline/julia_time.jl - times the REFERENCE's own CPU implementation of the hot path (BASELINE.md section 4.5): one
# derivative-training step of MeshGraphNets.jl / GraphNetCore.jl on the CylinderFlow-shaped synthetic workload of
# SURVEY.md 8d, on the host cores of the machine it runs on.  This image has no Julia: bench.py --impl reference times the
# torch-CPU port instead (kind "port").  Where Julia + GraphNetCore.jl exist:
#     julia --project=<MeshGraphNets.jl checkout> -t auto baseline/julia_time.jl [steps]
# prints one JSON line in bench.py's reference-arm format with kind = "reference".
using MeshGraphNets, GraphNetCore, Lux, Optimisers, Random, Statistics, LinearAlgebra, Printf
import MeshGraphNets: create_base_graph, build_graph

const NX, NY, MPS, LATENT, HIDDEN = 65, 29, 15, 128, 2

function cylinder_flow_mesh(nx, ny; lx = 1.6f0, ly = 0.41f0)
    xs = range(0f0, lx; length = nx); ys = range(0f0, ly; length = ny)
    pos = Float32[(d == 1 ? xs[i] : ys[j]) for d in 1:2, j in 1:ny, i in 1:nx]      # node id = (i-1)*ny + (j-1), 0-based
    pos = reshape(pos, 2, nx * ny)
    id(i, j) = Int32((i - 1) * ny + (j - 1))
    cells = Int32[]
    for tri in 1:2, i in 1:(nx - 1), j in 1:(ny - 1)
        a, b, c, d = id(i, j), id(i + 1, j), id(i + 1, j + 1), id(i, j + 1)
        append!(cells, tri == 1 ? (a, b, c) : (a, c, d))
    end
    nt = zeros(Int32, ny, nx); nt[1, :] .= 6; nt[end, :] .= 6; nt[:, 1] .= 4; nt[:, end] .= 5; nt[2:(end - 1), 2] .= 1
    pos, reshape(cells, 3, :), vec(nt)
end

function main(steps)
    pos, cells, nt = cylinder_flow_mesh(NX, NY)
    N = size(pos, 2)
    rng = Random.MersenneTwister(1234)
    vel = randn(rng, Float32, 2, N, 65)
    data = Dict{String,Any}("node_type" => reshape(nt, 1, N, 1), "mesh_pos" => reshape(pos, 2, N, 1),
        "cells" => reshape(cells, 3, :, 1), "velocity" => vel[:, :, 1:64], "target|velocity" => vel[:, :, 2:65])
    node_type, senders, receivers, edge_features = create_base_graph(data, 6, 0, cpu_device())   # src/graph.jl:25
    model, ps, st = GraphNetCore.build_model(2 + 7, 2, 2, MPS, LATENT, HIDDEN, cpu_device())
    e_norm = NormaliserOnline(3, cpu_device()); o_norm = Dict("velocity" => NormaliserOnline(2, cpu_device()))
    n_norm = Dict{String,Any}("velocity" => NormaliserOnline(2, cpu_device()), "node_type" => NormaliserOfflineMinMax(0f0, 6f0))
    mgn = GraphNetwork(model, ps, st, e_norm, n_norm, o_norm)
    opt_state = Optimisers.setup(Optimisers.Adam(1f-4), mgn.ps)
    mask = Int32.(findall(x -> x in (0, 5), nt))
    function one(t)
        target = o_norm["velocity"]((data["target|velocity"][:, :, t] .- data["velocity"][:, :, t]) ./ 0.01f0)   # strategies.jl:399-410
        graph = build_graph(mgn, data, ["velocity"], t, node_type, edge_features, senders, receivers)                # graph.jl:75
        gs, loss = GraphNetCore.step!(mgn, graph, target, mask, GraphNetCore.mse_reduce)                             # strategies.jl:421
        for g in gs
            opt_state, mgn.ps = Optimisers.update(opt_state, mgn.ps, g)                                              # MeshGraphNets.jl:374-378
        end
        loss
    end
    one(1)                                           # compile
    t0 = time(); for s in 1:steps; one(1 + s % 60); end; el = time() - t0
    E = length(senders)
    @printf("{\\"impl\\": \\"reference\\", \\"metric\\": \\"mp_step_edges_per_sec_train\\", \\"value\\": %.6g, \\"unit\\": \\"edges/s\\", \\"steps\\": %d, \\"ms_per_step\\": %.4f, \\"cpu_baseline\\": {\\"kind\\": \\"reference\\", \\"cores\\": %d, \\"sample\\": \\"%d batch-1 derivative-training steps, Julia %s\\"}}\\n",
        E * MPS * steps / el, steps, 1000 * el / steps, Threads.nthreads(), steps, string(VERSION))
end

main(length(ARGS) > 0 ? parse(Int, ARGS[1]) : 20)
