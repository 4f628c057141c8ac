# GraphNetCoreB200.jl - the binding a MeshGraphNets.jl maintainer adds to run the Encode-Process-Decode
# hot path on libmgn_b200.so (include/mgn_b200.h).  UNTESTED HERE: this image has no Julia; the same
# ABI is exercised from Python ctypes (meshgraphnets.jl_b200/_lib.py) and by tests/.
#
# It provides the GraphNetCore names MeshGraphNets.jl uses on this path (docs/src/graph_net_core.md:5-36):
#   mgn.model(graph, ps, st)            src/solve.jl:200
#   step!(mgn, graph, target, mask, mse_reduce)   src/strategies.jl:421
# plus a ChainRulesCore.rrule so that Zygote / SciMLSensitivity's ZygoteVJP differentiate the model call
# w.r.t. ps and graph.nf (src/strategies.jl:183-194).  Everything else of GraphNetCore (FeatureGraph,
# normaliser structs, load/save!) keeps its Julia definition; only the arithmetic moves.
module GraphNetCoreB200

using CUDA, ChainRulesCore

const LIB = get(ENV, "MGN_B200_LIB", "libmgn_b200.so")

struct MgnConfig            # mirrors mgn_model_config
    node_in::Int32; edge_in::Int32; out_dim::Int32; latent::Int32
    mps::Int32; hidden_layers::Int32; ln_eps::Float32; compute_mode::Int32
end

function check(status::Int32)
    status == 0 && return
    buf = Vector{UInt8}(undef, 1024)
    ccall((:mgn_last_error, LIB), Int32, (Ptr{UInt8}, Csize_t), buf, 1024)
    error("libmgn_b200: ", unsafe_string(pointer(buf)))
end

mutable struct B200Model     # stands in for the Lux chain held in GraphNetwork.model
    handle::Ptr{Cvoid}
    cfg::MgnConfig
    ws::Dict{Tuple{Ptr{Cvoid},Bool},CuVector{UInt8}}
end

function B200Model(node_in, edge_in, out_dim, mps, layer_size, hidden_layers; bf16 = true)
    cfg = MgnConfig(node_in, edge_in, out_dim, layer_size, mps, hidden_layers, 1f-5, bf16 ? 1 : 0)
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:mgn_model_create, LIB), Int32, (Ref{MgnConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    m = B200Model(h[], cfg, Dict())
    finalizer(x -> ccall((:mgn_model_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), m)
    m
end

# one mgn_graph per (senders, receivers) pair: the FeatureGraphs of a trajectory share them
const GRAPHS = IdDict{Any,Ptr{Cvoid}}()
function graph_handle(g)      # g::GraphNetCore.FeatureGraph
    get!(GRAPHS, g.senders) do
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:mgn_graph_create, LIB), Int32,
            (Int64, Int64, CuPtr{Int32}, CuPtr{Int32}, Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
            size(g.nf, 2), length(g.senders), g.senders, g.receivers, 1, CUDA.stream().handle, h))
        h[]
    end
end

function workspace(m::B200Model, gh, training)
    get!(m.ws, (gh, training)) do
        n = Ref{Csize_t}(0)
        check(ccall((:mgn_workspace_bytes, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{Csize_t}),
            m.handle, gh, training, n))
        CUDA.zeros(UInt8, n[])
    end
end

# ps is the flat Float32 parameter vector (ComponentArray data); mgn_model_param_layout gives the table
function forward(m::B200Model, g, ps::CuVector{Float32}; training = false)
    gh = graph_handle(g); ws = workspace(m, gh, training)
    out = CUDA.zeros(Float32, m.cfg.out_dim, size(g.nf, 2))
    check(ccall((:mgn_forward, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{UInt8}, Csize_t, Int32, Ptr{Cvoid}),
        m.handle, gh, ps, g.nf, g.ef, out, ws, length(ws), training, CUDA.stream().handle))
    out
end

function backward(m::B200Model, g, ps, dout; want_dnf = true)
    gh = graph_handle(g); ws = workspace(m, gh, true)
    dps = similar(ps); dnf = want_dnf ? similar(g.nf) : CuPtr{Float32}(0)
    check(ccall((:mgn_backward, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, CuPtr{UInt8}, Csize_t, Ptr{Cvoid}),
        m.handle, gh, ps, g.nf, g.ef, dout, dps, dnf, ws, length(ws), CUDA.stream().handle))
    dps, dnf
end

# `output, st = mgn.model(graph, ps, st)`  (src/solve.jl:200)
(m::B200Model)(g, ps, st) = (forward(m, g, ps), st)

function ChainRulesCore.rrule(m::B200Model, g, ps, st)
    out = forward(m, g, ps; training = true)
    function pullback(ȳ)
        dps, dnf = backward(m, g, ps, CuArray{Float32}(unthunk(ȳ[1])))
        g̃ = Tangent{typeof(g)}(nf = dnf)          # only node features carry the ODE state
        return NoTangent(), g̃, dps, NoTangent()
    end
    (out, st), pullback
end

# GraphNetCore.step!(mgn, graph, target, mask, mse_reduce) -> (gs, loss)   (src/strategies.jl:421)
function step!(mgn, g, target::CuMatrix{Float32}, mask::CuVector{Int32}, _loss)
    m, ps = mgn.model, mgn.ps
    out = forward(m, g, ps; training = true)
    loss = CUDA.zeros(Float32, 1); dout = similar(out)
    check(ccall((:mgn_loss_mse_masked, LIB), Int32,
        (CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Int32}, Int64, Int32, CuPtr{Float32},
         CuPtr{Float32}, Ptr{Cvoid}),
        out, target, size(out, 2), size(out, 1), mask, length(mask), 1, loss, dout, CUDA.stream().handle))
    dps, _ = backward(m, g, ps, dout; want_dnf = false)
    (dps,), loss                                  # gs is iterated at src/MeshGraphNets.jl:375-377
end

# ---- NeuralODE callers (include/mgn_b200.h, "NeuralODE callers") -----------------------------------------------
# The solver strategies (src/strategies.jl:229-386) need nothing beyond the rrule above: OrdinaryDiffEq drives
# ode_func_train and SciMLSensitivity's ZygoteVJP pulls it back through `rrule(::B200Model, ...)`.  The entry points
# below let the REST of the right-hand side stay on the device without allocating broadcasts: the inflow overwrite
# (src/solve.jl:104-107), `.* val_mask` (:218), Runge-Kutta stage combinations, and the shooting losses
# (src/strategies.jl:263-286, :367-383).  meshgraphnets.jl_b200/shooting.py is the executable specification of how
# they compose into a lock-step MultipleShooting step (all intervals as one block-diagonal graph).
stream() = CUDA.stream().handle

# y = x + sum_j coef[j] * k[j]   (n_terms <= 8)
function ode_lincomb!(y::CuArray{Float32}, x::CuArray{Float32}, ks::Vector{<:CuArray{Float32}}, coef::Vector{Float32})
    ptrs = [Ptr{Cvoid}(UInt(pointer(k))) for k in ks]      # device addresses, passed by value in a HOST array
    GC.@preserve ks check(ccall((:mgn_ode_lincomb, LIB), Int32,
        (CuPtr{Float32}, Ptr{Ptr{Cvoid}}, Ptr{Float32}, Int32, Int64, CuPtr{Float32}, Ptr{Cvoid}),
        x, ptrs, coef, length(ks), length(y), y, stream()))
    y
end

# bx[inflow_mask] = data[inflow_mask]   (mask::CuArray{UInt8}; src == nothing gives the transposed Jacobian)
function masked_overwrite!(y, x, src, mask::CuArray{UInt8})
    check(ccall((:mgn_masked_overwrite, LIB), Int32,
        (CuPtr{Float32}, CuPtr{Float32}, CuPtr{UInt8}, Int64, CuPtr{Float32}, Ptr{Cvoid}),
        x, src === nothing ? CuPtr{Float32}(0) : src, mask, length(y), y, stream()))
    y
end

vec_mul!(y, a, b) = (check(ccall((:mgn_vec_mul, LIB), Int32,
    (CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}, Ptr{Cvoid}), a, b, length(y), y, stream())); y)

# mean((gt .- pred).^2 .* val_mask) over n_saves saved states and its gradient w.r.t. pred
function shooting_mse!(loss::CuVector{Float32}, dpred, pred, gt, val_mask; accumulate = false)
    n_saves = size(pred, 3)
    check(ccall((:mgn_shooting_mse, LIB), Int32,
        (CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Int64, Float32, Int32, CuPtr{Float32},
         CuPtr{Float32}, Ptr{Cvoid}),
        pred, gt, val_mask, n_saves, length(val_mask), 1.0f0 / length(pred), accumulate, loss, dpred, stream()))
    loss
end

# loss += w * sum(abs, a .- b); da .+= w .* sign.(a .- b)
shooting_continuity!(loss, da, a, b, w) = (check(ccall((:mgn_shooting_continuity, LIB), Int32,
    (CuPtr{Float32}, CuPtr{Float32}, Int64, Float32, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
    a, b, length(a), Float32(w), loss, da, stream())); loss)

# ---- lock-step MultipleShooting step (transcription of meshgraphnets.jl_b200/shooting.py; UNTESTED like the rest) ----
# All K shooting intervals advance together as ONE block-diagonal graph (K copies of the mesh), fixed-step explicit
# Euler (the examples/cylinder_flow configuration; the Python file also carries RK4 / Tsit5 tableaus), exact reverse
# sweep through `backward` (d_params and d_nf).  Normaliser statistics are frozen during the step.
#   rhs_forward(X, idx; training)  -> dX/dt for the stacked state X (S x K*N), idx[k] = data index of interval k's inflow
#   rhs_backward(dY)               -> (d_params, dX) of the matching training forward
# are closures the caller builds from build_graph / inverse_data exactly as ode_step does (src/solve.jl:188-219), with
# the inflow overwrite (masked_overwrite!) in front and `.* val_mask` (vec_mul!) behind, over the K-fold repeated
# senders / receivers (copy k shifted by k*N) - see DeviceRhs in shooting.py.
function multiple_shooting_step(rhs_forward, rhs_backward, ps::CuVector{Float32}, gt::CuArray{Float32,3},
        val_mask::CuMatrix{Float32}, tsteps::AbstractVector{Float32}, dt::Float32, interval_size::Int,
        continuity_term)
    S, N, _ = size(gt)
    ranges = [i:min(length(tsteps), i + interval_size - 1) for i in 1:(interval_size - 1):(length(tsteps) - 1)]  # strategies.jl:346-347
    K, M = length(ranges), maximum(length.(ranges)) - 1
    firsts = first.(ranges)
    idx_of(t) = min(floor(Int, t / dt) + 1, size(gt, 3))                       # src/solve.jl:106
    X = reduce(hcat, [gt[:, :, f] for f in firsts])                            # S x K*N, u0 of every interval
    saves, chk = [X], Tuple{typeof(X),Vector{Int}}[]
    for m in 0:(M - 1)                                                         # forward sweep, states checkpointed
        idx = [idx_of(tsteps[min(f + m, length(tsteps))]) for f in firsts]
        push!(chk, (X, idx))
        k1 = rhs_forward(X, idx; training = false)
        X = ode_lincomb!(similar(X), X, [k1], [dt])
        push!(saves, X)
    end
    loss = CUDA.zeros(Float32, 1)
    dsaves = [CUDA.zeros(Float32, S, K * N) for _ in 0:M]
    for (k, rg) in enumerate(ranges)                                           # strategies.jl:367-383
        cols = ((k - 1) * N + 1):(k * N)
        P = cat([saves[m][:, cols] for m in 1:length(rg)]...; dims = 3)
        dP = similar(P)
        shooting_mse!(loss, dP, P, gt[:, :, rg], val_mask; accumulate = true)
        if k < K                                                               # continuity term of interval k + 1
            last = dP[:, :, end]
            shooting_continuity!(loss, last, P[:, :, end], gt[:, :, first(ranges[k + 1])], continuity_term)
            dP[:, :, end] .= last
        end
        for m in 1:length(rg)
            dsaves[m][:, cols] .= dP[:, :, m]
        end
    end
    g, lam = CUDA.zeros(Float32, length(ps)), copy(dsaves[end])
    for n in M:-1:1                                                            # reverse sweep (Euler: one stage)
        x, idx = chk[n]
        rhs_forward(x, idx; training = true)
        gi, dx = rhs_backward(ode_lincomb!(similar(lam), CUDA.zeros(Float32, size(lam)), [lam], [dt]))
        ode_lincomb!(g, g, [gi], [1.0f0])
        ode_lincomb!(lam, lam, [dx], [1.0f0])
        n > 1 && ode_lincomb!(lam, lam, [dsaves[n]], [1.0f0])
    end
    (g,), loss
end

end # module
