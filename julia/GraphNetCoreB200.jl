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

using CUDA, ChainRulesCore, ComponentArrays

const LIB = get(ENV, "MGN_B200_LIB", "libmgn_b200.so")

struct MgnConfig            # mirrors mgn_model_config (ABI version 2)
    node_in::Int32; edge_in::Int32; out_dim::Int32; latent::Int32
    mps::Int32; hidden_layers::Int32; ln_eps::Float32; compute_mode::Int32
    dense_layers::Int32            # 0 = hidden_layers + 2 (recalled build_mlp)
    ln_scale_first::Int32          # 0 = LayerNorm (bias, scale) order in the flat vector (recalled Lux 0.5)
    aggregate_post_residual::Int32 # 0 = scatter-sum the new messages (recalled); 1 = scatter-sum the updated edge latent
end

# `mgn.ps` is a ComponentArray (src/MeshGraphNets.jl:288,376; handed to ODEProblem at src/strategies.jl:187): the library
# takes its flat Float32 data.  check_layout(m, ps) compares mgn_model_param_layout with ComponentArrays.labels once.
flat(ps::CuVector{Float32}) = ps
flat(ps::ComponentArray) = ComponentArrays.getdata(ps)::CuVector{Float32}
retangent(ps::CuVector{Float32}, d) = d
retangent(ps::ComponentArray, d) = ComponentArray(d, ComponentArrays.getaxes(ps))

function check(status::Int32)
    status == 0 && return
    buf = Vector{UInt8}(undef, 1024)
    ccall((:mgn_last_error, LIB), Int32, (Ptr{UInt8}, Csize_t), buf, 1024)
    error("libmgn_b200: ", unsafe_string(pointer(buf)))
end

mutable struct B200Model     # stands in for the Lux chain held in GraphNetwork.model
    handle::Ptr{Cvoid}
    cfg::MgnConfig
    ws_infer::Dict{Ptr{Cvoid},CuVector{UInt8}}          # one inference workspace per graph
    ws_pool::Dict{Ptr{Cvoid},Vector{CuVector{UInt8}}}   # free TRAINING workspaces per graph (see rrule)
end

function B200Model(node_in, edge_in, out_dim, mps, layer_size, hidden_layers; bf16 = true, dense_layers = 0,
        ln_scale_first = false)
    cfg = MgnConfig(node_in, edge_in, out_dim, layer_size, mps, hidden_layers, 1f-5, bf16 ? 1 : 0, dense_layers,
        ln_scale_first ? 1 : 0, 0)
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:mgn_model_create, LIB), Int32, (Ref{MgnConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    m = B200Model(h[], cfg, Dict(), Dict())
    finalizer(x -> ccall((:mgn_model_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), m)
    m
end

# The flat layout the library expects against the ComponentArray the caller holds (run once after build_model / load).
struct ParamEntry; name::NTuple{48,UInt8}; offset::Int64; rows::Int32; cols::Int32; end
function param_layout(m::B200Model)
    n = Ref{Int32}(0)
    check(ccall((:mgn_model_param_layout, LIB), Int32, (Ptr{Cvoid}, Ptr{ParamEntry}, Int32, Ref{Int32}), m.handle, C_NULL, 0, n))
    e = Vector{ParamEntry}(undef, n[])
    check(ccall((:mgn_model_param_layout, LIB), Int32, (Ptr{Cvoid}, Ptr{ParamEntry}, Int32, Ref{Int32}), m.handle, e, n[], n))
    [(String(UInt8[c for c in x.name if c != 0]), x.offset, x.rows, x.cols) for x in e]
end
function check_layout(m::B200Model, ps::ComponentArray)
    sizes = [r * c for (_, _, r, c) in param_layout(m)]
    sum(sizes) == length(ps) || error("libmgn_b200: parameter count mismatch: ", sum(sizes), " vs ", length(ps))
    # per-leaf sizes in memory order; a mismatch means one of the recalled switches of MgnConfig has to be flipped
    leaves = [length(getproperty(ps, k)) for k in propertynames(ps)]
    return sizes, leaves
end

# one mgn_graph per (senders, receivers) pair: the FeatureGraphs of a trajectory share them.  Weak keys: when the index
# array of a trajectory is collected, the handle's finalizer destroys the device CSR.
mutable struct GraphHandle
    h::Ptr{Cvoid}
end
const GRAPHS = WeakKeyDict{Any,GraphHandle}()
function graph_handle(g)      # g::GraphNetCore.FeatureGraph
    gh = get!(GRAPHS, g.senders) do
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:mgn_graph_create, LIB), Int32,
            (Int64, Int64, CuPtr{Int32}, CuPtr{Int32}, Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
            size(g.nf, 2), length(g.senders), g.senders, g.receivers, 1, CUDA.stream().handle, h))
        finalizer(x -> ccall((:mgn_graph_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), GraphHandle(h[]))
    end
    gh.h
end

function workspace_bytes(m::B200Model, gh, training)
    n = Ref{Csize_t}(0)
    check(ccall((:mgn_workspace_bytes, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ref{Csize_t}), m.handle, gh, training, n))
    n[]
end
infer_workspace(m::B200Model, gh) = get!(() -> CUDA.zeros(UInt8, workspace_bytes(m, gh, false)), m.ws_infer, gh)
# A TRAINING workspace holds the activations of ONE forward until its pullback has run.  Sensitivity algorithms may
# issue several model calls before pulling any of them back (ZygoteAdjoint / ReverseDiff over a discrete solve), so every
# rrule invocation checks a workspace out of the pool and its pullback returns it.
function checkout!(m::B200Model, gh)
    pool = get!(() -> CuVector{UInt8}[], m.ws_pool, gh)
    isempty(pool) ? CUDA.zeros(UInt8, workspace_bytes(m, gh, true)) : pop!(pool)
end
checkin!(m::B200Model, gh, ws) = push!(get!(() -> CuVector{UInt8}[], m.ws_pool, gh), ws)

function forward(m::B200Model, g, ps; ws = nothing)
    gh = graph_handle(g)
    training = ws !== nothing
    w = training ? ws : infer_workspace(m, gh)
    out = CUDA.zeros(Float32, m.cfg.out_dim, size(g.nf, 2))
    check(ccall((:mgn_forward, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{UInt8}, Csize_t, Int32, Ptr{Cvoid}),
        m.handle, gh, flat(ps), g.nf, g.ef, out, w, length(w), training, CUDA.stream().handle))
    out
end

function backward(m::B200Model, g, ps, dout, ws; want_dnf = true)
    gh = graph_handle(g)
    p = flat(ps)
    dps = similar(p); dnf = want_dnf ? similar(g.nf) : CuPtr{Float32}(0)
    check(ccall((:mgn_backward, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, CuPtr{UInt8}, Csize_t, Ptr{Cvoid}),
        m.handle, gh, p, g.nf, g.ef, dout, dps, dnf, ws, length(ws), CUDA.stream().handle))
    dps, dnf
end

# `output, st = mgn.model(graph, ps, st)`  (src/solve.jl:200)
(m::B200Model)(g, ps, st) = (forward(m, g, ps), st)

function ChainRulesCore.rrule(m::B200Model, g, ps, st)
    gh = graph_handle(g)
    ws = checkout!(m, gh)                          # this call's own activations
    out = forward(m, g, ps; ws = ws)
    function pullback(ȳ)
        dps, dnf = backward(m, g, ps, CuArray{Float32}(unthunk(ȳ[1])), ws)
        checkin!(m, gh, ws)                        # a second pullback of the same call would need a fresh forward
        g̃ = Tangent{typeof(g)}(nf = dnf)          # only node features carry the ODE state
        return NoTangent(), g̃, retangent(ps, dps), NoTangent()
    end
    (out, st), pullback
end

# GraphNetCore.step!(mgn, graph, target, mask, mse_reduce) -> (gs, loss)   (src/strategies.jl:421)
function step!(mgn, g, target::CuMatrix{Float32}, mask::CuVector{Int32}, _loss)
    m, ps = mgn.model, mgn.ps
    gh = graph_handle(g); ws = checkout!(m, gh)
    out = forward(m, g, ps; ws = ws)
    loss = CUDA.zeros(Float32, 1); dout = similar(out)
    check(ccall((:mgn_loss_mse_masked, LIB), Int32,
        (CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Int32}, Int64, Int32, CuPtr{Float32},
         CuPtr{Float32}, Ptr{Cvoid}),
        out, target, size(out, 2), size(out, 1), mask, length(mask), 1, loss, dout, CUDA.stream().handle))
    dps, _ = backward(m, g, ps, dout, ws; want_dnf = false)
    checkin!(m, gh, ws)
    (retangent(ps, dps),), loss                   # gs is iterated at src/MeshGraphNets.jl:375-377
end

# ---- multi-GPU transport (include/mgn_b200.h, "multi-GPU transport"): one Julia process (or task) per GPU ------------
# Bootstrap: rank 0 calls dp_unique_id() and hands the 128 bytes to the others (MPI.bcast, a shared file, Distributed.jl
# remotecall); every rank then builds its Communicator on ITS device (CUDA.device!(rank), src/MeshGraphNets.jl:257).
mutable struct Communicator
    h::Ptr{Cvoid}; rank::Int; world::Int
end
function dp_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:mgn_dp_unique_id, LIB), Int32, (Ptr{UInt8},), id)); id
end
function Communicator(id::Vector{UInt8}, rank, world)
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:mgn_dp_init, LIB), Int32, (Ptr{UInt8}, Int32, Int32, Ref{Ptr{Cvoid}}), id, rank, world, h))
    finalizer(c -> ccall((:mgn_dp_finalize, LIB), Int32, (Ptr{Cvoid},), c.h), Communicator(h[], rank, world))
end
allreduce!(c::Communicator, x::CuArray{Float32}; mean = false) = (check(ccall((:mgn_dp_allreduce, LIB), Int32,
    (Ptr{Cvoid}, CuPtr{Float32}, Int64, Int32, Ptr{Cvoid}), c.h, x, length(x), mean ? 1 : 0, CUDA.stream().handle)); x)
# online-normaliser statistics after every rank accumulated its own window on top of the common `prev`
allreduce_normaliser!(c::Communicator, state::CuVector{Float32}, prev::CuVector{Float32}) = (check(ccall(
    (:mgn_dp_allreduce_normaliser, LIB), Int32, (Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, Int32, Ptr{Cvoid}),
    c.h, state, prev, length(state), CUDA.stream().handle)); state)
# halo rows of a partitioned mesh: send_rows[p] rows go to rank p, recv_rows[p] rows come from rank p
halo_exchange!(c::Communicator, recv::CuArray{UInt8}, send::CuArray{UInt8}, send_rows::Vector{Int64},
    recv_rows::Vector{Int64}, row_bytes) = (check(ccall((:mgn_halo_exchange, LIB), Int32,
    (Ptr{Cvoid}, CuPtr{UInt8}, Ptr{Int64}, CuPtr{UInt8}, Ptr{Int64}, Int64, Ptr{Cvoid}),
    c.h, send, send_rows, recv, recv_rows, row_bytes, CUDA.stream().handle)); recv)

struct AdamConfig   # mirrors mgn_adam_config
    lr::Float32; beta1::Float32; beta2::Float32; eps::Float32
    m::CuPtr{Float32}; v::CuPtr{Float32}; state16::CuPtr{Cvoid}
end
# step! fused with what train_mgn! does with its result (src/MeshGraphNets.jl:374-378): the gradient buckets are
# all-reduced (mean over `comm`) and fed to Adam on a side stream while the rest of the backward pass runs.
# opt = (lr, beta1, beta2, eps, m::CuVector, v::CuVector, state16::CuVector{UInt8} of 16 zero bytes)
function step_dp!(mgn, g, target::CuMatrix{Float32}, mask::CuVector{Int32}; comm = nothing, opt = nothing, n_buckets = 6)
    m, ps = mgn.model, flat(mgn.ps)
    gh = graph_handle(g); ws = checkout!(m, gh)
    out = forward(m, g, ps; ws = ws)
    loss = CUDA.zeros(Float32, 1); dout = similar(out); dps = similar(ps)
    check(ccall((:mgn_loss_mse_masked, LIB), Int32,
        (CuPtr{Float32}, CuPtr{Float32}, Int64, Int32, CuPtr{Int32}, Int64, Int32, CuPtr{Float32},
         CuPtr{Float32}, Ptr{Cvoid}),
        out, target, size(out, 2), size(out, 1), mask, length(mask), 1, loss, dout, CUDA.stream().handle))
    cfg = opt === nothing ? C_NULL : Ref(AdamConfig(opt.lr, opt.beta1, opt.beta2, opt.eps, pointer(opt.m), pointer(opt.v),
        reinterpret(CuPtr{Cvoid}, pointer(opt.state16))))
    GC.@preserve opt check(ccall((:mgn_backward_dp, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{UInt8}, Csize_t, Ptr{Cvoid}, Ptr{AdamConfig}, Int32, Ptr{Cvoid}),
        m.handle, gh, ps, g.nf, g.ef, dout, dps, CuPtr{Float32}(0), ws, length(ws),
        comm === nothing ? C_NULL : comm.h, cfg, n_buckets, CUDA.stream().handle))
    checkin!(m, gh, ws)
    (dps,), loss
end

# ---- create_base_graph on the device (src/graph.jl:25-55 for a trajectory already resident in HBM) -------------------
function triangles_to_edges_device(cells::CuMatrix{Int32})        # 3 x C
    C = size(cells, 2); s = CUDA.zeros(Int32, 6C); r = CUDA.zeros(Int32, 6C); n = Ref{Int64}(0)
    check(ccall((:mgn_triangles_to_edges_device, LIB), Int32, (CuPtr{Int32}, Int64, CuPtr{Int32}, CuPtr{Int32}, Ref{Int64},
        Ptr{Cvoid}), cells, C, s, r, n, CUDA.stream().handle))
    s[1:n[]], r[1:n[]]
end
function edge_features_device(pos::CuMatrix{Float32}, s::CuVector{Int32}, r::CuVector{Int32})   # dim x N -> (dim+1) x E
    out = CUDA.zeros(Float32, size(pos, 1) + 1, length(s))
    check(ccall((:mgn_edge_features_device, LIB), Int32, (CuPtr{Float32}, Int64, Int32, CuPtr{Int32}, CuPtr{Int32}, Int64,
        Int32, CuPtr{Float32}, Ptr{Cvoid}), pos, size(pos, 2), size(pos, 1), s, r, length(s), 1, out, CUDA.stream().handle))
    out
end
# mgn_one_hot_device / mgn_parse_edges_device / mgn_shift_one_based_device bind the same way (see INTEGRATION.md).

# ---- build_graph / inverse_data fused into the model (mgn_forward_fused, mgn_backward_fused) --------------------------
struct FeatureSeg   # mirrors mgn_feature_seg
    x::CuPtr{Float32}; ld::Int32; col::Int32; width::Int32; kind::Int32
    scale::Float32; shift::Float32; state::CuPtr{Float32}; std_eps::Float32
end
const NOSEG = FeatureSeg(CuPtr{Float32}(0), 0, 0, 0, 0, 0f0, 0f0, CuPtr{Float32}(0), 0f0)
struct FusedIo      # mirrors mgn_fused_io
    n_node::Int32; node::NTuple{8,FeatureSeg}; n_edge::Int32; edge::NTuple{8,FeatureSeg}
    n_out::Int32; out::NTuple{8,FeatureSeg}; val_mask::CuPtr{Float32}
end
pad8(v) = ntuple(i -> i <= length(v) ? v[i] : NOSEG, 8)
# online(norm) block of a matrix `x` (features x entities, i.e. ld = size(x, 1)), rows col+1 : col+width of it
online_seg(x, col, width, state::CuVector{Float32}; eps = 1f-8) =
    FeatureSeg(pointer(x), size(x, 1), col, width, 1, 0f0, 0f0, pointer(state), eps)
affine_seg(x, col, width, scale, shift) = FeatureSeg(pointer(x), size(x, 1), col, width, 0, scale, shift, CuPtr{Float32}(0), 0f0)
# ode_step (src/solve.jl:188-219) as ONE call: io describes vcat(n_norm[f](x[f])..., n_norm["node_type"](onehot)),
# e_norm(edge_features), inverse_data(o_norm[tf], .) per target field and val_mask.
function forward_fused(m::B200Model, gh::Ptr{Cvoid}, ps, io::FusedIo, n_nodes; ws = nothing)
    training = ws !== nothing
    w = training ? ws : infer_workspace(m, gh)
    out = CUDA.zeros(Float32, m.cfg.out_dim, n_nodes)
    check(ccall((:mgn_forward_fused, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, Ref{FusedIo}, CuPtr{Float32},
        CuPtr{UInt8}, Csize_t, Int32, Ptr{Cvoid}), m.handle, gh, flat(ps), io, out, w, length(w), training, CUDA.stream().handle))
    out
end
function backward_fused(m::B200Model, gh::Ptr{Cvoid}, ps, io::FusedIo, dout, ws, n_nodes)
    p = flat(ps); dps = similar(p); dx = CUDA.zeros(Float32, m.cfg.node_in, n_nodes)
    check(ccall((:mgn_backward_fused, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Float32}, Ref{FusedIo}, CuPtr{Float32},
        CuPtr{Float32}, CuPtr{Float32}, CuPtr{UInt8}, Csize_t, Ptr{Cvoid}), m.handle, gh, p, io, dout, dps, dx, ws, length(ws),
        CUDA.stream().handle))
    dps, dx
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
        rhs_forward(x, idx; training = true)          # into a checked-out training workspace the closure keeps
        gi, dx = rhs_backward(ode_lincomb!(similar(lam), CUDA.zeros(Float32, size(lam)), [lam], [dt]))
        ode_lincomb!(g, g, [gi], [1.0f0])
        ode_lincomb!(lam, lam, [dx], [1.0f0])
        n > 1 && ode_lincomb!(lam, lam, [dsaves[n]], [1.0f0])
    end
    (g,), loss
end

end # module
