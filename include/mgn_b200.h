/*
 * mgn_b200.h - C ABI of libmgn_b200.so: the B200-native (sm_100a) Encode-Process-Decode hot
 * path of una-auxme/MeshGraphNets.jl, i.e. the GraphNetCore.jl operations MeshGraphNets.jl calls.
 *
 * Drop-in boundary.  MeshGraphNets.jl (src/graph.jl, src/solve.jl, src/strategies.jl,
 * src/MeshGraphNets.jl) is unchanged; a GraphNetCore-compatible Julia module `ccall`s these
 * entry points (see INTEGRATION.md and julia/GraphNetCoreB200.jl).  Each entry point cites the
 * reference interface it replaces as  file:line  relative to the MeshGraphNets.jl v0.4.1 tree.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  Every function returns an int32_t status
 *     (MGN_OK == 0); the message for the last failure on the calling thread is read with
 *     mgn_last_error().  Nothing throws or aborts across the ABI.
 *   - Julia matrices are column-major (features, entities); the same bytes are a row-major
 *     [entities][features] C array.  That is the layout of every tensor below.
 *   - Pointers prefixed d_ are DEVICE pointers (CuPtr{T}); h_ are HOST pointers.  The caller
 *     owns every tensor buffer; the library owns only the opaque handles it creates.
 *   - `stream` is a cudaStream_t passed as void* (Julia: CUDA.stream().handle).  forward /
 *     backward / loss / adam / normaliser calls only enqueue work on that stream: no hidden
 *     synchronisation, no allocation - they are CUDA-graph capturable.
 *   - Node / edge ids follow `index_base` (1 for Julia arrays, src/graph.jl:31-34).
 *   - There is NO CPU fallback: without a CUDA device every device entry point fails with
 *     MGN_ERR_CUDA.
 */
#ifndef MGN_B200_H
#define MGN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGN_ABI_VERSION 2

enum {
  MGN_OK = 0,
  MGN_ERR_INVALID = 1, /* bad argument (null pointer, negative size, unknown enum) */
  MGN_ERR_CUDA = 2,    /* CUDA runtime error, message holds cudaGetErrorString */
  MGN_ERR_INDEX = 3,   /* node id outside [index_base, index_base + N) */
  MGN_ERR_WORKSPACE = 4, /* workspace too small / missing forward state */
  MGN_ERR_UNSUPPORTED = 5
};

/* Arithmetic of the MLP GEMMs (the latents' master copy is always fp32). */
enum {
  MGN_COMPUTE_FP32 = 0, /* fp32 FMA on CUDA cores: the parity mode                      */
  MGN_COMPUTE_BF16 = 1  /* bf16 operands, fp32 accumulate on tcgen05 tensor cores (TMEM) */
};

/* GraphNetCore.build_model(quantities, dims, outputs, mps, layer_size, hidden_layers, device),
 * reached through GraphNetCore.load at src/MeshGraphNets.jl:282-285. */
typedef struct mgn_model_config {
  int32_t node_in;       /* quantities                (src/MeshGraphNets.jl:274)        */
  int32_t edge_in;       /* dims + 1                  (src/graph.jl:51-52)              */
  int32_t out_dim;       /* outputs                   (src/MeshGraphNets.jl:277-280)    */
  int32_t latent;        /* layer_size, must be 128 for MGN_COMPUTE_BF16 (Args :37)     */
  int32_t mps;           /* message passing steps     (Args :36)                        */
  int32_t hidden_layers; /* Dense layers per MLP = hidden_layers + 2 (Args :38)         */
  float ln_eps;          /* Lux.LayerNorm epsilon, 1e-5                                 */
  int32_t compute_mode;  /* MGN_COMPUTE_*                                               */
  /* GraphNetCore.jl is not vendored in the reference tree, so three of its internals are recalled, not read
   * (SURVEY.md section 9).  Each is a field here: if oracle/julia/dump_reference.jl disagrees, flip the field - no
   * kernel changes.  All-zero = the recalled defaults. */
  int32_t dense_layers;  /* Dense layers per MLP; 0 = hidden_layers + 2 (recalled build_mlp), else explicit
                          * (hidden_layers + 1 would be DeepMind's _make_mlp).  2..4 in MGN_COMPUTE_BF16.          */
  int32_t ln_scale_first; /* LayerNorm parameter order in the flat vector: 0 = (bias, scale) (recalled Lux 0.5
                          * ComponentArray order), 1 = (scale, bias)                                              */
  int32_t aggregate_post_residual; /* 0 = scatter-sum the NEW messages m (DeepMind / recalled GraphNetCore order);
                          * 1 = scatter-sum the updated edge latent ef + m (both arithmetic modes, forward and
                          * backward: the aggregation adjoint then also flows down the edge residual path)         */
} mgn_model_config;

/* One tensor of the flat Float32 parameter vector (mirrors the ComponentArray `mgn.ps` that
 * src/strategies.jl:187 hands to ODEProblem and src/MeshGraphNets.jl:376 to Optimisers.update). */
typedef struct mgn_param_entry {
  char name[48];  /* e.g. "processor3.edge.dense2.weight" */
  int64_t offset; /* element offset into the flat vector */
  int32_t rows;   /* Julia size(.,1): out for a Dense weight */
  int32_t cols;   /* Julia size(.,2): in for a Dense weight, 1 for vectors */
} mgn_param_entry;

typedef struct mgn_model mgn_model;
typedef struct mgn_graph mgn_graph;

/* ------------------------------------------------------------------ errors / info */
int32_t mgn_abi_version(void);
/* Copies the calling thread's last error message (NUL terminated) into buf. */
int32_t mgn_last_error(char* buf, size_t n);
/* Number of visible CUDA devices (0 and MGN_OK when the driver is absent). */
int32_t mgn_device_count(int32_t* count);

/* ------------------------------------------------------------------ graph indexing (host, integer, bit exact) */
/* GraphNetCore.one_hot(v, depth, offset)  <- src/graph.jl:26-27.  h_out is [n][depth]. */
int32_t mgn_one_hot(const int32_t* h_v, int64_t n, int32_t depth, int32_t offset, float* h_out);
/* GraphNetCore.triangles_to_edges(cells)  <- src/graph.jl:30.  h_cells is [C][3]; senders /
 * receivers need room for 6*C entries; *n_edges receives E = 2 * (#unique undirected edges). */
int32_t mgn_triangles_to_edges(const int32_t* h_cells, int64_t n_cells, int32_t* h_senders,
                               int32_t* h_receivers, int64_t* n_edges);
/* GraphNetCore.parse_edges(edges)  <- src/graph.jl:38.  h_edges is [U][2]; outputs hold 2*U. */
int32_t mgn_parse_edges(const int32_t* h_edges, int64_t n_pairs, int32_t* h_senders,
                        int32_t* h_receivers);
/* The 0 -> 1 based shift of src/graph.jl:31-34 / :39-42.  *shifted is set to 1 if applied. */
int32_t mgn_shift_one_based(int32_t* h_senders, int32_t* h_receivers, int64_t n_edges,
                            int32_t* shifted);
/* Edge features [rel ; ||rel||]  <- src/graph.jl:35-36,49-52.  h_pos [N][dim], h_out [E][dim+1]. */
int32_t mgn_edge_features(const float* h_pos, int64_t n_nodes, int32_t dim, const int32_t* h_senders,
                          const int32_t* h_receivers, int64_t n_edges, int32_t index_base,
                          float* h_out);

/* ------------------------------------------------------------------ create_base_graph on the device */
/* The same five operations with inputs and outputs resident in HBM (src/graph.jl:25-55 when the trajectory already
 * lives on the device): bit-exact equal to the host functions above.  Graph construction, like mgn_graph_create, may
 * allocate scratch and synchronise `stream` (to hand E / the shift flag back to the host); it is not for capture. */
int32_t mgn_one_hot_device(const int32_t* d_v, int64_t n, int32_t depth, int32_t offset, float* d_out, void* stream);
/* d_senders / d_receivers need room for 6*C entries; *h_n_edges receives E. */
int32_t mgn_triangles_to_edges_device(const int32_t* d_cells, int64_t n_cells, int32_t* d_senders,
                                      int32_t* d_receivers, int64_t* h_n_edges, void* stream);
int32_t mgn_parse_edges_device(const int32_t* d_edges, int64_t n_pairs, int32_t* d_senders, int32_t* d_receivers,
                               void* stream);
int32_t mgn_shift_one_based_device(int32_t* d_senders, int32_t* d_receivers, int64_t n_edges, int32_t* h_shifted,
                                   void* stream);
int32_t mgn_edge_features_device(const float* d_pos, int64_t n_nodes, int32_t dim, const int32_t* d_senders,
                                 const int32_t* d_receivers, int64_t n_edges, int32_t index_base, float* d_out,
                                 void* stream);

/* ------------------------------------------------------------------ graph handle (device CSR; NEW) */
/* Builds, on the device, the stable sort of edge ids by receiver (CSR) and by sender (CSC) that
 * turns GraphNetCore's NNlib.scatter(+) into an atomics-free segmented sum.  d_senders /
 * d_receivers are the FeatureGraph index vectors of src/graph.jl:87-96 (Int32, device).
 * Synchronises `stream` once (validation of the ids).  Ids are copied; the caller may free them. */
int32_t mgn_graph_create(int64_t n_nodes, int64_t n_edges, const int32_t* d_senders,
                         const int32_t* d_receivers, int32_t index_base, void* stream,
                         mgn_graph** out);
int32_t mgn_graph_destroy(mgn_graph* g);
int32_t mgn_graph_sizes(const mgn_graph* g, int64_t* n_nodes, int64_t* n_edges);
/* Copies the index structures to the host for the bit-exact check against the oracle.  Any
 * pointer may be NULL.  row_ptr/col_ptr: [N+1]; perm: CSR slot -> 0-based original edge id [E];
 * perm_sender: CSC slot -> 0-based original edge id [E]. */
int32_t mgn_graph_get_index(const mgn_graph* g, int32_t* h_row_ptr, int32_t* h_perm,
                            int32_t* h_col_ptr, int32_t* h_perm_sender);

/* ------------------------------------------------------------------ model */
int32_t mgn_model_create(const mgn_model_config* cfg, mgn_model** out);
int32_t mgn_model_destroy(mgn_model* m);
int32_t mgn_model_param_count(const mgn_model* m, int64_t* count);
/* Fills up to `capacity` entries (may be 0 / NULL to query) and returns the total in *n. */
int32_t mgn_model_param_layout(const mgn_model* m, mgn_param_entry* entries, int32_t capacity,
                               int32_t* n);
/* Bytes of caller-owned device scratch that forward (+backward when training != 0) needs. */
int32_t mgn_workspace_bytes(const mgn_model* m, const mgn_graph* g, int32_t training,
                            size_t* bytes);

/* `mgn.model(graph, ps, st)`  <- src/solve.jl:200 and inside step! (src/strategies.jl:421).
 * d_nf [N][node_in] and d_ef [E][edge_in] are FeatureGraph.node_features / edge_features in
 * the ORIGINAL edge order; d_out is [N][out_dim].  With training != 0 the activations the
 * backward pass needs are kept in the workspace. */
int32_t mgn_forward(const mgn_model* m, const mgn_graph* g, const float* d_params,
                    const float* d_nf, const float* d_ef, float* d_out, void* d_workspace,
                    size_t workspace_bytes, int32_t training, void* stream);

/* Reverse mode of mgn_forward: what Zygote derives for step! (src/strategies.jl:421) and what
 * SciMLSensitivity's ZygoteVJP asks of ode_step (src/strategies.jl:183-194).  Needs the
 * workspace of the matching training forward.  d_dparams [P] is overwritten; d_dnf
 * [N][node_in] (gradient w.r.t. the node features, for the NeuralODE adjoint) may be NULL. */
int32_t mgn_backward(const mgn_model* m, const mgn_graph* g, const float* d_params,
                     const float* d_nf, const float* d_ef, const float* d_dout,
                     float* d_dparams, float* d_dnf, void* d_workspace, size_t workspace_bytes,
                     void* stream);

/* ------------------------------------------------------------------ graph-partitioned meshes (halo exchange) */
/* A mesh too large for one GPU is split by nodes: a rank owns a set of nodes and every edge whose receiver it owns
 * (so the scatter-sum stays local and deterministic); sender nodes owned elsewhere are appended to the local node
 * list as HALO rows.  The model then runs stage by stage, and between stages the caller exchanges halo rows with
 * the owners (NCCL send/recv, all-to-all, or peer copies - the transport is the caller's):
 *   forward : ENCODE, [unpack latent(0)], step 0, pack owned rows of latent(1) -> peers unpack into their halo rows,
 *             step 1, ..., DECODE          (the reference has no counterpart: it is single device, MeshGraphNets.jl:257)
 *   backward: DECODE, step mps-1, pack+zero the halo rows of the latent gradient -> owners add them, step mps-2, ...,
 *             ENCODE; the parameter gradient is then summed over ranks.
 * mgn_forward / mgn_backward are exactly the stage sequences without exchanges. */
enum { MGN_STAGE_ENCODE = -1, MGN_STAGE_DECODE = -2 }; /* stage >= 0: message-passing step k */
int32_t mgn_forward_stage(const mgn_model* m, const mgn_graph* g, const float* d_params, const float* d_nf,
                          const float* d_ef, float* d_out, void* d_workspace, size_t workspace_bytes,
                          int32_t training, int32_t stage, void* stream);
/* d_dparams is WRITTEN piecewise: every stage overwrites the gradient of the MLPs it covers. */
int32_t mgn_backward_stage(const mgn_model* m, const mgn_graph* g, const float* d_params, const float* d_nf,
                           const float* d_ef, const float* d_dout, float* d_dparams, float* d_dnf,
                           void* d_workspace, size_t workspace_bytes, int32_t stage, void* stream);
enum { MGN_HALO_LATENT = 0, /* node latent READ by message-passing step `step` (bf16 rows in bf16 mode, fp32 in fp32 mode) */
       MGN_HALO_GRAD = 1 }; /* gradient of the node latent (fp32 rows) */
enum { MGN_ROWS_PACK = 0,      /* buf[i]  = tensor[rows[i]]            */
       MGN_ROWS_UNPACK = 1,    /* tensor[rows[i]]  = buf[i]            */
       MGN_ROWS_ADD = 2,       /* tensor[rows[i]] += buf[i]  (fp32; rows must be distinct) */
       MGN_ROWS_PACK_ZERO = 3  /* pack, then zero the rows             */ };
/* Bytes of one row of the tensor (latent * 2 or latent * 4). */
int32_t mgn_halo_row_bytes(const mgn_model* m, int32_t what, size_t* bytes);
/* d_rows: n_rows local node ids (0-based, device); d_buf: n_rows contiguous rows (device). */
int32_t mgn_halo_rows(const mgn_model* m, const mgn_graph* g, void* d_workspace, size_t workspace_bytes,
                      int32_t training, int32_t what, int32_t step, const int32_t* d_rows, int64_t n_rows,
                      void* d_buf, int32_t op, void* stream);

/* Loss of GraphNetCore.step!(mgn, graph, target, mask, mse_reduce)  <- src/strategies.jl:421:
 * loss = mean(sum_rows((target - out)^2)[mask]); d_mask holds n_mask node ids (the Int32 vector
 * of src/MeshGraphNets.jl:352).  Writes the scalar to d_loss[0] and dloss/dout to d_dout. */
int32_t mgn_loss_mse_masked(const float* d_out, const float* d_target, int64_t n_nodes,
                            int32_t out_dim, const int32_t* d_mask, int64_t n_mask,
                            int32_t index_base, float* d_loss, float* d_dout, void* stream);

/* Optimisers.update with Adam  <- src/MeshGraphNets.jl:374-378.  t is the 1-based step. */
int32_t mgn_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n,
                      float lr, float beta1, float beta2, float eps, int64_t t, void* stream);

/* Same update with the step counter on the device, so that a whole training step can be replayed
 * as a CUDA graph: d_state16 is 16 bytes {int64 step; float c1; float c2}, zero-initialised by the
 * caller; each call increments step and uses t = step. */
int32_t mgn_adam_step_device(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n,
                             float lr, float beta1, float beta2, float eps, void* d_state16,
                             void* stream);

/* ------------------------------------------------------------------ multi-GPU transport (NEW: SURVEY 8b mgn_dp_*, 8e) */
/* The reference is single device (src/MeshGraphNets.jl:257).  One process (or Julia task) per GPU; NCCL over NVLink is
 * loaded at run time (dlopen "libnccl.so.2": no link-time dependency, MGN_ERR_UNSUPPORTED when absent).  Bootstrap:
 * rank 0 calls mgn_dp_unique_id and hands the 128 bytes to the other ranks by any means (MPI.jl, a file, a socket, or
 * torch.distributed in the Python mirror); every rank then calls mgn_dp_init on ITS device.  The collectives below only
 * enqueue on `stream` and are CUDA-graph capturable. */
#define MGN_DP_UNIQUE_ID_BYTES 128
typedef struct mgn_comm mgn_comm;
enum { MGN_DP_SUM = 0, MGN_DP_MEAN = 1 };
int32_t mgn_dp_unique_id(void* h_id128);
int32_t mgn_dp_init(const void* h_id128, int32_t rank, int32_t world, mgn_comm** out);
int32_t mgn_dp_finalize(mgn_comm* c);
int32_t mgn_dp_rank(const mgn_comm* c, int32_t* rank, int32_t* world);
/* In-place all-reduce of a flat Float32 buffer: the data-parallel gradient of src/MeshGraphNets.jl:364-378 run as
 * batch-P SGD (MEAN), or loss / gradient of interval-sharded MultipleShooting and of a partitioned mesh (SUM). */
int32_t mgn_dp_allreduce(mgn_comm* c, float* d_buf, int64_t n, int32_t op, void* stream);
/* Online-normaliser state after a step in which every rank accumulated its own window on top of the common d_prev:
 * state = prev + sum over ranks of (state_r - prev); n_floats = 2 * features + 2.  d_prev may alias nothing. */
int32_t mgn_dp_allreduce_normaliser(mgn_comm* c, float* d_state, const float* d_prev, int32_t n_floats, void* stream);
/* Halo exchange of a partitioned mesh (one per message-passing step): rank p receives
 * h_send_rows[p] rows of row_bytes bytes from d_send (packed per peer in rank order) and this rank receives
 * h_recv_rows[p] rows from rank p into d_recv (same packing): grouped ncclSend / ncclRecv. */
int32_t mgn_halo_exchange(mgn_comm* c, const void* d_send, const int64_t* h_send_rows, void* d_recv,
                          const int64_t* h_recv_rows, int64_t row_bytes, void* stream);

/* Optimisers.update fused with the gradient all-reduce, bucketed and overlapped with the backward pass
 * (src/MeshGraphNets.jl:374-378).  mgn_backward_dp is mgn_backward plus, per bucket of consecutive MLPs
 * (gradients complete in reverse parameter order: decoder first, encoders last): as soon as a bucket's gradients are
 * final, on an internal side stream, (1) ncclAllReduce(avg) over `comm` when comm != NULL, (2) the Adam update of that
 * bucket's parameters when adam != NULL (its own parameters are no longer read by the rest of the backward pass).
 * `stream` joins the side stream before the call returns control to later work on `stream`.  With comm == NULL and
 * adam == NULL it is mgn_backward.  d_params is updated in place when adam != NULL. */
typedef struct mgn_adam_config {
  float lr, beta1, beta2, eps;
  float* d_m;       /* [P] first moments  */
  float* d_v;       /* [P] second moments */
  void* d_state16;  /* {int64 step; float c1; float c2}, see mgn_adam_step_device */
} mgn_adam_config;
int32_t mgn_backward_dp(const mgn_model* m, const mgn_graph* g, float* d_params, const float* d_nf, const float* d_ef,
                        const float* d_dout, float* d_dparams, float* d_dnf, void* d_workspace, size_t workspace_bytes,
                        mgn_comm* comm, const mgn_adam_config* adam, int32_t n_buckets, void* stream);
/* Frees the library's per-(device, stream) scratch buffers; handles stay valid. */
int32_t mgn_library_release(void);

/* ------------------------------------------------------------------ measurement hooks (bench.py) */
/* Counts every kernel launch of the library and brackets the launches of one kernel family
 * (`tag`, see mgn_profile_tag_name; -1 = count only) with CUDA events on their own stream.
 * mgn_profile_end synchronises the device and returns the launch count, the number of tagged
 * launches and the sum of their device durations.  Not for use under CUDA-graph capture. */
int32_t mgn_profile_begin(int32_t tag);
int32_t mgn_profile_end(int64_t* n_launches, int64_t* n_tagged, float* tagged_ms, int64_t* per_tag,
                        int32_t per_tag_capacity);
int32_t mgn_profile_tag_name(int32_t tag, char* buf, size_t n);

/* ------------------------------------------------------------------ normalisers */
/* NormaliserOnline state on the device: d_state = [acc_sum[F] | acc_sum_sq[F] | acc_count |
 * num_acc] (2F+2 floats).  mgn_norm_online_update is the accumulate branch of the callable
 * (src/graph.jl:80,84,93; src/strategies.jl:399-410); it is skipped on-device once
 * num_acc >= max_acc.  d_x is [M][F], F <= 64.  Deterministic two-stage sum; the first call on a device allocates a
 * 64 KB scratch private to (device, stream) - also legal inside a stream capture; a captured graph keeps the scratch of its
 * capture stream, so replay one instance of it at a time. */
int32_t mgn_norm_online_update(const float* d_x, int64_t rows, int32_t features, float* d_state,
                               float max_acc, void* stream);
/* y = (x - mean)/std (inverse == 0) or y = x*std + mean (inverse != 0, inverse_data at
 * src/solve.jl:207-209) from an online state.  y has leading dimension ld_y and starts at
 * column col_y, so build_graph's vcat (src/graph.jl:80-86) is written in place. */
int32_t mgn_norm_online_apply(const float* d_x, int64_t rows, int32_t features,
                              const float* d_state, float std_eps, int32_t inverse, float* d_y,
                              int32_t ld_y, int32_t col_y, void* stream);
/* Offline normalisers and any other per-feature affine map: y = x*scale + shift. */
int32_t mgn_affine_apply(const float* d_x, int64_t rows, int32_t features, float scale,
                         float shift, float* d_y, int32_t ld_y, int32_t col_y, void* stream);

/* ------------------------------------------------------------------ build_graph / inverse_data fused into the model */
/* `build_graph` (src/graph.jl:75-97) is `nf = vcat(n_norm[f](data[f][:, :, t]) for f in fields..., n_norm["node_type"](
 * onehot))`, `ef = e_norm(edge_features)`; ode_step (src/solve.jl:205-218) post-processes the output with
 * `inverse_data(o_norm[tf], out[cols]) .* val_mask`.  Instead of materialising those matrices the caller DESCRIBES them:
 * a list of column blocks, each the image of a raw matrix under an offline (affine) or an online normaliser.  In the
 * tensor-core mode the first Dense layer of the encoders evaluates the recipe while staging its operand, the decoder's
 * last epilogue applies the inverse map and the mask, and the backward kernels apply the transposed maps where they
 * read the features / the cotangent and where they write d_dx; the fp32 mode materialises each recipe with one launch.
 * Statistics are READ (never accumulated) here: run mgn_norm_online_update_multi first when the step accumulates. */
enum { MGN_FEAT_AFFINE = 0,   /* y = x * scale + shift   (NormaliserOfflineMinMax / MeanStd; {1, 0} = copy)      */
       MGN_FEAT_ONLINE = 1 }; /* y = (x - mean) / std from d_state (node / edge blocks); y * std + mean (outputs) */
typedef struct mgn_feature_seg {
  const float* d_x;     /* [rows][ld] source matrix (unused for output blocks)                      */
  int32_t ld, col, width; /* the block is columns [col, col + width) of d_x                          */
  int32_t kind;         /* MGN_FEAT_*                                                                */
  float scale, shift;   /* MGN_FEAT_AFFINE (for output blocks: the INVERSE map's scale and shift)    */
  const float* d_state; /* MGN_FEAT_ONLINE: [sum | sum_sq | count | num_acc], 2 * width + 2 floats   */
  float std_eps;
} mgn_feature_seg;
typedef struct mgn_fused_io {
  int32_t n_node_segs;  /* blocks of the node-feature matrix in vcat order (src/graph.jl:80-86); widths sum to node_in */
  mgn_feature_seg node[8];
  int32_t n_edge_segs;  /* blocks of the edge-feature matrix (src/graph.jl:93), rows in ORIGINAL edge order */
  mgn_feature_seg edge[8];
  int32_t n_out_segs;   /* inverse_data per target field (src/solve.jl:205-210); 0 = return the network output */
  mgn_feature_seg out[8];
  const float* d_val_mask; /* [N][out_dim] (src/solve.jl:218) or NULL */
} mgn_fused_io;
/* mgn_forward with the recipes in place of d_nf / d_ef; d_out = val_mask .* inverse_data(model(...)). */
int32_t mgn_forward_fused(const mgn_model* m, const mgn_graph* g, const float* d_params, const mgn_fused_io* io,
                          float* d_out, void* d_workspace, size_t workspace_bytes, int32_t training, void* stream);
/* Pullback of mgn_forward_fused (what ZygoteVJP derives for ode_step, src/strategies.jl:183-194): d_dout is the
 * cotangent of the FUSED output; d_dx [N][node_in] (nullable) receives the gradient w.r.t. the RAW source columns of the
 * node blocks, in vcat order (the transposed normalisers are applied). */
int32_t mgn_backward_fused(const mgn_model* m, const mgn_graph* g, const float* d_params, const mgn_fused_io* io,
                           const float* d_dout, float* d_dparams, float* d_dx, void* d_workspace,
                           size_t workspace_bytes, void* stream);
/* The accumulate branch of up to 8 online normalisers (every n_norm / e_norm / o_norm call of one step: src/graph.jl:
 * 80,84,93 and src/strategies.jl:399-410) in TWO launches: x is columns [col, col + features) of a [rows][ld] matrix. */
typedef struct mgn_norm_update {
  const float* d_x;
  int64_t rows;
  int32_t ld, col, features;
  float* d_state;
  float max_acc;
} mgn_norm_update;
int32_t mgn_norm_online_update_multi(const mgn_norm_update* h_jobs, int32_t n_jobs, void* stream);

/* ------------------------------------------------------------------ NeuralODE callers (src/solve.jl, SolverStrategy of src/strategies.jl) */
/* The state of the NeuralODE is x [N][S] (S = sum of the target feature dims; src/strategies.jl:171-172).  A solver
 * strategy evaluates ode_func_train (src/solve.jl:101-115) many times per training step; the pieces of that right-hand
 * side that are not the model itself, and the pieces of its pullback (what SciMLSensitivity's ZygoteVJP derives,
 * src/strategies.jl:183-194), are the entry points below.  All of them only enqueue work on `stream`. */

/* One explicit Runge-Kutta combination: y = x + sum_j coef[j] * k[j], j < n_terms <= 8 (stage inputs x + h*sum a_ij k_j,
 * the update x + h*sum b_i k_i, and - transposed - the adjoint accumulations).  h_k is a HOST array of n_terms DEVICE
 * pointers, h_coef a HOST array; d_x may be NULL (treated as zero); d_y may alias d_x or any k[j].  Every product and
 * sum is rounded separately in ascending j (what the broadcast `x .+ c1 .* k1 .+ ...` does). */
int32_t mgn_ode_lincomb(const float* d_x, const float* const* h_k, const float* h_coef, int32_t n_terms, int64_t n,
                        float* d_y, void* stream);
/* Inflow overwrite `bx[inflow_mask] = data[...][inflow_mask]`  <- src/solve.jl:104-107 (train), :151 (eval):
 * y[i] = mask[i] ? src[i] : x[i].  With d_src == NULL it is the transposed Jacobian: y[i] = mask[i] ? 0 : x[i]. */
int32_t mgn_masked_overwrite(const float* d_x, const float* d_src, const uint8_t* d_mask, int64_t n, float* d_y,
                             void* stream);
/* `copy(buf) .* val_mask`  <- src/solve.jl:218 (and its pullback): y = a .* b. */
int32_t mgn_vec_mul(const float* d_a, const float* d_b, int64_t n, float* d_y, void* stream);

/* Strided normaliser maps without accumulation: x is read from columns [col_x, col_x + features) of a matrix with
 * leading dimension ld_x, y written likewise.  FORWARD / INVERSE are mgn_norm_online_apply's two directions; the _VJP
 * modes are their transposed Jacobians (dy / std and dy * std), used by the pullback of build_graph (src/graph.jl:80-86)
 * and of inverse_data (src/solve.jl:205-210). */
enum { MGN_NORM_FORWARD = 0, MGN_NORM_INVERSE = 1, MGN_NORM_FORWARD_VJP = 2, MGN_NORM_INVERSE_VJP = 3 };
int32_t mgn_norm_online_apply_ld(const float* d_x, int32_t ld_x, int32_t col_x, int64_t rows, int32_t features,
                                 const float* d_state, float std_eps, int32_t mode, float* d_y, int32_t ld_y,
                                 int32_t col_y, void* stream);
int32_t mgn_affine_apply_ld(const float* d_x, int32_t ld_x, int32_t col_x, int64_t rows, int32_t features,
                            float scale, float shift, float* d_y, int32_t ld_y, int32_t col_y, void* stream);

/* train_loss(::SolverTraining)  <- src/strategies.jl:253-286, and the per-interval error term of
 * train_loss(::MultipleShooting)  <- src/strategies.jl:367-378:  with e = (gt - pred).^2 .* val_mask over n_saves saved
 * states of mask_elems = N*S values each,  d_loss[0] = (accumulate ? d_loss[0] : 0) + weight * sum(e)  (weight =
 * 1 / (n_saves * mask_elems) gives `mean`), and d_dpred = d loss / d pred.  Fixed summation order (deterministic).
 * Scratch (1 KB) is private to (device, stream), allocated on first use - also legal inside a stream capture. */
int32_t mgn_shooting_mse(const float* d_pred, const float* d_gt, const float* d_val_mask, int64_t n_saves,
                         int64_t mask_elems, float weight, int32_t accumulate, float* d_loss, float* d_dpred,
                         void* stream);
/* Continuity term `continuity_term * sum(abs, pred_prev[:, :, end] - gt[:, :, first(rg)])`  <- src/strategies.jl:380-383:
 * d_loss[0] += weight * sum|a - b| and d_da += weight * sign(a - b)  (d_da is ACCUMULATED into). */
int32_t mgn_shooting_continuity(const float* d_a, const float* d_b, int64_t n, float weight, float* d_loss,
                                float* d_da, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGN_B200_H */
