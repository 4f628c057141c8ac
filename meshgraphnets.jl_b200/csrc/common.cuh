// Internal declarations shared by the translation units of libmgn_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mgn_b200.h"

namespace mgn {

void set_error(const std::string& msg);
int32_t fail(int32_t code, const std::string& msg);

#define MGN_CUDA_TRY(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::mgn::fail(MGN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
  } while (0)

#define MGN_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) return ::mgn::fail(MGN_ERR_INVALID, (msg));    \
  } while (0)

#define MGN_TRY(expr)                  \
  do {                                 \
    int32_t _s = (expr);               \
    if (_s != MGN_OK) return _s;       \
  } while (0)

// ---- per-device library state (device_state.cu) -----------------------------------------------------
// The library keeps NO process-global mutable state on the launch path: what has to be set up once per
// CUDA device (opt-in shared-memory limits of the big kernels, the SM count) is guarded by one
// std::call_once per device, so a process that switches devices (CUDA.device!, src/MeshGraphNets.jl:257)
// or calls from several host threads launches correctly configured kernels on every device.
constexpr int kMaxDevices = 64;
class PerDeviceOnce {
 public:
  // Runs f(device) the first time it is called on the current device; every call returns f's status.
  template <class F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    std::call_once(flag_[dev], [&] { err_[dev] = f(dev); });
    return err_[dev];
  }

 private:
  std::once_flag flag_[kMaxDevices];
  cudaError_t err_[kMaxDevices] = {};
};
int device_sm_count();  // SMs of the current device (cached per device, thread-safe)
// Scratch private to (current device, stream, kind): two calls on different streams or host threads never share
// partial sums.  Allocated on first use under a mutex; if that first use happens inside a stream capture the
// allocation is made with the thread's capture mode relaxed, so the capture stays valid.  A captured graph keeps
// the scratch of its capture stream: replay one graph instance at a time.
enum ScratchKind : int { SCRATCH_LOSS = 0, SCRATCH_NORM = 1, SCRATCH_NORM_MULTI = 2, SCRATCH_KINDS };
cudaError_t stream_scratch(cudaStream_t st, int kind, size_t bytes, void** out);
void release_device_state();  // frees every scratch buffer (mgn_library_release)
// A second stream per device for work that may run BESIDE the main pass: the fixed-order reduction of the weight-gradient
// partials of MLP i overlaps the backward kernels of MLP i-1 (tc_pipeline.cu).  fork / done are indexed by the partial
// buffer set (double buffered); every user joins the lane back into its stream before returning (capturable fork/join).
struct ReduceLane {
  cudaStream_t side = nullptr;
  cudaEvent_t fork[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, join = nullptr;
};
ReduceLane* reduce_lane();    // of the current device; nullptr if it cannot be created

// Tuning knobs, read from the environment ONCE per model handle (mgn_model_create), never on the launch path.
struct TuneKnobs {
  int fwd_epi_warps = 8;   // MGN_FWD_EPI_WARPS=4 selects the one-thread-per-row epilogue
  int fwd_stagger_ns = 0;  // MGN_FWD_STAGGER_NS
  int fwd_deep_ring = 1;   // MGN_FWD_DEEP_RING=0 disables the deep-ring variant for small graphs
  int fwd_persist = 1;     // inference pass of a graph with no more tiles than SMs as ONE persistent cooperative launch:
                           // MGN_FWD_PERSIST=0 never, 1 unless the call is being captured into a CUDA graph, 2 always
  int recompute = 0;       // MGN_RECOMPUTE=1: no activation saves of the processor MLPs in the forward pass; the backward pass
                           // re-runs each MLP for them (5x less workspace per edge row and MP step, ~1.3x the step time)
  int reduce_lane = 1;     // MGN_REDUCE_LANE=0: reduce the weight-gradient partials inline on the caller's stream
  int pdl = 0;             // MGN_PDL=1: programmatic dependent launch between the library's kernels (measured: no gain
                           // inside a CUDA graph, -3 % on the 32-window step; kept as an opt-in for eager callers)
};
TuneKnobs read_tune_knobs();

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// Consecutive kernels of a forward / backward pass are launched with programmatic stream serialization: a kernel's CTAs
// may become resident (and run their prologue: barrier init, TMEM allocation) while the previous kernel drains; every
// kernel executes pdl_wait() BEFORE its first global-memory access (read or write), which returns only when the whole
// previous grid has completed and its writes are visible - so the data dependencies are exactly those of plain
// launches.  pdl_trigger() at the top of a kernel lets its successor be scheduled as soon as SM resources free up.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
cudaError_t launch_kernel(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#endif

constexpr int kMaxDense = 8;
constexpr int kOdeMaxTerms = 8;  // terms of one explicit Runge-Kutta combination (mgn_ode_lincomb)

// ---- launch counter / per-kernel-family device timer (mgn_profile_begin / mgn_profile_end) ------
enum KernelTag : int {
  TAG_SIMT_GEMM_FWD = 0, TAG_SIMT_GEMM_DX, TAG_SIMT_DW, TAG_REDUCE_PARTIALS, TAG_SEGMENT_SUM,
  TAG_LN_BWD, TAG_LN_REDUCE, TAG_NODE_GRAD_GATHER, TAG_ADD_COLS, TAG_LOSS, TAG_ADAM, TAG_NORM,
  TAG_TC_PACK, TAG_TC_MLP_FWD, TAG_TC_MLP_BWD, TAG_TC_DW, TAG_TC_MISC, TAG_SOLVER, TAG_COUNT
};
const char* tag_name(int tag);
struct ProfScope {  // records a CUDA event pair around one launch when its family is being timed
  cudaStream_t st;
  int slot;
  ProfScope(int tag, cudaStream_t st);
  ~ProfScope();
};

// One MLP of the flat parameter vector: n_dense Dense layers (+ LayerNorm).
struct MlpLayout {
  std::string name;
  int in_dim = 0, out_dim = 0, n_dense = 0;
  bool layer_norm = false;
  int64_t w_off[kMaxDense], b_off[kMaxDense];
  int in[kMaxDense], out[kMaxDense];
  int64_t ln_bias_off = -1, ln_scale_off = -1;
};

}  // namespace mgn

namespace mgn { namespace tc { struct ModelImages; } }

struct mgn_model {
  mgn_model_config cfg;
  mgn::tc::ModelImages* images = nullptr;  // packed-weight image plan (MGN_COMPUTE_BF16)
  mgn::TuneKnobs knobs;                    // environment knobs, frozen at creation
  std::vector<mgn::MlpLayout> mlps;  // encoder.node, encoder.edge, (edge, node) x mps, decoder
  int64_t n_params = 0;
  int n_dense() const { return cfg.dense_layers > 0 ? cfg.dense_layers : cfg.hidden_layers + 2; }
};

struct mgn_graph {
  int64_t N = 0, E = 0;
  int32_t index_base = 1;
  // device arrays (0-based)
  int32_t* row_ptr = nullptr;    // [N+1] CSR by receiver
  int32_t* perm = nullptr;       // [E]   CSR slot -> original edge id (stable)
  int32_t* send_csr = nullptr;   // [E]   sender node of CSR slot
  int32_t* recv_csr = nullptr;   // [E]   receiver node of CSR slot
  int32_t* col_ptr = nullptr;    // [N+1] CSC by sender
  int32_t* perm_sender = nullptr;  // [E] CSC slot -> original edge id (stable)
  int32_t* csc_slot = nullptr;   // [E]   CSC slot -> CSR slot of the same edge
  int32_t max_in_degree = 0;
  // node-aligned 128-row tiles of the CSR edge list (tensor-core path): tile t holds the CSR slots
  // [tile_row_start[t], tile_row_start[t+1]) = all in-edges of nodes [tile_node_start[t], ..[t+1])
  int32_t* tile_row_start = nullptr;   // [n_edge_tiles + 1]
  int32_t* tile_node_start = nullptr;  // [n_edge_tiles + 1]
  int32_t n_edge_tiles = 0;
  int32_t* csc_pos = nullptr;          // [E] CSC slot -> row of the same edge in tile-image space: tile * 128 + row in tile
  bool tiles_ok = false;               // false when a node has more than 128 in-edges
};

namespace mgn {

// A row-concatenated, optionally gathered operand: row r of the logical matrix is
// [seg0[idx0[r]] | seg1[idx1[r]] | seg2[idx2[r]]] (idx == nullptr -> identity).
struct Seg {
  const float* base;
  const int32_t* idx;
  int width;
  int ld;
};
struct Operand {
  Seg s[3];
  int nseg;
  int K() const {
    int k = 0;
    for (int i = 0; i < nseg; ++i) k += s[i].width;
    return k;
  }
};

// ---- fp32 (CUDA-core) kernels: simt_kernels.cu ------------------------------------------------
struct LnEpilogue {          // final Dense layer: bias -> LayerNorm -> (residual)
  const float* ln_scale;     // [n]
  const float* ln_bias;      // [n]
  float eps;
  float* xhat;               // [M][n] saved normalised pre-affine values (nullable)
  float* rstd;               // [M]    (nullable)
  float* out;                // [M][n] LayerNorm result (nullable)
  const float* resid_in;     // [M][n] (nullable)
  float* resid_out;          // [M][n] = resid_in + LN (nullable)
};
// Y[M][n] = act(X W + b); W is [K][n] row-major (Julia (out x in) column-major).
cudaError_t dense_forward(const Operand& x, int64_t M, const float* W, const float* bias, int n,
                          bool relu, float* Y, const LnEpilogue* ln, cudaStream_t st);
// dX[M][K] = dZ[M][n] W^T, optionally masked by (mask_src > 0) (ReLU of the previous layer).
cudaError_t dense_backward_dx(const float* dZ, int64_t M, int n, const float* W, int K,
                              const float* relu_src, float* dX, cudaStream_t st);
// g_w[K][n] = X^T dZ and g_b[n] = colsum(dZ), written to grad (bias directly follows weight).
// partial: scratch of dw_partial_floats(K, n, M) floats.
size_t dw_partial_floats(int K, int n, int64_t M);
cudaError_t dense_backward_dw(const Operand& x, const float* dZ, int64_t M, int n, float* partial,
                              float* g_w_and_b, cudaStream_t st);
// agg[v] = sum over CSR row v of m[j] (ascending slot = ascending original edge id).
cudaError_t segment_sum(const float* m, const int32_t* row_ptr, int64_t N, int D, float* agg,
                        cudaStream_t st);
// dz = LayerNorm backward of (a[r] + b[bidx[r]]) ; column sums -> g_scale, g_bias.
size_t ln_partial_floats(int64_t M, int D);
cudaError_t layernorm_backward(const float* a, int lda, const float* b, int ldb,
                               const int32_t* bidx, const float* xhat, const float* rstd,
                               const float* scale, int64_t M, int D, float* dz, float* partial,
                               float* g_scale, float* g_bias, cudaStream_t st);
// d_nf[v] = base[v] + add[v] + sum_{CSR row v} dxe[j][D:2D] + sum_{CSC row v} dxe[slot][0:D]
cudaError_t node_grad_gather(const float* base, const float* add, int ld_add, const float* dxe,
                             const int32_t* row_ptr, const int32_t* col_ptr,
                             const int32_t* csc_slot, int64_t N, int D, float* out,
                             cudaStream_t st);
// out[r][0:D] = a[r][0:D] + b[r][col_b : col_b + D]
cudaError_t add_cols(const float* a, const float* b, int ldb, int col_b, int64_t M, int D,
                     float* out, cudaStream_t st, const float* g = nullptr, int ldg = 0, int col_g = 0,
                     const int32_t* idx = nullptr);   // + g[idx[r], col_g : col_g + D] when g != nullptr
cudaError_t loss_mse_masked(const float* out, const float* target, int64_t N, int out_dim,
                            const int32_t* mask, int64_t n_mask, int base, float* loss,
                            float* dout, cudaStream_t st);
cudaError_t adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1,
                      float b2, float eps, int64_t t, cudaStream_t st);
cudaError_t adam_step_device(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                             float b1, float b2, float eps, void* state16, cudaStream_t st);
// The two halves of adam_step_device, for bucketed updates: advance the device-side step counter once, then update
// any number of parameter ranges with it.
cudaError_t adam_tick(void* state16, float b1, float b2, cudaStream_t st);
cudaError_t adam_apply_range(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
                             float eps, const void* state16, cudaStream_t st);
cudaError_t norm_online_update(const float* x, int64_t rows, int F, float* state, float max_acc,
                               cudaStream_t st);
cudaError_t norm_online_apply(const float* x, int64_t rows, int F, const float* state,
                              float std_eps, int inverse, float* y, int ld_y, int col_y,
                              cudaStream_t st);
cudaError_t affine_apply(const float* x, int64_t rows, int F, float scale, float shift, float* y,
                         int ld_y, int col_y, cudaStream_t st);

// ---- NeuralODE callers: solver_kernels.cu -----------------------------------------------------
cudaError_t ode_lincomb(const float* x, const float* const* k, const float* coef, int n_terms, int64_t n, float* y,
                        cudaStream_t st);
cudaError_t masked_overwrite(const float* x, const float* src, const uint8_t* mask, int64_t n, float* y,
                             cudaStream_t st);
cudaError_t vec_mul(const float* a, const float* b, int64_t n, float* y, cudaStream_t st);
cudaError_t norm_online_apply_ld(const float* x, int ld_x, int col_x, int64_t rows, int F, const float* state,
                                 float std_eps, int mode, float* y, int ld_y, int col_y, cudaStream_t st);
cudaError_t affine_apply_ld(const float* x, int ld_x, int col_x, int64_t rows, int F, float scale, float shift,
                            float* y, int ld_y, int col_y, cudaStream_t st);
cudaError_t shooting_mse(const float* pred, const float* gt, const float* vm, int64_t period, int64_t n, float w,
                         int accumulate, float* loss, float* dpred, cudaStream_t st);
cudaError_t shooting_continuity(const float* a, const float* b, int64_t n, float w, float* loss, float* da,
                                cudaStream_t st);

// ---- CSR build: csr.cu -----------------------------------------------------------------------
int32_t build_graph_index(mgn_graph* g, const int32_t* d_senders, const int32_t* d_receivers,
                          cudaStream_t st);

// ---- orchestration: pipeline.cu --------------------------------------------------------------
struct FusedIo;  // features.cuh: build_graph / inverse_data recipes evaluated inside the model kernels (nullptr: plain tensors)
// Notified by the backward pass, on the host, right after the launches that FINISH the gradient of MLP `mi` have been
// enqueued on the stream; MLPs finish in descending index order (decoder first), so mlp_done(mi) means that the flat
// gradient range from MLP mi to the end is final on the stream (dp.cu buckets the all-reduce / Adam update on it).
struct GradHook {
  // `where`: the stream on which the gradient of MLP mi (and of every MLP after it) is final
  virtual int32_t mlp_done(size_t mi, cudaStream_t where) = 0;
  virtual ~GradHook() = default;
};
int32_t workspace_bytes(const mgn_model* m, const mgn_graph* g, bool training, size_t* bytes);
int32_t forward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                const float* ef, float* out, void* ws, size_t ws_bytes, bool training,
                cudaStream_t st, const FusedIo* io = nullptr);
int32_t backward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                 const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                 size_t ws_bytes, cudaStream_t st, GradHook* hook = nullptr, const FusedIo* io = nullptr);
int32_t forward_stage(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                      const float* ef, float* out, void* ws, size_t ws_bytes, bool training, int stage,
                      cudaStream_t st, const FusedIo* io = nullptr);
int32_t backward_stage(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                       const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                       size_t ws_bytes, int stage, cudaStream_t st, GradHook* hook = nullptr,
                       const FusedIo* io = nullptr);
int32_t halo_rows(const mgn_model* m, const mgn_graph* g, void* ws, size_t ws_bytes, bool training, int what,
                  int step, const int32_t* rows, int64_t n_rows, void* buf, int op, cudaStream_t st);
// Row pack / unpack / add between a [N][row_elems] tensor of elem_bytes-wide elements and a contiguous buffer.
cudaError_t rows_op(void* base, int elem_bytes, int row_elems, const int32_t* rows, int64_t n_rows, int64_t N,
                    void* buf, int op, cudaStream_t st);

}  // namespace mgn
