// tcgen05 (5th-gen tensor core) path: bf16 operands, fp32 accumulation in TMEM.
#pragma once
#include "common.cuh"
#include "features.cuh"

namespace mgn {

// Model-handle hooks (abi.cu): build / free the packed-weight image plan
int32_t tc_model_init(mgn_model* m);
void tc_model_free(mgn_model* m);

// Entry points used by pipeline.cu
int32_t tc_workspace_bytes(const mgn_model* m, const mgn_graph* g, bool training, size_t* bytes);
int32_t tc_forward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                   const float* ef, float* out, void* ws, size_t ws_bytes, bool training,
                   cudaStream_t st, const FusedIo* io = nullptr);
int32_t tc_backward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                    const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                    size_t ws_bytes, cudaStream_t st, GradHook* hook = nullptr, const FusedIo* io = nullptr);
int32_t tc_backward_scratch_bytes(const mgn_model* m, const mgn_graph* g, size_t* bytes);
int32_t tc_forward_stage(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                         const float* ef, float* out, void* ws, size_t ws_bytes, bool training, int stage,
                         cudaStream_t st, const FusedIo* io = nullptr);
int32_t tc_backward_stage(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                          const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                          size_t ws_bytes, int stage, cudaStream_t st, GradHook* hook = nullptr,
                          const FusedIo* io = nullptr);
int32_t tc_halo_rows(const mgn_model* m, const mgn_graph* g, void* ws, size_t ws_bytes, bool training, int what,
                     int step, const int32_t* rows, int64_t n_rows, void* buf, int op, cudaStream_t st);

namespace tc {

constexpr int kTile = 128;           // rows per tile (UMMA M)
constexpr size_t kTileB = 16384;     // bytes of one T128 operand tile (128 rows x 64 bf16)
constexpr int kMaxLayers = 4;

// ---- packed weight images ------------------------------------------------------------------------
// One entry per 16 KB image tile.  kind 0 (forward, B operand of  H W^T... i.e. D = X * W):
//   tile element (n, k) = W[kb*64 + k][n]        (n = output feature, K-major over the input)
// kind 1 (backward, B operand of dX = dZ * W^T):
//   tile element (n, k) = W[nb*128 + n][kb*64 + k]   (n = input feature, K-major over the output)
struct PackTile {
  int64_t w_off;   // element offset of the Dense weight in the flat parameter vector
  int32_t in_dim, out_dim;
  int32_t kind, kb, nb;
  int32_t pad;
};

struct MlpImages {          // tile offsets (units of kTileB) into the image buffer
  int fwd_off[kMaxLayers];  // layer l forward image: nkb_f tiles
  int nkb_f[kMaxLayers];
  int bwd_off[kMaxLayers];  // layer l backward image: nb_b x nkb_b tiles, index nb * nkb_b + kb
  int nb_b[kMaxLayers], nkb_b[kMaxLayers];
};

struct ModelImages {
  std::vector<MlpImages> mlps;
  std::vector<PackTile> tiles;
  PackTile* d_tiles = nullptr;  // device copy (owned by the model handle)
  int n_tiles = 0;
};

cudaError_t pack_weights(const ModelImages& im, const float* params, __nv_bfloat16* images, cudaStream_t st, bool pdl);

// ---- fused MLP forward -----------------------------------------------------------------------------
enum InMode { IN_RAW = 0, IN_PLAIN = 1, IN_CONCAT2 = 2, IN_GATHER3 = 3 };
enum FinMode { FIN_LN = 0, FIN_LN_RESID = 1, FIN_LN_RESID_AGG = 2, FIN_LINEAR = 3 };

// Everything a forward launch needs except the two feature recipes (which only the encoders / the decoder read, once, in
// their prologue): the persistent kernel keeps one FwdCore per stage in kernel-parameter space.
struct FwdCore {
  // tiling
  int n_tiles;
  int64_t M;                          // rows
  const int32_t* tile_row_start;      // [n_tiles+1] (nullptr: tile t covers rows [128 t, 128 t + 128))
  const int32_t* tile_node_start;     // [n_tiles+1] FIN_LN_RESID_AGG: first node of each tile
  const int32_t* row_ptr;             // CSR row pointer (aggregation)
  // input operand
  int in_mode;
  const __nv_bfloat16 *x0, *x1, *x2;  // row-major [rows][128] bf16
  const __nv_bfloat16* x2_img;        // IN_GATHER3: the third segment as tile images [tile][2][16 KB] (bulk copies)
  const int32_t *idx0, *idx1;         // IN_GATHER3: rows of x0 for K-blocks {0,1} / {2,3}
  const int32_t* raw_idx;             // IN_RAW: optional row gather (CSR perm)
  int raw_F;                          // == FwdParams::feat.F
  // layers
  int n_layers;
  int nkb[kMaxLayers];                // K-blocks (64 wide) of each layer
  int ksteps0;                        // UMMA K-steps per K-block in layer 0 (4, or ceil(raw_F/16))
  const __nv_bfloat16* wimg[kMaxLayers];  // forward image of each layer
  const float* bias[kMaxLayers];
  int n_out_last;                     // 128, or out_dim for FIN_LINEAR
  const float *ln_scale, *ln_bias;
  float eps;
  // outputs
  int fin_mode;
  const float* lat_in;                // fp32 [rows][128] residual input (node latents: fp32 master) ...
  const __nv_bfloat16* lat_img_in;    // ... or the bf16 tile images of the latent itself (edge latents: stored in bf16 only)
  float* lat_out;                     // fp32 [rows][128] (nullptr: no fp32 master is kept)
  __nv_bfloat16* lat_bf16_out;        // bf16 shadow of lat_out (row-major) ...
  __nv_bfloat16* lat_img_out;         // ... or, when non-null, as tile images (one 32 KB bulk store per tile)
  __nv_bfloat16* agg_bf16;            // [nodes][128]
  int agg_post_residual;              // FIN_LN_RESID_AGG: aggregate the updated latent (residual added) instead of the message
  float* out;                         // FIN_LINEAR: [rows][out_dim]
  int out_dim;
  const float* val_mask;              // FIN_LINEAR: `.* val_mask` [rows][out_dim] (nullable)   <- src/solve.jl:218
  // training saves (nullptr when not training)
  __nv_bfloat16* save_h[kMaxLayers - 1];  // image [tile][2 tiles]
  __nv_bfloat16* save_xhat;               // image [tile][2 tiles]
  float* save_rstd;                       // [rows]
  unsigned long long* trace;              // debug: per-role %globaltimer stamps of CTA 0 (nullptr = off)
  uint32_t stagger_ns;                    // start delay unit that de-phases co-resident CTAs (0 = off)
  int epi_warps;                          // 8 (two threads per tile row) or 4: mgn_model::knobs
  int deep_ring;                          // allow the deep-ring variant when the graph has no more tiles than SMs
  int pdl;                                // programmatic dependent launch (common.cuh)
};
struct FwdParams : FwdCore {
  FeatRecipe feat;                    // IN_RAW: the raw fp32 features as a recipe (features.cuh): normalise + concat on the fly
  FeatRecipe out_feat;                // FIN_LINEAR: inverse_data per output column (n == 0: none)   <- src/solve.jl:205-210
};

cudaError_t mlp_forward_tc(const FwdParams& p, cudaStream_t st);

// Persistent forward pass (graphs with no more tiles than SMs): every MLP of the pass is a stage of ONE cooperative launch.
constexpr int kMaxStages = 36;   // 3 + 2 * mps: mps <= 16
struct PersistParams {
  int n_stages;
  unsigned int* sync;            // grid-barrier counter, zeroed on the stream before the launch
  unsigned long long* dbg;       // debug build only: CTA 0 stamps %globaltimer at [3 * stage + {0: top, 1: body done, 2: barrier passed}]
  FeatRecipe feat[3];            // node encoder input, edge encoder input, decoder output
  FwdCore stage[kMaxStages];     // stage 0: node encoder, 1: edge encoder, then (edge, node) per MP step, last: decoder
};
// True when the pass can run persistently on the current device (tile counts, stage count).
bool forward_persist_ok(int max_tiles, int n_stages);
cudaError_t mlp_forward_persist_tc(const PersistParams& pp, int max_tiles, cudaStream_t st);
// Debug: the `skip`-th next launch of kernel family `kernel` (0 forward, 1 backward chain, 2 backward input)
// records timestamps into d_buf [4 roles][kTraceLen] (u64 nanoseconds).
constexpr int kTraceLen = 512;
void set_trace(unsigned long long* d_buf, int kernel, int skip);
unsigned long long* take_trace(int kernel);

// ---- fused MLP backward ----------------------------------------------------------------------------
// Chain kernel: LayerNorm backward (or a precomputed top-level dZ image) followed by `nsteps` steps;
// step j handles Dense layer l = top - j:   dW_l += H_{l-1}^T dZ_l  (accumulated in TMEM over all tiles
// of the CTA),  dZ_{l-1} = (dZ_l W_l^T) .* (H_{l-1} > 0),  db = column sums of every dZ.  The last dZ
// (dZ of layer top - nsteps) is written as a tile image for the input kernel.
enum HeadMode { HEAD_LN = 0, HEAD_IMAGE = 1 };
constexpr int kMaxSteps = kMaxLayers - 1;

struct ChainParams {
  int n_tiles;
  int64_t M;
  const int32_t* tile_row_start;       // nullable (plain 128-row tiles)
  int head_mode;
  // HEAD_LN: dy[r] = dy_a[r]  (fp32 row-major: node MLPs, encoders)   or
  //          dy[r] = dy_a_img[r] + dy_b16[b_idx ? b_idx[r] : r]   (edge MLPs; either term may be null)
  const float* dy_a;
  const __nv_bfloat16* dy_a_img;       // bf16 tile images (gradient of the edge latent)
  const __nv_bfloat16* dy_b16;         // bf16 row-major [nodes][128] (gradient of the aggregated messages), gathered
  const int32_t* b_idx;
  const __nv_bfloat16* xhat;           // image
  const float* rstd;                   // [rows]
  const float* ln_scale;               // [128]
  // HEAD_IMAGE
  const __nv_bfloat16* z_top;          // image
  int nsteps;                          // 1..kMaxSteps
  const __nv_bfloat16* h_img[kMaxSteps];   // step j: H_{l-1} image
  const __nv_bfloat16* wt_img[kMaxSteps];  // step j: W_l^T image (2 tiles)
  __nv_bfloat16* dy_out_img;           // optional: dy (as used by the head) written back as a bf16 tile image - in place over
                                       // dy_a_img: with aggregate_post_residual the residual path carries d_ef + d_agg[recv]
  __nv_bfloat16* dz_out;               // image of the last dZ
  float* partial;                      // [grid][chain_partial_floats(nsteps)]
  unsigned long long* trace;           // debug (see FwdParams::trace)
  int pdl;
};
// per-CTA partial layout: dW[j] at j*16384 ; db[i] at nsteps*16384 + i*128 (i = 0: top dZ, i = j+1: dZ
// produced by step j) ; g_scale, g_bias after the db block.
__host__ __device__ inline size_t chain_partial_floats(int nsteps) { return (size_t)nsteps * 16384 + (size_t)(nsteps + 1) * 128 + 256; }
int backward_grid(int n_tiles);   // CTAs a backward kernel launches for n_tiles tiles
cudaError_t mlp_backward_chain_tc(const ChainParams& p, int* grid_out, cudaStream_t st);

// Input kernel: first Dense layer of an MLP whose input is made of 128-wide bf16 blocks.
//   dW_0[block b] += X_b^T dZ_0 ; dX_b = dZ_0 W_0[block b]^T -> sink b
enum SinkMode { SINK_NONE = 0, SINK_STORE_BF16 = 1, SINK_ADD_F32 = 2, SINK_SEGSUM_F32 = 3,
                SINK_STORE_IMG = 4, /* bf16 tile image [tile][2][16 KB]: one bulk store of the staged tile */
                SINK_ADD_IMG = 5    /* dst image = bf16(src image + dX): the tile's own rows, in place allowed */ };
struct InputParams {
  int n_tiles;
  int64_t M;
  const int32_t* tile_row_start;
  const int32_t* tile_node_start;      // SINK_SEGSUM_F32
  const int32_t* row_ptr;
  const __nv_bfloat16* dz0;            // image
  int nblk;                            // 1..3
  const __nv_bfloat16* x[3];           // block b source, row-major [*][128] ...
  int x_is_img[3];                     // ... or tile images [tile][2][16 KB] (identity rows; one bulk copy per tile)
  const int32_t* idx[3];               // optional row gather
  const __nv_bfloat16* wt_img;         // W_0^T image: tile (nb, kb) at (nb * 2 + kb) * 16 KB
  int sink[3];
  float* f32_dst[3];                   // SINK_ADD_F32: dst[r] = (src ? src[r] : 0) + dX ; SINK_SEGSUM_F32: per node
  const float* f32_src[3];
  __nv_bfloat16* bf16_dst[3];
  const __nv_bfloat16* img_src[3];     // SINK_ADD_IMG
  float* partial;                      // [grid][nblk * 16384]
  unsigned long long* trace;           // debug (see FwdParams::trace)
  int pdl;
};
cudaError_t mlp_backward_input_tc(const InputParams& p, int* grid_out, cudaStream_t st);

// ---- small CUDA-core helpers of the backward pass ---------------------------------------------------
struct Piece { const float* src; int64_t stride; int32_t n_parts; float* dst; int64_t count; };
constexpr int kMaxPieces = 12;
struct Pieces { Piece p[kMaxPieces]; int n; };
// For every piece: dst[i] = sum_k src[k * stride + i], k < n_parts   (fixed order: deterministic).  One launch reduces
// the partials of all kernels of one MLP (chain + input layer, or decoder head + chain + input layer).
cudaError_t reduce_pieces(const Pieces& pieces, cudaStream_t st, bool pdl);
// Decoder head: dZ_{L-2} = (dout W_{L-1}^T) .* (H_{L-2} > 0) as an image, plus per-tile partials of
// dW_{L-1} [128][od], db_{L-1} [od] and db_{L-2} [128]  (stride 128*od + od + 128 floats per tile).
// out_feat / val_mask: the cotangent is first pulled back through `inverse_data(...) .* val_mask` (fused output).
cudaError_t decoder_head_bwd(const float* dout, int out_dim, const float* w_last, const __nv_bfloat16* h_img,
                             int n_tiles, int64_t M, __nv_bfloat16* z_img, float* partial, const FeatRecipe& out_feat,
                             const float* val_mask, cudaStream_t st, bool pdl);
// Encoder input layer: dW_0 [F][128] per-tile partials from the dZ_0 image and the raw fp32 features;
// d_raw [rows][F] = dZ_0 W_0^T when requested.
// raw features come from a recipe; d_raw is the gradient w.r.t. the recipe's SOURCE columns (transposed normaliser applied).
cudaError_t encoder_input_bwd(const __nv_bfloat16* dz0, const FeatRecipe& feat, const int32_t* raw_idx, int F,
                              const float* w0, int n_tiles, int64_t M, const int32_t* tile_row_start,
                              float* partial, float* d_raw, cudaStream_t st, bool pdl);
// d_nf[v] += recv_sum[v] + sum over CSC row v of dxs[csc_slot[j]]  (adjoints of the receiver and sender gathers;
// recv_sum is the tile-local segmented sum the input kernel stored; fixed order: deterministic)
// dxs is a tile-image tensor; csc_pos maps a CSC entry to its row in image space (tile * 128 + row in tile).
cudaError_t sender_gather_add(float* d_nf, const float* recv_sum, const __nv_bfloat16* dxs_img, const int32_t* col_ptr,
                              const int32_t* csc_pos, int64_t N, cudaStream_t st, bool pdl);

}  // namespace tc
}  // namespace mgn
