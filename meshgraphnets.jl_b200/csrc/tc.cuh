// tcgen05 (5th-gen tensor core) path: bf16 operands, fp32 accumulation in TMEM.
#pragma once
#include "common.cuh"

namespace mgn {

// Model-handle hooks (abi.cu): build / free the packed-weight image plan
int32_t tc_model_init(mgn_model* m);
void tc_model_free(mgn_model* m);

// Entry points used by pipeline.cu
int32_t tc_workspace_bytes(const mgn_model* m, const mgn_graph* g, bool training, size_t* bytes);
int32_t tc_forward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                   const float* ef, float* out, void* ws, size_t ws_bytes, bool training,
                   cudaStream_t st);
int32_t tc_backward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                    const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                    size_t ws_bytes, cudaStream_t st);
int32_t tc_backward_scratch_bytes(const mgn_model* m, const mgn_graph* g, size_t* bytes);

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

cudaError_t pack_weights(const ModelImages& im, const float* params, __nv_bfloat16* images, cudaStream_t st);

// ---- fused MLP forward -----------------------------------------------------------------------------
enum InMode { IN_RAW = 0, IN_PLAIN = 1, IN_CONCAT2 = 2, IN_GATHER3 = 3 };
enum FinMode { FIN_LN = 0, FIN_LN_RESID = 1, FIN_LN_RESID_AGG = 2, FIN_LINEAR = 3 };

struct FwdParams {
  // tiling
  int n_tiles;
  int64_t M;                          // rows
  const int32_t* tile_row_start;      // [n_tiles+1] (nullptr: tile t covers rows [128 t, 128 t + 128))
  const int32_t* tile_node_start;     // [n_tiles+1] FIN_LN_RESID_AGG: first node of each tile
  const int32_t* row_ptr;             // CSR row pointer (aggregation)
  // input operand
  int in_mode;
  const __nv_bfloat16 *x0, *x1, *x2;  // row-major [rows][128] bf16
  const int32_t *idx0, *idx1;         // IN_GATHER3: rows of x0 for K-blocks {0,1} / {2,3}
  const float* raw;                   // IN_RAW: fp32 [rows][raw_F]
  const int32_t* raw_idx;             // IN_RAW: optional row gather (CSR perm)
  int raw_F;
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
  const float* lat_in;                // fp32 [rows][128] residual input
  float* lat_out;                     // fp32 [rows][128]
  __nv_bfloat16* lat_bf16_out;        // bf16 shadow of lat_out
  __nv_bfloat16* agg_bf16;            // [nodes][128]
  float* out;                         // FIN_LINEAR: [rows][out_dim]
  int out_dim;
  // training saves (nullptr when not training)
  __nv_bfloat16* save_h[kMaxLayers - 1];  // image [tile][2 tiles]
  __nv_bfloat16* save_xhat;               // image [tile][2 tiles]
  float* save_rstd;                       // [rows]
};

cudaError_t mlp_forward_tc(const FwdParams& p, cudaStream_t st);

}  // namespace tc
}  // namespace mgn
