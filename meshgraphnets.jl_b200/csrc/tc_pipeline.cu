// Orchestration of the tensor-core (MGN_COMPUTE_BF16) forward / backward.
//
// Data layout in HBM (all owned by the caller's workspace):
//   images      bf16 weight images, 16 KB 128B-swizzled K-major tiles (repacked every forward)
//   nf32 / ef32 fp32 master copies of the node / edge latents, row-major [rows][128]
//   nf16 / ef16 bf16 shadows used as GEMM operands / gather sources, one per MP step when training
//   agg16       bf16 aggregated messages per MP step
//   saves       per MLP: hidden activations and LayerNorm xhat as tile images [tile][2][16 KB]
//               (written and read back with 1-D bulk copies), rstd fp32 [rows]
// Edge tensors are in CSR order and tiled node-aligned (mgn_graph::tile_row_start).
#include "tc.cuh"

#include <algorithm>

namespace mgn {

using namespace tc;

namespace {

struct Bump {
  char* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<char*>(b)) {}
  void* raw(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  float* f(size_t n) { return static_cast<float*>(raw(n * 4)); }
  __nv_bfloat16* h(size_t n) { return static_cast<__nv_bfloat16*>(raw(n * 2)); }
};

struct MlpSave {
  __nv_bfloat16* h[kMaxLayers - 1] = {nullptr, nullptr, nullptr};
  __nv_bfloat16* xhat = nullptr;
  float* rstd = nullptr;
};

struct TcWorkspace {
  __nv_bfloat16* images = nullptr;
  float *nf32 = nullptr, *ef32 = nullptr;
  std::vector<__nv_bfloat16*> nf16, ef16, agg16;
  std::vector<MlpSave> saves;
  size_t bytes = 0;
};

bool is_edge_mlp(const mgn_model* m, size_t i) {
  return i == 1 || (i >= 2 && i + 1 < m->mlps.size() && ((i - 2) % 2 == 0));
}

void tc_layout(const mgn_model* m, const mgn_graph* g, bool training, void* base, TcWorkspace& w) {
  const int64_t N = g->N, E = g->E;
  const int mps = m->cfg.mps, L = m->n_dense();
  const int64_t node_tiles = (N + kTile - 1) / kTile, edge_tiles = g->n_edge_tiles;
  Bump b(base);
  w.images = static_cast<__nv_bfloat16*>(b.raw((size_t)m->images->n_tiles * kTileB));
  w.nf32 = b.f((size_t)N * 128);
  w.ef32 = b.f((size_t)std::max<int64_t>(E, 1) * 128);
  const int nlat = training ? mps + 1 : 1;
  w.nf16.resize(nlat);
  w.ef16.resize(nlat);
  for (int k = 0; k < nlat; ++k) {
    w.nf16[k] = b.h((size_t)N * 128);
    w.ef16[k] = b.h((size_t)std::max<int64_t>(E, 1) * 128);
  }
  w.agg16.resize(training ? std::max(mps, 1) : 1);
  for (auto& a : w.agg16) a = b.h((size_t)N * 128);
  w.saves.resize(m->mlps.size());
  if (training) {
    for (size_t i = 0; i < m->mlps.size(); ++i) {
      const bool edge = is_edge_mlp(m, i);
      const int64_t tiles = edge ? edge_tiles : node_tiles, rows = edge ? E : N;
      for (int l = 0; l < L - 1; ++l) w.saves[i].h[l] = static_cast<__nv_bfloat16*>(b.raw((size_t)tiles * 2 * kTileB));
      if (m->mlps[i].layer_norm) {
        w.saves[i].xhat = static_cast<__nv_bfloat16*>(b.raw((size_t)tiles * 2 * kTileB));
        w.saves[i].rstd = b.f((size_t)std::max<int64_t>(rows, 1));
      }
    }
  }
  w.bytes = b.off;
}

void fill_layers(const mgn_model* m, size_t mi, const float* params, const TcWorkspace& w, bool training,
                 FwdParams& p) {
  const MlpLayout& L = m->mlps[mi];
  const MlpImages& im = m->images->mlps[mi];
  p.n_layers = L.n_dense;
  for (int l = 0; l < L.n_dense; ++l) {
    p.nkb[l] = im.nkb_f[l];
    p.wimg[l] = w.images + (size_t)im.fwd_off[l] * (kTileB / 2);
    p.bias[l] = params + L.b_off[l];
  }
  p.ksteps0 = 4;
  p.n_out_last = L.out_dim;
  p.ln_scale = L.layer_norm ? params + L.ln_scale_off : nullptr;
  p.ln_bias = L.layer_norm ? params + L.ln_bias_off : nullptr;
  p.eps = m->cfg.ln_eps;
  for (int l = 0; l < kMaxLayers - 1; ++l) p.save_h[l] = training ? w.saves[mi].h[l] : nullptr;
  p.save_xhat = training ? w.saves[mi].xhat : nullptr;
  p.save_rstd = training ? w.saves[mi].rstd : nullptr;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
int32_t tc_model_init(mgn_model* m) {
  MGN_REQUIRE(m->cfg.latent == 128, "MGN_COMPUTE_BF16 needs latent == 128");
  MGN_REQUIRE(m->n_dense() <= kMaxLayers, "MGN_COMPUTE_BF16 supports at most 2 hidden layers");
  MGN_REQUIRE(m->cfg.node_in <= 64 && m->cfg.edge_in <= 64, "MGN_COMPUTE_BF16 needs <= 64 input features");
  MGN_REQUIRE(m->cfg.out_dim <= 16, "MGN_COMPUTE_BF16 needs out_dim <= 16");
  ModelImages* im = new ModelImages();
  m->images = im;
  int off = 0;
  for (const MlpLayout& L : m->mlps) {
    MlpImages mi{};
    for (int l = 0; l < L.n_dense; ++l) {
      mi.fwd_off[l] = off;
      mi.nkb_f[l] = (L.in[l] + 63) / 64;
      for (int kb = 0; kb < mi.nkb_f[l]; ++kb) im->tiles.push_back({L.w_off[l], L.in[l], L.out[l], 0, kb, 0, 0});
      off += mi.nkb_f[l];
    }
    for (int l = 0; l < L.n_dense; ++l) {
      mi.bwd_off[l] = off;
      mi.nb_b[l] = (L.in[l] + 127) / 128;
      mi.nkb_b[l] = (L.out[l] + 63) / 64;
      for (int nb = 0; nb < mi.nb_b[l]; ++nb)
        for (int kb = 0; kb < mi.nkb_b[l]; ++kb) im->tiles.push_back({L.w_off[l], L.in[l], L.out[l], 1, kb, nb, 0});
      off += mi.nb_b[l] * mi.nkb_b[l];
    }
    im->mlps.push_back(mi);
  }
  im->n_tiles = off;
  MGN_CUDA_TRY(cudaMalloc(&im->d_tiles, sizeof(PackTile) * im->tiles.size()));
  MGN_CUDA_TRY(cudaMemcpy(im->d_tiles, im->tiles.data(), sizeof(PackTile) * im->tiles.size(), cudaMemcpyHostToDevice));
  return MGN_OK;
}

void tc_model_free(mgn_model* m) {
  if (!m->images) return;
  cudaFree(m->images->d_tiles);
  delete m->images;
  m->images = nullptr;
}

int32_t tc_workspace_bytes(const mgn_model* m, const mgn_graph* g, bool training, size_t* bytes) {
  TcWorkspace w;
  tc_layout(m, g, training, nullptr, w);
  size_t extra = 0;
  if (training) MGN_TRY(tc_backward_scratch_bytes(m, g, &extra));
  *bytes = w.bytes + extra;
  return MGN_OK;
}

int32_t tc_forward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                   const float* ef, float* out, void* ws, size_t ws_bytes, bool training,
                   cudaStream_t st) {
  if (!g->tiles_ok)
    return fail(MGN_ERR_UNSUPPORTED, "MGN_COMPUTE_BF16 needs every node to have at most 128 in-edges");
  TcWorkspace w;
  tc_layout(m, g, training, ws, w);
  if (w.bytes > ws_bytes) return fail(MGN_ERR_WORKSPACE, "workspace too small for mgn_forward (bf16)");
  const int64_t N = g->N, E = g->E;
  const int mps = m->cfg.mps;
  const int node_tiles = (int)((N + kTile - 1) / kTile);

  MGN_CUDA_TRY(pack_weights(*m->images, params, w.images, st));

  // Encoder (a9): raw fp32 features -> latent; edge features arrive in original order (perm gather)
  {
    FwdParams p{};
    fill_layers(m, 0, params, w, training, p);
    p.n_tiles = node_tiles;
    p.M = N;
    p.in_mode = IN_RAW;
    p.raw = nf;
    p.raw_F = m->cfg.node_in;
    p.ksteps0 = (m->cfg.node_in + 15) / 16;
    p.fin_mode = FIN_LN;
    p.lat_out = w.nf32;
    p.lat_bf16_out = w.nf16[0];
    MGN_CUDA_TRY(mlp_forward_tc(p, st));
  }
  if (E > 0) {
    FwdParams p{};
    fill_layers(m, 1, params, w, training, p);
    p.n_tiles = g->n_edge_tiles;
    p.M = E;
    p.tile_row_start = g->tile_row_start;
    p.in_mode = IN_RAW;
    p.raw = ef;
    p.raw_idx = g->perm;
    p.raw_F = m->cfg.edge_in;
    p.ksteps0 = (m->cfg.edge_in + 15) / 16;
    p.fin_mode = FIN_LN;
    p.lat_out = w.ef32;
    p.lat_bf16_out = w.ef16[0];
    MGN_CUDA_TRY(mlp_forward_tc(p, st));
  }
  for (int k = 0; k < mps; ++k) {
    const int cur = training ? k : 0, nxt = training ? k + 1 : 0;
    __nv_bfloat16* agg = w.agg16[training ? k : 0];
    if (E > 0) {  // edge update + residual + aggregation (a10, a11, a12)
      FwdParams p{};
      fill_layers(m, 2 + 2 * k, params, w, training, p);
      p.n_tiles = g->n_edge_tiles;
      p.M = E;
      p.tile_row_start = g->tile_row_start;
      p.tile_node_start = g->tile_node_start;
      p.row_ptr = g->row_ptr;
      p.in_mode = IN_GATHER3;
      p.x0 = w.nf16[cur];
      p.x2 = w.ef16[cur];
      p.idx0 = g->send_csr;
      p.idx1 = g->recv_csr;
      p.fin_mode = FIN_LN_RESID_AGG;
      p.lat_in = w.ef32;
      p.lat_out = w.ef32;
      p.lat_bf16_out = w.ef16[nxt];
      p.agg_bf16 = agg;
      MGN_CUDA_TRY(mlp_forward_tc(p, st));
    } else {
      MGN_CUDA_TRY(cudaMemsetAsync(agg, 0, (size_t)N * 128 * 2, st));
    }
    {  // node update + residual (a12)
      FwdParams p{};
      fill_layers(m, 3 + 2 * k, params, w, training, p);
      p.n_tiles = node_tiles;
      p.M = N;
      p.in_mode = IN_CONCAT2;
      p.x0 = w.nf16[cur];
      p.x1 = agg;
      p.fin_mode = FIN_LN_RESID;
      p.lat_in = w.nf32;
      p.lat_out = w.nf32;
      p.lat_bf16_out = w.nf16[nxt];
      MGN_CUDA_TRY(mlp_forward_tc(p, st));
    }
  }
  {  // Decoder (a13)
    const size_t di = m->mlps.size() - 1;
    FwdParams p{};
    fill_layers(m, di, params, w, training, p);
    p.n_tiles = node_tiles;
    p.M = N;
    p.in_mode = IN_PLAIN;
    p.x0 = w.nf16[training ? mps : 0];
    p.fin_mode = FIN_LINEAR;
    p.out = out;
    p.out_dim = m->cfg.out_dim;
    MGN_CUDA_TRY(mlp_forward_tc(p, st));
  }
  return MGN_OK;
}

}  // namespace mgn
