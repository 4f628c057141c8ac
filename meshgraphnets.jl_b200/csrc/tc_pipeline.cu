// Orchestration of the tensor-core (MGN_COMPUTE_BF16) forward / backward.
//
// Data layout in HBM (all owned by the caller's workspace):
//   images      bf16 weight images, 16 KB 128B-swizzled K-major tiles (repacked every forward)
//   nf32        fp32 master copy of the node latent, row-major [N][128]; nf16 = its bf16 shadow (gather source /
//               GEMM operand), one per MP step when training
//   ef16        the edge latent, stored in bf16 ONLY, as tile images (a tile only ever reads its own rows: bulk copies),
//               one per MP step when training.  No fp32 master: every MLP input is rounded to bf16 anyway, and the CPU
//               model of this arithmetic shows the fp32 edge master buys nothing at 15 MP steps (output error 7.0e-3
//               with and without it, gradient 1.65e-2 vs 1.71e-2) while costing 1024 B per edge
//               row and step of HBM traffic
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
  float* nf32 = nullptr;
  std::vector<__nv_bfloat16*> nf16, ef16, agg16;
  std::vector<MlpSave> saves;
  unsigned int* sync = nullptr;   // grid-barrier counter of the persistent forward kernel
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
  w.sync = static_cast<unsigned int*>(b.raw(1024));
  w.nf32 = b.f((size_t)N * 128);
  const int nlat = training ? mps + 1 : 1;
  w.nf16.resize(nlat);
  w.ef16.resize(nlat);
  for (int k = 0; k < nlat; ++k) {
    w.nf16[k] = b.h((size_t)N * 128);
    w.ef16[k] = static_cast<__nv_bfloat16*>(b.raw((size_t)std::max<int64_t>(edge_tiles, 1) * 2 * kTileB));  // tile images
  }
  w.agg16.resize(training ? std::max(mps, 1) : 1);
  for (auto& a : w.agg16) a = b.h((size_t)N * 128);
  w.saves.resize(m->mlps.size());
  if (training) {
    // MGN_RECOMPUTE=1 (TuneKnobs::recompute): the processor MLPs write no activation saves in the forward pass; the
    // backward pass re-runs an MLP (FIN_LN form: saves only) right before its own kernels, into ONE save set per kind of
    // MLP - 1 284 B per edge row and MP step less workspace (only the bf16 latents of every step stay), one extra MLP
    // forward per MLP and step of time.  The recomputed saves are the same bits, so are the gradients.
    const size_t first_proc = 2, last_proc = m->mlps.size() - 1;   // [first_proc, last_proc): (edge, node) x mps
    MlpSave shared[2];
    bool have_shared[2] = {false, false};   // (the pointers are null when only the size is computed)
    for (size_t i = 0; i < m->mlps.size(); ++i) {
      const bool edge = is_edge_mlp(m, i);
      const bool proc = i >= first_proc && i < last_proc;
      if (proc && m->knobs.recompute && have_shared[edge ? 0 : 1]) {
        w.saves[i] = shared[edge ? 0 : 1];
        continue;
      }
      const int64_t tiles = edge ? edge_tiles : node_tiles, rows = edge ? E : N;
      for (int l = 0; l < L - 1; ++l) w.saves[i].h[l] = static_cast<__nv_bfloat16*>(b.raw((size_t)tiles * 2 * kTileB));
      if (m->mlps[i].layer_norm) {
        w.saves[i].xhat = static_cast<__nv_bfloat16*>(b.raw((size_t)tiles * 2 * kTileB));
        w.saves[i].rstd = b.f((size_t)std::max<int64_t>(rows, 1));
      }
      if (proc && m->knobs.recompute) {
        shared[edge ? 0 : 1] = w.saves[i];
        have_shared[edge ? 0 : 1] = true;
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
  p.epi_warps = m->knobs.fwd_epi_warps;
  p.deep_ring = m->knobs.fwd_deep_ring;
  p.stagger_ns = (uint32_t)m->knobs.fwd_stagger_ns;
  p.pdl = m->knobs.pdl;
}

// Scratch of the backward pass, placed after the forward workspace.
struct BwdScratch {
  float *d_nf = nullptr, *recv_sum = nullptr;                // fp32: gradient of the node latent; receiver-adjoint sums
  __nv_bfloat16* d_agg = nullptr;                            // gradient of the aggregated messages, bf16 [N][128]: it is the
                                                             // staged (bf16) dX block of the node MLP, stored as it is
  __nv_bfloat16* d_ef = nullptr;                             // gradient of the edge latent: bf16 tile images (as the latent)
  __nv_bfloat16* dxs = nullptr;                              // sender adjoint rows as tile images [edge tile][2][16 KB]
  __nv_bfloat16 *dz0 = nullptr, *ztop = nullptr;             // tile images
  // weight-gradient partials of ONE MLP at a time: chain kernel (per CTA), input kernel (per CTA), CUDA-core helper
  // (per tile: decoder head or encoder input layer) - three regions so that one launch reduces them all
  // double buffered (set = MLP counter & 1): the reduction of one MLP's partials runs on the reduce lane while the
  // kernels of the next MLP already fill the other set
  float *partial_chain[2] = {nullptr, nullptr}, *partial_input[2] = {nullptr, nullptr}, *partial_misc[2] = {nullptr, nullptr};
  size_t bytes = 0;
};

void bwd_layout(const mgn_model* m, const mgn_graph* g, void* base, BwdScratch& b) {
  const int64_t N = g->N;
  const int64_t node_tiles = (N + kTile - 1) / kTile, edge_tiles = g->n_edge_tiles;
  const int64_t max_tiles = std::max<int64_t>(std::max(node_tiles, edge_tiles), 1);
  Bump bump(base);
  b.d_nf = bump.f((size_t)std::max<int64_t>(N, 1) * 128);
  b.d_ef = static_cast<__nv_bfloat16*>(bump.raw((size_t)std::max<int64_t>(edge_tiles, 1) * 2 * kTileB));
  b.recv_sum = bump.f((size_t)std::max<int64_t>(N, 1) * 128);
  b.d_agg = bump.h((size_t)std::max<int64_t>(N, 1) * 128);
  b.dxs = static_cast<__nv_bfloat16*>(bump.raw((size_t)std::max<int64_t>(edge_tiles, 1) * 2 * kTileB));  // tile images
  b.dz0 = static_cast<__nv_bfloat16*>(bump.raw((size_t)max_tiles * 2 * kTileB));
  b.ztop = static_cast<__nv_bfloat16*>(bump.raw((size_t)std::max<int64_t>(node_tiles, 1) * 2 * kTileB));
  const int grid = backward_grid((int)max_tiles);
  const size_t enc_f = (size_t)std::max(m->cfg.node_in, m->cfg.edge_in) * 128;
  for (int s = 0; s < 2; ++s) {
    b.partial_chain[s] = bump.f((size_t)grid * chain_partial_floats(kMaxSteps));
    b.partial_input[s] = bump.f((size_t)grid * 3 * 16384);
    b.partial_misc[s] = bump.f(std::max((size_t)max_tiles * enc_f,
                                        (size_t)node_tiles * (size_t)(128 * m->cfg.out_dim + m->cfg.out_dim + 128)));
  }
  b.bytes = bump.off;
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

int32_t tc_backward_scratch_bytes(const mgn_model* m, const mgn_graph* g, size_t* bytes) {
  BwdScratch b;
  bwd_layout(m, g, nullptr, b);
  *bytes = b.bytes;
  return MGN_OK;
}

constexpr int kStageAll = -3;

int32_t tc_forward_stage(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                         const float* ef, float* out, void* ws, size_t ws_bytes, bool training, int stage,
                         cudaStream_t st, const FusedIo* io) {
  if (!g->tiles_ok)
    return fail(MGN_ERR_UNSUPPORTED, "MGN_COMPUTE_BF16 needs every node to have at most 128 in-edges");
  TcWorkspace w;
  tc_layout(m, g, training, ws, w);
  if (w.bytes > ws_bytes) return fail(MGN_ERR_WORKSPACE, "workspace too small for mgn_forward (bf16)");
  const int64_t N = g->N, E = g->E;
  const int mps = m->cfg.mps;
  const int node_tiles = (int)((N + kTile - 1) / kTile);
  const bool all = stage == kStageAll;

  // ---- the FwdParams of every MLP of the pass
  auto no_saves = [](FwdParams& p) {   // MGN_RECOMPUTE: the backward pass re-runs the processor MLPs for their saves
    for (int l = 0; l < kMaxLayers - 1; ++l) p.save_h[l] = nullptr;
    p.save_xhat = nullptr;
    p.save_rstd = nullptr;
  };
  auto enc_node = [&]() {   // Encoder (a9): raw fp32 features -> latent
    FwdParams p{};
    fill_layers(m, 0, params, w, training, p);
    p.n_tiles = node_tiles;
    p.M = N;
    p.in_mode = IN_RAW;
    p.feat = io ? io->node : identity_recipe(nf, m->cfg.node_in);
    p.raw_F = m->cfg.node_in;
    p.ksteps0 = (m->cfg.node_in + 15) / 16;
    p.fin_mode = FIN_LN;
    p.lat_out = w.nf32;
    p.lat_bf16_out = w.nf16[0];
    return p;
  };
  auto enc_edge = [&]() {   // edge features arrive in original order (perm gather)
    FwdParams p{};
    fill_layers(m, 1, params, w, training, p);
    p.n_tiles = g->n_edge_tiles;
    p.M = E;
    p.tile_row_start = g->tile_row_start;
    p.in_mode = IN_RAW;
    p.feat = io ? io->edge : identity_recipe(ef, m->cfg.edge_in);
    p.raw_idx = g->perm;
    p.raw_F = m->cfg.edge_in;
    p.ksteps0 = (m->cfg.edge_in + 15) / 16;
    p.fin_mode = FIN_LN;
    p.lat_img_out = w.ef16[0];
    return p;
  };
  auto edge_step = [&](int k) {   // edge update + residual + aggregation (a10, a11, a12)
    const int cur = training ? k : 0, nxt = training ? k + 1 : 0;
    FwdParams p{};
    fill_layers(m, 2 + 2 * k, params, w, training, p);
    p.n_tiles = g->n_edge_tiles;
    p.M = E;
    p.tile_row_start = g->tile_row_start;
    p.tile_node_start = g->tile_node_start;
    p.row_ptr = g->row_ptr;
    p.in_mode = IN_GATHER3;
    p.x0 = w.nf16[cur];
    p.x2_img = w.ef16[cur];
    p.idx0 = g->send_csr;
    p.idx1 = g->recv_csr;
    p.fin_mode = FIN_LN_RESID_AGG;
    if (k + 1 < mps) {  // the edge latent after the last MP step is never read (the decoder takes the nodes only)
      p.lat_img_in = w.ef16[cur];   // residual = the bf16 latent itself (in place when not training: tile-local)
      p.lat_img_out = w.ef16[nxt];
    }
    p.agg_post_residual = m->cfg.aggregate_post_residual;
    if (p.agg_post_residual) p.lat_img_in = w.ef16[cur];   // ... but the aggregation of the last step still sums ef + m
    p.agg_bf16 = w.agg16[training ? k : 0];
    if (m->knobs.recompute) no_saves(p);
    return p;
  };
  auto node_step = [&](int k) {   // node update + residual (a12)
    const int cur = training ? k : 0, nxt = training ? k + 1 : 0;
    FwdParams p{};
    fill_layers(m, 3 + 2 * k, params, w, training, p);
    p.n_tiles = node_tiles;
    p.M = N;
    p.in_mode = IN_CONCAT2;
    p.x0 = w.nf16[cur];
    p.x1 = w.agg16[training ? k : 0];
    p.fin_mode = FIN_LN_RESID;
    p.lat_in = w.nf32;
    p.lat_out = w.nf32;
    p.lat_bf16_out = w.nf16[nxt];
    if (m->knobs.recompute) no_saves(p);
    return p;
  };
  auto decoder = [&]() {   // Decoder (a13)
    FwdParams p{};
    fill_layers(m, m->mlps.size() - 1, params, w, training, p);
    p.n_tiles = node_tiles;
    p.M = N;
    p.in_mode = IN_PLAIN;
    p.x0 = w.nf16[training ? mps : 0];
    p.fin_mode = FIN_LINEAR;
    p.out = out;
    p.out_dim = m->cfg.out_dim;
    if (io) {
      p.out_feat = io->out;
      p.val_mask = io->val_mask;
    }
    return p;
  };

  // ---- whole inference pass of a small graph (no more tiles than SMs): one persistent cooperative launch, the MLPs are
  //      its stages (tc_kernels.cu).  The buffers of an inference pass are the same in every MP step, so a template per
  //      kind of stage plus the stage's weights describes the pass: node encoder, edge encoder, edge step, node step,
  //      decoder, and the edge step of the LAST MP step (no residual, no latent output: only the aggregation is wanted).
  //      Policy (measured on the 1 885-node CylinderFlow mesh, 15 MP steps): plain launches 0.603 -> 0.587 ms per pass
  //      and 34 -> 2 launches for the host to issue; replayed inside a CUDA graph, where launch gaps are already gone,
  //      the launch-per-MLP form is faster (0.526 vs 0.564 ms) - so a call that is being captured keeps it.
  const int n_stages = 3 + 2 * mps;
  bool persist = all && !training && E > 0 && mps >= 1 && m->knobs.fwd_persist != 0 &&
                 forward_persist_ok(std::max(node_tiles, g->n_edge_tiles), n_stages);
  if (persist && m->knobs.fwd_persist == 1 && st != nullptr && st != cudaStreamLegacy) {
    // (the legacy default stream cannot be captured; querying it while another stream captures would invalidate that capture)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    MGN_CUDA_TRY(cudaStreamIsCapturing(st, &cs));
    persist = cs == cudaStreamCaptureStatusNone;
  }
  if (persist) {
    MGN_CUDA_TRY(pack_weights(*m->images, params, w.images, st, m->knobs.pdl != 0));
    static_assert(sizeof(PersistParams) < 32000, "kernel parameter space");
    PersistParams pp{};
    pp.sync = w.sync;
    const FwdParams en = enc_node(), ee = enc_edge(), de = decoder();
    pp.feat[0] = en.feat;
    pp.feat[1] = ee.feat;
    pp.feat[2] = de.out_feat;
    pp.stage[pp.n_stages++] = en;
    pp.stage[pp.n_stages++] = ee;
    for (int k = 0; k < mps; ++k) {
      pp.stage[pp.n_stages++] = edge_step(k);
      pp.stage[pp.n_stages++] = node_step(k);
    }
    pp.stage[pp.n_stages++] = de;
    const cudaError_t ce = mlp_forward_persist_tc(pp, std::max(node_tiles, g->n_edge_tiles), st);
    if (ce != cudaErrorCooperativeLaunchTooLarge) {
      MGN_CUDA_TRY(ce);
      return MGN_OK;
    }
    // Not every CTA can be resident (the process owns only part of the GPU: MPS thread percentage, a green context):
    // clear the launch error and run the pass with one launch per MLP below (the weights are packed again: harmless).
    (void)cudaGetLastError();
  }

  if (all || stage == MGN_STAGE_ENCODE) {
    MGN_CUDA_TRY(pack_weights(*m->images, params, w.images, st, m->knobs.pdl != 0));
    MGN_CUDA_TRY(mlp_forward_tc(enc_node(), st));
    if (E > 0) MGN_CUDA_TRY(mlp_forward_tc(enc_edge(), st));
  }
  for (int k = 0; k < mps; ++k) {
    if (!all && stage != k) continue;
    if (E > 0) MGN_CUDA_TRY(mlp_forward_tc(edge_step(k), st));
    else MGN_CUDA_TRY(cudaMemsetAsync(w.agg16[training ? k : 0], 0, (size_t)N * 128 * 2, st));
    MGN_CUDA_TRY(mlp_forward_tc(node_step(k), st));
  }
  if (all || stage == MGN_STAGE_DECODE) MGN_CUDA_TRY(mlp_forward_tc(decoder(), st));
  return MGN_OK;
}

int32_t tc_forward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                   const float* ef, float* out, void* ws, size_t ws_bytes, bool training,
                   cudaStream_t st, const FusedIo* io) {
  return tc_forward_stage(m, g, params, nf, ef, out, ws, ws_bytes, training, kStageAll, st, io);
}

// ---------------------------------------------------------------------------------------------------
// Backward: the Zygote pullback of mgn.model(graph, ps, st) (src/strategies.jl:189-194, :421).
// ---------------------------------------------------------------------------------------------------
namespace {

struct BwdCtx {
  const mgn_model* m;
  const mgn_graph* g;
  const float* params;
  float* grads;
  TcWorkspace* w;
  BwdScratch* b;
  cudaStream_t st;
  ReduceLane* lane = nullptr;   // nullptr: reductions run inline on st
  GradHook* hook = nullptr;
  mutable int set = 0;          // partial buffer set of the MLP in flight
  mutable int used[2] = {0, 0}; // lane->done[set] has been recorded in this call
  mutable bool forked = false;  // the lane has work of this call: join it before returning
  float* pchain() const { return b->partial_chain[set]; }
  float* pinput() const { return b->partial_input[set]; }
  float* pmisc() const { return b->partial_misc[set]; }
};

// Before the first kernel of an MLP writes into the current partial set: its previous reduction (two MLPs ago) must be done.
int32_t begin_mlp(const BwdCtx& c) {
  if (c.lane && c.used[c.set]) MGN_CUDA_TRY(cudaStreamWaitEvent(c.st, c.lane->done[c.set], 0));
  return MGN_OK;
}
// The fixed-order reduction of every partial of MLP `mi` (gathered in pc): on the reduce lane, beside the next MLP's
// kernels, or inline.  Then the gradient of MLP mi is final: notify the hook with the stream it is final on.
int32_t finish_mlp(const BwdCtx& c, size_t mi, const Pieces& pc) {
  cudaStream_t where = c.st;
  if (c.lane) {
    MGN_CUDA_TRY(cudaEventRecord(c.lane->fork[c.set], c.st));
    MGN_CUDA_TRY(cudaStreamWaitEvent(c.lane->side, c.lane->fork[c.set], 0));
    MGN_CUDA_TRY(reduce_pieces(pc, c.lane->side, false));
    MGN_CUDA_TRY(cudaEventRecord(c.lane->done[c.set], c.lane->side));
    c.used[c.set] = 1;
    c.forked = true;
    where = c.lane->side;
  } else {
    MGN_CUDA_TRY(reduce_pieces(pc, c.st, c.m->knobs.pdl != 0));
  }
  c.set ^= 1;
  return c.hook ? c.hook->mlp_done(mi, where) : MGN_OK;
}

// An MLP whose gradient is identically zero (no edges): cleared on the stream the other gradients become final on.
int32_t zero_mlp(const BwdCtx& c, size_t mi, int64_t lo, int64_t hi) {
  cudaStream_t where = c.st;
  if (c.lane) {
    MGN_CUDA_TRY(cudaEventRecord(c.lane->fork[c.set], c.st));
    MGN_CUDA_TRY(cudaStreamWaitEvent(c.lane->side, c.lane->fork[c.set], 0));
    c.forked = true;
    where = c.lane->side;
  }
  MGN_CUDA_TRY(cudaMemsetAsync(c.grads + lo, 0, sizeof(float) * (size_t)(hi - lo), where));
  return c.hook ? c.hook->mlp_done(mi, where) : MGN_OK;
}
int32_t join_lane(const BwdCtx& c) {
  if (c.lane && c.forked) {
    MGN_CUDA_TRY(cudaEventRecord(c.lane->join, c.lane->side));
    MGN_CUDA_TRY(cudaStreamWaitEvent(c.st, c.lane->join, 0));
  }
  return MGN_OK;
}

// MGN_RECOMPUTE: the saves (hidden activations, xhat, rstd) of processor MLP `mi` of MP step k are produced again - the MLP
// in its FIN_LN form, same operands as in the forward pass (the bf16 latents of every step are kept), no latent outputs.
int32_t recompute_saves(const BwdCtx& c, int k, bool edge) {
  const mgn_model* m = c.m;
  const mgn_graph* g = c.g;
  const size_t mi = (edge ? 2 : 3) + 2 * (size_t)k;
  FwdParams p{};
  fill_layers(m, mi, c.params, *c.w, true, p);
  p.fin_mode = FIN_LN;
  if (edge) {
    p.n_tiles = g->n_edge_tiles;
    p.M = g->E;
    p.tile_row_start = g->tile_row_start;
    p.in_mode = IN_GATHER3;
    p.x0 = c.w->nf16[k];
    p.x2_img = c.w->ef16[k];
    p.idx0 = g->send_csr;
    p.idx1 = g->recv_csr;
  } else {
    p.n_tiles = (int)((g->N + kTile - 1) / kTile);
    p.M = g->N;
    p.in_mode = IN_CONCAT2;
    p.x0 = c.w->nf16[k];
    p.x1 = c.w->agg16[k];
  }
  MGN_CUDA_TRY(mlp_forward_tc(p, c.st));
  return MGN_OK;
}

// Chain kernel + fixed-order reduction of its partials for MLP `mi`.  HEAD_LN when the MLP ends in a
// LayerNorm (dy = dy_a[r], or dy_a_img[r] + dy_b16[b_idx[r]] for edge rows); HEAD_IMAGE for the decoder (top dZ precomputed in b->ztop).
int32_t run_chain(const BwdCtx& c, size_t mi, bool edge_rows, const float* dy_a, const __nv_bfloat16* dy_b16,
                  const int32_t* b_idx, Pieces& pc, const __nv_bfloat16* dy_a_img = nullptr,
                  __nv_bfloat16* dy_out_img = nullptr) {
  const MlpLayout& L = c.m->mlps[mi];
  const MlpImages& im = c.m->images->mlps[mi];
  const MlpSave& sv = c.w->saves[mi];
  const int nd = L.n_dense;
  ChainParams p{};
  p.n_tiles = edge_rows ? c.g->n_edge_tiles : (int)((c.g->N + kTile - 1) / kTile);
  p.M = edge_rows ? c.g->E : c.g->N;
  p.tile_row_start = edge_rows ? c.g->tile_row_start : nullptr;
  int top;
  if (L.layer_norm) {
    p.head_mode = HEAD_LN;
    p.dy_a = dy_a;
    p.dy_a_img = dy_a_img;
    p.dy_out_img = dy_out_img;
    p.dy_b16 = dy_b16;
    p.b_idx = b_idx;
    p.xhat = sv.xhat;
    p.rstd = sv.rstd;
    p.ln_scale = c.params + L.ln_scale_off;
    top = nd - 1;
  } else {
    p.head_mode = HEAD_IMAGE;
    p.z_top = c.b->ztop;
    top = nd - 2;
  }
  p.nsteps = top;  // layers top .. 1
  if (p.nsteps == 0) return MGN_OK;
  for (int j = 0; j < p.nsteps; ++j) {
    const int l = top - j;
    p.h_img[j] = sv.h[l - 1];
    p.wt_img[j] = c.w->images + (size_t)im.bwd_off[l] * (kTileB / 2);
  }
  p.dz_out = c.b->dz0;
  p.partial = c.pchain();
  p.pdl = c.m->knobs.pdl;
  int grid = 0;
  MGN_CUDA_TRY(mlp_backward_chain_tc(p, &grid, c.st));
  const float* base = c.pchain();
  const int64_t stride = (int64_t)chain_partial_floats(p.nsteps), base_db = (int64_t)p.nsteps * 16384;
  for (int j = 0; j < p.nsteps; ++j) {
    const int l = top - j;
    pc.p[pc.n++] = {base + (int64_t)j * 16384, stride, grid, c.grads + L.w_off[l], 16384};
    pc.p[pc.n++] = {base + base_db + (int64_t)(j + 1) * 128, stride, grid, c.grads + L.b_off[l - 1], 128};
  }
  if (L.layer_norm) {
    pc.p[pc.n++] = {base + base_db, stride, grid, c.grads + L.b_off[top], 128};
    pc.p[pc.n++] = {base + base_db + (int64_t)(p.nsteps + 1) * 128, stride, grid, c.grads + L.ln_scale_off, 128};
    pc.p[pc.n++] = {base + base_db + (int64_t)(p.nsteps + 1) * 128 + 128, stride, grid, c.grads + L.ln_bias_off, 128};
  }
  return MGN_OK;
}

// Input kernel of MLP `mi`, then ONE fixed-order reduction of every partial of this MLP gathered in `pc`.
int32_t run_input(const BwdCtx& c, size_t mi, InputParams& p, Pieces& pc) {
  const MlpLayout& L = c.m->mlps[mi];
  const MlpImages& im = c.m->images->mlps[mi];
  p.dz0 = (L.layer_norm || L.n_dense > 2) ? c.b->dz0 : c.b->ztop;
  p.wt_img = c.w->images + (size_t)im.bwd_off[0] * (kTileB / 2);
  p.partial = c.pinput();
  p.pdl = c.m->knobs.pdl;
  int grid = 0;
  MGN_CUDA_TRY(mlp_backward_input_tc(p, &grid, c.st));
  pc.p[pc.n++] = {c.pinput(), (int64_t)p.nblk * 16384, grid, c.grads + L.w_off[0], (int64_t)p.nblk * 16384};
  return finish_mlp(c, mi, pc);
}

int32_t run_encoder_input(const BwdCtx& c, size_t mi, bool edge_rows, const FeatRecipe& raw, const int32_t* raw_idx,
                          float* d_raw, Pieces& pc) {
  const MlpLayout& L = c.m->mlps[mi];
  const int n_tiles = edge_rows ? c.g->n_edge_tiles : (int)((c.g->N + kTile - 1) / kTile);
  const int F = L.in[0];
  MGN_CUDA_TRY(encoder_input_bwd(c.b->dz0, raw, raw_idx, F, c.params + L.w_off[0], n_tiles,
                                 edge_rows ? c.g->E : c.g->N, edge_rows ? c.g->tile_row_start : nullptr,
                                 c.pmisc(), d_raw, c.st, c.m->knobs.pdl != 0));
  pc.p[pc.n++] = {c.pmisc(), (int64_t)F * 128, n_tiles, c.grads + L.w_off[0], (int64_t)F * 128};
  return finish_mlp(c, mi, pc);
}

}  // namespace

int32_t tc_backward_stage(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                          const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                          size_t ws_bytes, int stage, cudaStream_t st, GradHook* hook, const FusedIo* io) {
  if (!g->tiles_ok)
    return fail(MGN_ERR_UNSUPPORTED, "MGN_COMPUTE_BF16 needs every node to have at most 128 in-edges");
  TcWorkspace w;
  tc_layout(m, g, true, ws, w);
  BwdScratch b;
  bwd_layout(m, g, static_cast<char*>(ws) + w.bytes, b);
  if (w.bytes + b.bytes > ws_bytes) return fail(MGN_ERR_WORKSPACE, "workspace too small for mgn_backward (bf16)");
  const int64_t N = g->N, E = g->E;
  const int mps = m->cfg.mps, nd = m->n_dense(), od = m->cfg.out_dim;
  const int node_tiles = (int)((N + kTile - 1) / kTile);
  const bool all = stage == kStageAll;
  BwdCtx c{m, g, params, dparams, &w, &b, st};
  c.hook = hook;
  // whole-pass calls reduce the weight-gradient partials of MLP i beside the kernels of MLP i-1 (stage-wise calls keep
  // everything on the caller's stream: the halo exchanges in between are the caller's)
  c.lane = (all && m->knobs.reduce_lane) ? reduce_lane() : nullptr;

  if (all || stage == MGN_STAGE_DECODE) {
    // ---- Decoder (no LayerNorm): last Dense on CUDA cores, the rest on the tensor cores
    const size_t di = m->mlps.size() - 1;
    const MlpLayout& L = m->mlps[di];
    FeatRecipe no_out{};
    MGN_TRY(begin_mlp(c));
    MGN_CUDA_TRY(decoder_head_bwd(dout, od, params + L.w_off[nd - 1], w.saves[di].h[nd - 2], node_tiles, N, b.ztop,
                                  c.pmisc(), io ? io->out : no_out, io ? io->val_mask : nullptr, st, m->knobs.pdl != 0));
    Pieces pc{};
    const int64_t hs = (int64_t)128 * od + od + 128;
    pc.p[pc.n++] = {c.pmisc(), hs, node_tiles, dparams + L.w_off[nd - 1], (int64_t)128 * od};
    pc.p[pc.n++] = {c.pmisc() + (int64_t)128 * od, hs, node_tiles, dparams + L.b_off[nd - 1], od};
    pc.p[pc.n++] = {c.pmisc() + (int64_t)128 * od + od, hs, node_tiles, dparams + L.b_off[nd - 2], 128};
    MGN_TRY(run_chain(c, di, false, nullptr, nullptr, nullptr, pc));
    InputParams p{};
    p.n_tiles = node_tiles;
    p.M = N;
    p.nblk = 1;
    p.x[0] = w.nf16[mps];
    p.sink[0] = SINK_ADD_F32;
    p.f32_dst[0] = b.d_nf;
    MGN_TRY(run_input(c, di, p, pc));
  }
  for (int k = mps - 1; k >= 0; --k) {
    if (!all && stage != k) continue;
    const bool d_ef_valid = k != mps - 1;  // the decoder does not read the edge latent
    {  // node update: nf[k+1] = nf[k] + LN(MLP_n([nf[k]; agg[k]]))
      const size_t mi = 3 + 2 * k;
      Pieces pc{};
      MGN_TRY(begin_mlp(c));
      if (m->knobs.recompute) MGN_TRY(recompute_saves(c, k, false));
      MGN_TRY(run_chain(c, mi, false, b.d_nf, nullptr, nullptr, pc));
      InputParams p{};
      p.n_tiles = node_tiles;
      p.M = N;
      p.nblk = 2;
      p.x[0] = w.nf16[k];
      p.x[1] = w.agg16[k];
      p.sink[0] = SINK_ADD_F32;  // residual + direct path
      p.f32_src[0] = b.d_nf;
      p.f32_dst[0] = b.d_nf;
      p.sink[1] = SINK_STORE_BF16;  // gradient of the aggregated messages (bf16 rows: the staged tile as it is)
      p.bf16_dst[1] = b.d_agg;
      MGN_TRY(run_input(c, mi, p, pc));
    }
    if (E > 0) {  // edge update: ef[k+1] = ef[k] + m, agg = segsum(m)  =>  dm[j] = d_ef[j] + d_agg[recv[j]]
      const size_t mi = 2 + 2 * k;
      Pieces pc{};
      MGN_TRY(begin_mlp(c));
      if (m->knobs.recompute) MGN_TRY(recompute_saves(c, k, true));
      // aggregate_post_residual: agg = segsum(ef[k+1]), so the residual path carries d_ef + d_agg[recv] as well: the chain
      // head writes that sum back over the gradient image and the input kernel's sink adds dX to it
      const bool post = m->cfg.aggregate_post_residual != 0;
      MGN_TRY(run_chain(c, mi, true, nullptr, b.d_agg, g->recv_csr, pc, d_ef_valid ? b.d_ef : nullptr,
                        post ? b.d_ef : nullptr));
      InputParams p{};
      p.n_tiles = g->n_edge_tiles;
      p.M = E;
      p.tile_row_start = g->tile_row_start;
      p.tile_node_start = g->tile_node_start;
      p.row_ptr = g->row_ptr;
      p.nblk = 3;
      p.x[0] = w.nf16[k];
      p.idx[0] = g->send_csr;
      p.x[1] = w.nf16[k];
      p.idx[1] = g->recv_csr;
      p.x[2] = w.ef16[k];
      p.x_is_img[2] = 1;
      p.sink[0] = SINK_STORE_IMG;   // sender adjoint rows (tile image), gathered per node through the CSC below
      p.bf16_dst[0] = b.dxs;
      p.sink[1] = SINK_SEGSUM_F32;  // receiver adjoint: CSR segments are tile-local; plain stores, added to d_nf by the
      p.f32_dst[1] = b.recv_sum;    // gather below
      p.sink[2] = SINK_ADD_IMG;     // edge-latent residual: d_ef = bf16(d_ef + dX), tile images, in place
      p.img_src[2] = (d_ef_valid || post) ? b.d_ef : nullptr;
      p.bf16_dst[2] = b.d_ef;
      MGN_TRY(run_input(c, mi, p, pc));
      MGN_CUDA_TRY(sender_gather_add(b.d_nf, b.recv_sum, b.dxs, g->col_ptr, g->csc_pos, N, st, m->knobs.pdl != 0));
    } else {  // no edges: the edge MLP of this step has a zero gradient
      MGN_TRY(zero_mlp(c, 2 + 2 * k, m->mlps[2 + 2 * k].w_off[0], m->mlps[3 + 2 * k].w_off[0]));
    }
  }
  if (all || stage == MGN_STAGE_ENCODE) {
    const MlpLayout& L = m->mlps[1];
    if (mps > 0 && E > 0) {
      Pieces pc{};
      MGN_TRY(begin_mlp(c));
      MGN_TRY(run_chain(c, 1, true, nullptr, nullptr, nullptr, pc, b.d_ef));
      MGN_TRY(run_encoder_input(c, 1, true, io ? io->edge : identity_recipe(ef, m->cfg.edge_in), g->perm, nullptr, pc));
    } else {
      MGN_TRY(zero_mlp(c, 1, L.w_off[0], m->mlps[2].w_off[0]));
    }
    Pieces pc{};
    MGN_TRY(begin_mlp(c));
    MGN_TRY(run_chain(c, 0, false, b.d_nf, nullptr, nullptr, pc));
    MGN_TRY(run_encoder_input(c, 0, false, io ? io->node : identity_recipe(nf, m->cfg.node_in), nullptr, dnf, pc));
  }
  return join_lane(c);
}

int32_t tc_backward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                    const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                    size_t ws_bytes, cudaStream_t st, GradHook* hook, const FusedIo* io) {
  return tc_backward_stage(m, g, params, nf, ef, dout, dparams, dnf, ws, ws_bytes, kStageAll, st, hook, io);
}

int32_t tc_halo_rows(const mgn_model* m, const mgn_graph* g, void* ws, size_t ws_bytes, bool training, int what,
                     int step, const int32_t* rows, int64_t n_rows, void* buf, int op, cudaStream_t st) {
  TcWorkspace w;
  tc_layout(m, g, training, ws, w);
  if (what == MGN_HALO_LATENT) {
    if (step < 0 || step > m->cfg.mps) return fail(MGN_ERR_INVALID, "halo_rows: bad step");
    if (w.bytes > ws_bytes) return fail(MGN_ERR_WORKSPACE, "workspace too small for mgn_halo_rows");
    if (op == MGN_ROWS_ADD) return fail(MGN_ERR_INVALID, "halo_rows: the bf16 latent cannot be accumulated");
    MGN_CUDA_TRY(rows_op(w.nf16[training ? step : 0], 2, 128, rows, n_rows, g->N, buf, op, st));
    return MGN_OK;
  }
  if (what == MGN_HALO_GRAD) {
    if (!training) return fail(MGN_ERR_INVALID, "halo_rows: gradients need a training workspace");
    BwdScratch b;
    bwd_layout(m, g, static_cast<char*>(ws) + w.bytes, b);
    if (w.bytes + b.bytes > ws_bytes) return fail(MGN_ERR_WORKSPACE, "workspace too small for mgn_halo_rows");
    MGN_CUDA_TRY(rows_op(b.d_nf, 4, 128, rows, n_rows, g->N, buf, op, st));
    return MGN_OK;
  }
  return fail(MGN_ERR_INVALID, "halo_rows: unknown tensor");
}

}  // namespace mgn
