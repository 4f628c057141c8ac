// C-ABI entry points of libmgn_b200.so (see include/mgn_b200.h for the reference interface each
// one replaces) and the host-side integer graph indexing that must match the reference
// bit-exactly (one_hot, triangles_to_edges, parse_edges, the 0->1 based shift).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <unordered_set>

#include "common.cuh"
#include "tc.cuh"

namespace mgn {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int32_t fail(int32_t code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// ---- profiler --------------------------------------------------------------------------------
static const char* kTagNames[TAG_COUNT] = {
    "simt_gemm_fwd", "simt_gemm_dx", "simt_dw", "reduce_partials", "segment_sum", "ln_bwd", "ln_reduce",
    "node_grad_gather", "add_cols", "loss", "adam", "normaliser", "tc_pack", "tc_mlp_fwd", "tc_mlp_bwd",
    "tc_dw", "tc_misc", "solver"};
const char* tag_name(int tag) { return tag >= 0 && tag < TAG_COUNT ? kTagNames[tag] : "?"; }

// Diagnostic only (bench.py's launch count / per-family timing): off by default, and when off the launch path reads
// one relaxed atomic and nothing else.  While a session is on, every access is serialised by `mu`.
struct Profiler {
  std::atomic<bool> on{false};
  std::mutex mu;
  int tag = -1;
  int64_t launches = 0;
  int64_t per_tag[TAG_COUNT] = {};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
};
static Profiler g_prof;

ProfScope::ProfScope(int tag, cudaStream_t s) : st(s), slot(-1) {
  if (!g_prof.on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  ++g_prof.launches;
  if (tag >= 0 && tag < TAG_COUNT) ++g_prof.per_tag[tag];
  if (tag != g_prof.tag) return;
  std::pair<cudaEvent_t, cudaEvent_t> ev;
  if (!g_prof.pool.empty()) {
    ev = g_prof.pool.back();
    g_prof.pool.pop_back();
  } else {
    cudaEventCreate(&ev.first);
    cudaEventCreate(&ev.second);
  }
  cudaEventRecord(ev.first, st);
  slot = (int)g_prof.events.size();
  g_prof.events.push_back(ev);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  if (slot < (int)g_prof.events.size()) cudaEventRecord(g_prof.events[slot].second, st);
}

static void add_mlp(mgn_model* m, const std::string& name, int in_dim, int out_dim, bool ln,
                    int64_t& off) {
  MlpLayout L;
  L.name = name;
  L.in_dim = in_dim;
  L.out_dim = out_dim;
  L.layer_norm = ln;
  L.n_dense = m->n_dense();
  const int D = m->cfg.latent;
  for (int l = 0; l < L.n_dense; ++l) {
    L.in[l] = l == 0 ? in_dim : D;
    L.out[l] = l == L.n_dense - 1 ? out_dim : D;
    L.w_off[l] = off;
    off += (int64_t)L.in[l] * L.out[l];
    L.b_off[l] = off;
    off += L.out[l];
  }
  if (ln) {
    // Lux 0.5 LayerNorm parameters are (bias, scale) in that order (recalled); cfg.ln_scale_first flips it
    if (m->cfg.ln_scale_first) {
      L.ln_scale_off = off;
      L.ln_bias_off = off + out_dim;
    } else {
      L.ln_bias_off = off;
      L.ln_scale_off = off + out_dim;
    }
    off += 2 * out_dim;
  }
  m->mlps.push_back(L);
}

}  // namespace mgn

using namespace mgn;

extern "C" {

int32_t mgn_abi_version(void) { return MGN_ABI_VERSION; }

int32_t mgn_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return MGN_ERR_INVALID;
  std::snprintf(buf, n, "%s", g_last_error.c_str());
  return MGN_OK;
}

int32_t mgn_device_count(int32_t* count) {
  if (!count) return fail(MGN_ERR_INVALID, "count is null");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  *count = c;
  return MGN_OK;
}

// ------------------------------------------------------------------------------------------
// Host integer path
// ------------------------------------------------------------------------------------------
int32_t mgn_one_hot(const int32_t* h_v, int64_t n, int32_t depth, int32_t offset, float* h_out) {
  MGN_REQUIRE(n >= 0 && depth > 0, "one_hot: bad sizes");
  MGN_REQUIRE((h_v && h_out) || n == 0, "one_hot: null pointer");
  std::memset(h_out, 0, sizeof(float) * (size_t)n * depth);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t j = (int64_t)h_v[i] + offset - 1;  // Julia row v+offset (1-based)
    if (j >= 0 && j < depth) h_out[i * depth + j] = 1.0f;
  }
  return MGN_OK;
}

int32_t mgn_triangles_to_edges(const int32_t* h_cells, int64_t n_cells, int32_t* h_senders,
                               int32_t* h_receivers, int64_t* n_edges) {
  MGN_REQUIRE(n_cells >= 0 && n_edges, "triangles_to_edges: bad arguments");
  MGN_REQUIRE((h_cells && h_senders && h_receivers) || n_cells == 0, "triangles_to_edges: null pointer");
  // edges = [f0f1 for all faces ; f1f2 ... ; f2f0 ...], stored (max, min), unique in
  // first-occurrence order, then made two-way.
  std::unordered_set<uint64_t> seen;
  seen.reserve((size_t)n_cells * 3);
  std::vector<int32_t> hi, lo;
  hi.reserve((size_t)n_cells * 3);
  lo.reserve((size_t)n_cells * 3);
  static const int pa[3] = {0, 1, 2}, pb[3] = {1, 2, 0};
  for (int side = 0; side < 3; ++side) {
    for (int64_t c = 0; c < n_cells; ++c) {
      const int32_t a = h_cells[c * 3 + pa[side]], b = h_cells[c * 3 + pb[side]];
      const int32_t mx = std::max(a, b), mn = std::min(a, b);
      const uint64_t key = ((uint64_t)(uint32_t)mx << 32) | (uint32_t)mn;
      if (seen.insert(key).second) {
        hi.push_back(mx);
        lo.push_back(mn);
      }
    }
  }
  const int64_t U = (int64_t)hi.size();
  for (int64_t i = 0; i < U; ++i) {
    h_senders[i] = hi[i];
    h_senders[U + i] = lo[i];
    h_receivers[i] = lo[i];
    h_receivers[U + i] = hi[i];
  }
  *n_edges = 2 * U;
  return MGN_OK;
}

int32_t mgn_parse_edges(const int32_t* h_edges, int64_t n_pairs, int32_t* h_senders,
                        int32_t* h_receivers) {
  MGN_REQUIRE(n_pairs >= 0, "parse_edges: bad size");
  MGN_REQUIRE((h_edges && h_senders && h_receivers) || n_pairs == 0, "parse_edges: null pointer");
  for (int64_t i = 0; i < n_pairs; ++i) {
    const int32_t s = h_edges[2 * i], r = h_edges[2 * i + 1];
    h_senders[i] = s;
    h_senders[n_pairs + i] = r;
    h_receivers[i] = r;
    h_receivers[n_pairs + i] = s;
  }
  return MGN_OK;
}

int32_t mgn_shift_one_based(int32_t* h_senders, int32_t* h_receivers, int64_t n_edges,
                            int32_t* shifted) {
  MGN_REQUIRE(n_edges >= 0, "shift_one_based: bad size");
  MGN_REQUIRE((h_senders && h_receivers) || n_edges == 0, "shift_one_based: null pointer");
  bool has0 = false;
  for (int64_t i = 0; i < n_edges && !has0; ++i) has0 = h_senders[i] == 0 || h_receivers[i] == 0;
  if (has0)
    for (int64_t i = 0; i < n_edges; ++i) {
      h_senders[i] += 1;
      h_receivers[i] += 1;
    }
  if (shifted) *shifted = has0 ? 1 : 0;
  return MGN_OK;
}

int32_t mgn_edge_features(const float* h_pos, int64_t n_nodes, int32_t dim, const int32_t* h_senders,
                          const int32_t* h_receivers, int64_t n_edges, int32_t index_base,
                          float* h_out) {
  MGN_REQUIRE(n_nodes >= 0 && n_edges >= 0 && dim > 0, "edge_features: bad sizes");
  MGN_REQUIRE((h_pos && h_senders && h_receivers && h_out) || n_edges == 0, "edge_features: null pointer");
  for (int64_t e = 0; e < n_edges; ++e) {
    const int64_t s = (int64_t)h_senders[e] - index_base, r = (int64_t)h_receivers[e] - index_base;
    if (s < 0 || s >= n_nodes || r < 0 || r >= n_nodes)
      return fail(MGN_ERR_INDEX, "edge_features: node id out of range");
    double acc = 0.0;  // LinearAlgebra.norm widens the accumulator, rounds once
    for (int d = 0; d < dim; ++d) {
      const float rel = h_pos[s * dim + d] - h_pos[r * dim + d];
      h_out[e * (dim + 1) + d] = rel;
      acc += (double)rel * (double)rel;
    }
    h_out[e * (dim + 1) + dim] = (float)std::sqrt(acc);
  }
  return MGN_OK;
}

// ------------------------------------------------------------------------------------------
// Graph handle
// ------------------------------------------------------------------------------------------
int32_t mgn_graph_create(int64_t n_nodes, int64_t n_edges, const int32_t* d_senders,
                         const int32_t* d_receivers, int32_t index_base, void* stream,
                         mgn_graph** out) {
  MGN_REQUIRE(out, "graph_create: out is null");
  *out = nullptr;
  MGN_REQUIRE(n_nodes > 0 && n_edges >= 0, "graph_create: bad sizes");
  MGN_REQUIRE(n_nodes < (int64_t)1 << 31 && n_edges < (int64_t)1 << 31, "graph_create: sizes exceed Int32");
  MGN_REQUIRE((d_senders && d_receivers) || n_edges == 0, "graph_create: null index pointer");
  mgn_graph* g = new mgn_graph();
  g->N = n_nodes;
  g->E = n_edges;
  g->index_base = index_base;
  const int32_t s = build_graph_index(g, d_senders, d_receivers, static_cast<cudaStream_t>(stream));
  if (s != MGN_OK) {
    mgn_graph_destroy(g);
    return s;
  }
  *out = g;
  return MGN_OK;
}

int32_t mgn_graph_destroy(mgn_graph* g) {
  if (!g) return MGN_OK;
  cudaFree(g->row_ptr);
  cudaFree(g->perm);
  cudaFree(g->send_csr);
  cudaFree(g->recv_csr);
  cudaFree(g->col_ptr);
  cudaFree(g->perm_sender);
  cudaFree(g->csc_slot);
  cudaFree(g->tile_row_start);
  cudaFree(g->csc_pos);
  cudaFree(g->tile_node_start);
  delete g;
  return MGN_OK;
}

int32_t mgn_graph_sizes(const mgn_graph* g, int64_t* n_nodes, int64_t* n_edges) {
  MGN_REQUIRE(g, "graph_sizes: null graph");
  if (n_nodes) *n_nodes = g->N;
  if (n_edges) *n_edges = g->E;
  return MGN_OK;
}

int32_t mgn_graph_get_index(const mgn_graph* g, int32_t* h_row_ptr, int32_t* h_perm,
                            int32_t* h_col_ptr, int32_t* h_perm_sender) {
  MGN_REQUIRE(g, "graph_get_index: null graph");
  if (h_row_ptr)
    MGN_CUDA_TRY(cudaMemcpy(h_row_ptr, g->row_ptr, sizeof(int32_t) * (g->N + 1), cudaMemcpyDeviceToHost));
  if (h_col_ptr)
    MGN_CUDA_TRY(cudaMemcpy(h_col_ptr, g->col_ptr, sizeof(int32_t) * (g->N + 1), cudaMemcpyDeviceToHost));
  if (h_perm && g->E)
    MGN_CUDA_TRY(cudaMemcpy(h_perm, g->perm, sizeof(int32_t) * g->E, cudaMemcpyDeviceToHost));
  if (h_perm_sender && g->E)
    MGN_CUDA_TRY(cudaMemcpy(h_perm_sender, g->perm_sender, sizeof(int32_t) * g->E, cudaMemcpyDeviceToHost));
  return MGN_OK;
}

// ------------------------------------------------------------------------------------------
// Model
// ------------------------------------------------------------------------------------------
int32_t mgn_model_create(const mgn_model_config* cfg, mgn_model** out) {
  MGN_REQUIRE(cfg && out, "model_create: null argument");
  *out = nullptr;
  MGN_REQUIRE(cfg->node_in > 0 && cfg->edge_in > 0 && cfg->out_dim > 0, "model_create: feature sizes must be positive");
  MGN_REQUIRE(cfg->latent > 0 && cfg->latent <= 128, "model_create: latent must be in [1, 128]");
  MGN_REQUIRE(cfg->out_dim <= 128, "model_create: out_dim must be <= 128");
  MGN_REQUIRE(cfg->mps >= 0 && cfg->hidden_layers >= 0 && cfg->hidden_layers + 2 <= kMaxDense,
              "model_create: bad mps / hidden_layers");
  MGN_REQUIRE(cfg->dense_layers == 0 || (cfg->dense_layers >= 2 && cfg->dense_layers <= kMaxDense),
              "model_create: dense_layers must be 0 (= hidden_layers + 2) or in [2, 8]");
  MGN_REQUIRE(cfg->ln_scale_first == 0 || cfg->ln_scale_first == 1, "model_create: ln_scale_first must be 0 or 1");
  MGN_REQUIRE(cfg->aggregate_post_residual == 0 || cfg->aggregate_post_residual == 1,
              "model_create: aggregate_post_residual must be 0 or 1");
  MGN_REQUIRE(cfg->compute_mode == MGN_COMPUTE_FP32 || cfg->compute_mode == MGN_COMPUTE_BF16,
              "model_create: unknown compute_mode");
  if (cfg->compute_mode == MGN_COMPUTE_BF16)
    MGN_REQUIRE(cfg->latent == 128, "model_create: MGN_COMPUTE_BF16 needs latent == 128");
  mgn_model* m = new mgn_model();
  m->cfg = *cfg;
  m->knobs = read_tune_knobs();  // the only place the environment is read
  int64_t off = 0;
  const int D = cfg->latent;
  add_mlp(m, "encoder.node", cfg->node_in, D, true, off);
  add_mlp(m, "encoder.edge", cfg->edge_in, D, true, off);
  for (int k = 0; k < cfg->mps; ++k) {
    add_mlp(m, "processor" + std::to_string(k + 1) + ".edge", 3 * D, D, true, off);
    add_mlp(m, "processor" + std::to_string(k + 1) + ".node", 2 * D, D, true, off);
  }
  add_mlp(m, "decoder", D, cfg->out_dim, false, off);
  m->n_params = off;
  if (cfg->compute_mode == MGN_COMPUTE_BF16) {
    const int32_t s = tc_model_init(m);
    if (s != MGN_OK) {
      mgn_model_destroy(m);
      return s;
    }
  }
  *out = m;
  return MGN_OK;
}

int32_t mgn_model_destroy(mgn_model* m) {
  if (m) tc_model_free(m);
  delete m;
  return MGN_OK;
}

int32_t mgn_model_param_count(const mgn_model* m, int64_t* count) {
  MGN_REQUIRE(m && count, "param_count: null argument");
  *count = m->n_params;
  return MGN_OK;
}

int32_t mgn_model_param_layout(const mgn_model* m, mgn_param_entry* entries, int32_t capacity,
                               int32_t* n) {
  MGN_REQUIRE(m && n, "param_layout: null argument");
  int32_t k = 0;
  auto put = [&](const std::string& name, int64_t off, int rows, int cols) {
    if (entries && k < capacity) {
      mgn_param_entry& e = entries[k];
      std::snprintf(e.name, sizeof(e.name), "%s", name.c_str());
      e.offset = off;
      e.rows = rows;
      e.cols = cols;
    }
    ++k;
  };
  for (const MlpLayout& L : m->mlps) {
    for (int l = 0; l < L.n_dense; ++l) {
      put(L.name + ".dense" + std::to_string(l + 1) + ".weight", L.w_off[l], L.out[l], L.in[l]);
      put(L.name + ".dense" + std::to_string(l + 1) + ".bias", L.b_off[l], L.out[l], 1);
    }
    if (L.layer_norm) {
      if (m->cfg.ln_scale_first) {
        put(L.name + ".layernorm.scale", L.ln_scale_off, L.out_dim, 1);
        put(L.name + ".layernorm.bias", L.ln_bias_off, L.out_dim, 1);
      } else {
        put(L.name + ".layernorm.bias", L.ln_bias_off, L.out_dim, 1);
        put(L.name + ".layernorm.scale", L.ln_scale_off, L.out_dim, 1);
      }
    }
  }
  *n = k;
  return MGN_OK;
}

int32_t mgn_workspace_bytes(const mgn_model* m, const mgn_graph* g, int32_t training, size_t* bytes) {
  MGN_REQUIRE(m && g && bytes, "workspace_bytes: null argument");
  return workspace_bytes(m, g, training != 0, bytes);
}

int32_t mgn_forward(const mgn_model* m, const mgn_graph* g, const float* d_params, const float* d_nf,
                    const float* d_ef, float* d_out, void* d_workspace, size_t workspace_bytes,
                    int32_t training, void* stream) {
  MGN_REQUIRE(m && g && d_params && d_nf && d_out && d_workspace, "forward: null argument");
  MGN_REQUIRE(d_ef || g->E == 0, "forward: null edge features");
  return forward(m, g, d_params, d_nf, d_ef, d_out, d_workspace, workspace_bytes, training != 0,
                 static_cast<cudaStream_t>(stream));
}

int32_t mgn_backward(const mgn_model* m, const mgn_graph* g, const float* d_params, const float* d_nf,
                     const float* d_ef, const float* d_dout, float* d_dparams, float* d_dnf,
                     void* d_workspace, size_t workspace_bytes, void* stream) {
  MGN_REQUIRE(m && g && d_params && d_nf && d_dout && d_dparams && d_workspace, "backward: null argument");
  MGN_REQUIRE(d_ef || g->E == 0, "backward: null edge features");
  return backward(m, g, d_params, d_nf, d_ef, d_dout, d_dparams, d_dnf, d_workspace, workspace_bytes,
                  static_cast<cudaStream_t>(stream));
}

int32_t mgn_forward_stage(const mgn_model* m, const mgn_graph* g, const float* d_params, const float* d_nf,
                          const float* d_ef, float* d_out, void* d_workspace, size_t workspace_bytes,
                          int32_t training, int32_t stage, void* stream) {
  MGN_REQUIRE(m && g && d_params && d_workspace, "forward_stage: null argument");
  MGN_REQUIRE(stage == MGN_STAGE_ENCODE || stage == MGN_STAGE_DECODE || (stage >= 0 && stage < m->cfg.mps),
              "forward_stage: bad stage");
  MGN_REQUIRE(stage != MGN_STAGE_ENCODE || (d_nf && (d_ef || g->E == 0)), "forward_stage: null features");
  MGN_REQUIRE(stage != MGN_STAGE_DECODE || d_out, "forward_stage: null output");
  return forward_stage(m, g, d_params, d_nf, d_ef, d_out, d_workspace, workspace_bytes, training != 0, stage,
                       static_cast<cudaStream_t>(stream));
}

int32_t mgn_backward_stage(const mgn_model* m, const mgn_graph* g, const float* d_params, const float* d_nf,
                           const float* d_ef, const float* d_dout, float* d_dparams, float* d_dnf,
                           void* d_workspace, size_t workspace_bytes, int32_t stage, void* stream) {
  MGN_REQUIRE(m && g && d_params && d_dparams && d_workspace, "backward_stage: null argument");
  MGN_REQUIRE(stage == MGN_STAGE_ENCODE || stage == MGN_STAGE_DECODE || (stage >= 0 && stage < m->cfg.mps),
              "backward_stage: bad stage");
  MGN_REQUIRE(stage != MGN_STAGE_ENCODE || (d_nf && (d_ef || g->E == 0)), "backward_stage: null features");
  MGN_REQUIRE(stage != MGN_STAGE_DECODE || d_dout, "backward_stage: null output gradient");
  return backward_stage(m, g, d_params, d_nf, d_ef, d_dout, d_dparams, d_dnf, d_workspace, workspace_bytes, stage,
                        static_cast<cudaStream_t>(stream));
}

int32_t mgn_halo_row_bytes(const mgn_model* m, int32_t what, size_t* bytes) {
  MGN_REQUIRE(m && bytes, "halo_row_bytes: null argument");
  MGN_REQUIRE(what == MGN_HALO_LATENT || what == MGN_HALO_GRAD, "halo_row_bytes: unknown tensor");
  const bool bf16 = m->cfg.compute_mode == MGN_COMPUTE_BF16 && what == MGN_HALO_LATENT;
  *bytes = (size_t)m->cfg.latent * (bf16 ? 2 : 4);
  return MGN_OK;
}

int32_t mgn_halo_rows(const mgn_model* m, const mgn_graph* g, void* d_workspace, size_t workspace_bytes,
                      int32_t training, int32_t what, int32_t step, const int32_t* d_rows, int64_t n_rows,
                      void* d_buf, int32_t op, void* stream) {
  MGN_REQUIRE(m && g && d_workspace, "halo_rows: null argument");
  MGN_REQUIRE(n_rows >= 0 && (n_rows == 0 || (d_rows && d_buf)), "halo_rows: null rows / buffer");
  return halo_rows(m, g, d_workspace, workspace_bytes, training != 0, what, step, d_rows, n_rows, d_buf, op,
                   static_cast<cudaStream_t>(stream));
}

int32_t mgn_loss_mse_masked(const float* d_out, const float* d_target, int64_t n_nodes,
                            int32_t out_dim, const int32_t* d_mask, int64_t n_mask,
                            int32_t index_base, float* d_loss, float* d_dout, void* stream) {
  MGN_REQUIRE(d_out && d_target && d_mask && d_loss && d_dout, "loss: null argument");
  MGN_REQUIRE(n_nodes > 0 && out_dim > 0 && n_mask > 0, "loss: bad sizes");
  MGN_CUDA_TRY(loss_mse_masked(d_out, d_target, n_nodes, out_dim, d_mask, n_mask, index_base, d_loss,
                               d_dout, static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n,
                      float lr, float beta1, float beta2, float eps, int64_t t, void* stream) {
  MGN_REQUIRE(d_params && d_grads && d_m && d_v, "adam: null argument");
  MGN_REQUIRE(n >= 0 && t >= 1, "adam: bad n / t");
  MGN_CUDA_TRY(adam_step(d_params, d_grads, d_m, d_v, n, lr, beta1, beta2, eps, t,
                         static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_adam_step_device(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n,
                             float lr, float beta1, float beta2, float eps, void* d_state16,
                             void* stream) {
  MGN_REQUIRE(d_params && d_grads && d_m && d_v && d_state16, "adam_device: null argument");
  MGN_REQUIRE(n >= 0, "adam_device: bad n");
  MGN_CUDA_TRY(adam_step_device(d_params, d_grads, d_m, d_v, n, lr, beta1, beta2, eps, d_state16,
                                static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_profile_begin(int32_t tag) {
  MGN_REQUIRE(tag >= -1 && tag < TAG_COUNT, "profile_begin: unknown tag");
  std::lock_guard<std::mutex> lock(g_prof.mu);
  g_prof.tag = tag;
  g_prof.launches = 0;
  for (auto& c : g_prof.per_tag) c = 0;
  for (auto& e : g_prof.events) g_prof.pool.push_back(e);
  g_prof.events.clear();
  g_prof.on = true;
  return MGN_OK;
}

int32_t mgn_profile_end(int64_t* n_launches, int64_t* n_tagged, float* tagged_ms, int64_t* per_tag,
                        int32_t per_tag_capacity) {
  g_prof.on = false;
  MGN_CUDA_TRY(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_prof.mu);
  float total = 0.f;
  for (auto& e : g_prof.events) {
    float ms = 0.f;
    MGN_CUDA_TRY(cudaEventElapsedTime(&ms, e.first, e.second));
    total += ms;
  }
  if (n_launches) *n_launches = g_prof.launches;
  if (n_tagged) *n_tagged = (int64_t)g_prof.events.size();
  if (tagged_ms) *tagged_ms = total;
  if (per_tag)
    for (int i = 0; i < per_tag_capacity && i < TAG_COUNT; ++i) per_tag[i] = g_prof.per_tag[i];
  return MGN_OK;
}

int32_t mgn_profile_tag_name(int32_t tag, char* buf, size_t n) {
  MGN_REQUIRE(buf && n > 0, "profile_tag_name: null buffer");
  std::snprintf(buf, n, "%s", tag >= 0 && tag < TAG_COUNT ? tag_name(tag) : "");
  return tag >= 0 && tag < TAG_COUNT ? MGN_OK : MGN_ERR_INVALID;
}

int32_t mgn_norm_online_update(const float* d_x, int64_t rows, int32_t features, float* d_state,
                               float max_acc, void* stream) {
  MGN_REQUIRE(d_x && d_state && rows >= 0 && features > 0, "norm_online_update: bad argument");
  MGN_CUDA_TRY(norm_online_update(d_x, rows, features, d_state, max_acc, static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_norm_online_apply(const float* d_x, int64_t rows, int32_t features, const float* d_state,
                              float std_eps, int32_t inverse, float* d_y, int32_t ld_y, int32_t col_y,
                              void* stream) {
  MGN_REQUIRE(d_x && d_state && d_y && rows >= 0 && features > 0, "norm_online_apply: bad argument");
  MGN_REQUIRE(ld_y >= features + col_y && col_y >= 0, "norm_online_apply: bad ld_y / col_y");
  MGN_CUDA_TRY(norm_online_apply(d_x, rows, features, d_state, std_eps, inverse, d_y, ld_y, col_y,
                                 static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_affine_apply(const float* d_x, int64_t rows, int32_t features, float scale, float shift,
                         float* d_y, int32_t ld_y, int32_t col_y, void* stream) {
  MGN_REQUIRE(d_x && d_y && rows >= 0 && features > 0, "affine_apply: bad argument");
  MGN_REQUIRE(ld_y >= features + col_y && col_y >= 0, "affine_apply: bad ld_y / col_y");
  MGN_CUDA_TRY(affine_apply(d_x, rows, features, scale, shift, d_y, ld_y, col_y,
                            static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

// ------------------------------------------------------------------------------------------
// NeuralODE callers (src/solve.jl, SolverStrategy of src/strategies.jl)
// ------------------------------------------------------------------------------------------
int32_t mgn_ode_lincomb(const float* d_x, const float* const* h_k, const float* h_coef, int32_t n_terms, int64_t n,
                        float* d_y, void* stream) {
  MGN_REQUIRE(d_y && n >= 0, "ode_lincomb: bad argument");
  MGN_REQUIRE(n_terms >= 0 && n_terms <= kOdeMaxTerms, "ode_lincomb: n_terms must be in [0, 8]");
  MGN_REQUIRE(n_terms == 0 || (h_k && h_coef), "ode_lincomb: null term list");
  for (int j = 0; j < n_terms; ++j) MGN_REQUIRE(h_k[j], "ode_lincomb: null term");
  MGN_CUDA_TRY(ode_lincomb(d_x, h_k, h_coef, n_terms, n, d_y, static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_masked_overwrite(const float* d_x, const float* d_src, const uint8_t* d_mask, int64_t n, float* d_y,
                             void* stream) {
  MGN_REQUIRE(d_x && d_mask && d_y && n >= 0, "masked_overwrite: bad argument");
  MGN_CUDA_TRY(masked_overwrite(d_x, d_src, d_mask, n, d_y, static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_vec_mul(const float* d_a, const float* d_b, int64_t n, float* d_y, void* stream) {
  MGN_REQUIRE(d_a && d_b && d_y && n >= 0, "vec_mul: bad argument");
  MGN_CUDA_TRY(vec_mul(d_a, d_b, n, d_y, static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_norm_online_apply_ld(const float* d_x, int32_t ld_x, int32_t col_x, int64_t rows, int32_t features,
                                 const float* d_state, float std_eps, int32_t mode, float* d_y, int32_t ld_y,
                                 int32_t col_y, void* stream) {
  MGN_REQUIRE(d_x && d_state && d_y && rows >= 0 && features > 0, "norm_online_apply_ld: bad argument");
  MGN_REQUIRE(mode >= MGN_NORM_FORWARD && mode <= MGN_NORM_INVERSE_VJP, "norm_online_apply_ld: unknown mode");
  MGN_REQUIRE(col_x >= 0 && ld_x >= features + col_x, "norm_online_apply_ld: bad ld_x / col_x");
  MGN_REQUIRE(col_y >= 0 && ld_y >= features + col_y, "norm_online_apply_ld: bad ld_y / col_y");
  MGN_CUDA_TRY(norm_online_apply_ld(d_x, ld_x, col_x, rows, features, d_state, std_eps, mode, d_y, ld_y, col_y,
                                    static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_affine_apply_ld(const float* d_x, int32_t ld_x, int32_t col_x, int64_t rows, int32_t features,
                            float scale, float shift, float* d_y, int32_t ld_y, int32_t col_y, void* stream) {
  MGN_REQUIRE(d_x && d_y && rows >= 0 && features > 0, "affine_apply_ld: bad argument");
  MGN_REQUIRE(col_x >= 0 && ld_x >= features + col_x, "affine_apply_ld: bad ld_x / col_x");
  MGN_REQUIRE(col_y >= 0 && ld_y >= features + col_y, "affine_apply_ld: bad ld_y / col_y");
  MGN_CUDA_TRY(affine_apply_ld(d_x, ld_x, col_x, rows, features, scale, shift, d_y, ld_y, col_y,
                               static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_shooting_mse(const float* d_pred, const float* d_gt, const float* d_val_mask, int64_t n_saves,
                         int64_t mask_elems, float weight, int32_t accumulate, float* d_loss, float* d_dpred,
                         void* stream) {
  MGN_REQUIRE(d_pred && d_gt && d_val_mask && d_loss && d_dpred, "shooting_mse: null argument");
  MGN_REQUIRE(n_saves > 0 && mask_elems > 0, "shooting_mse: bad sizes");
  MGN_CUDA_TRY(shooting_mse(d_pred, d_gt, d_val_mask, mask_elems, n_saves * mask_elems, weight, accumulate, d_loss,
                            d_dpred, static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_shooting_continuity(const float* d_a, const float* d_b, int64_t n, float weight, float* d_loss,
                                float* d_da, void* stream) {
  MGN_REQUIRE(d_a && d_b && d_loss && d_da && n > 0, "shooting_continuity: bad argument");
  MGN_CUDA_TRY(shooting_continuity(d_a, d_b, n, weight, d_loss, d_da, static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

}  // extern "C"
