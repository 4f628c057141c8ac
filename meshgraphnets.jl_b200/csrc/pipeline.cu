// Orchestration of the Encode-Process-Decode forward / backward over the caller's workspace.
//
//   forward  = `mgn.model(graph, ps, st)`                    (src/solve.jl:200, step! src/strategies.jl:421)
//   backward = the Zygote pullback of the same call           (src/strategies.jl:189-194, :421)
//
// Model structure (GraphNetCore.jl, recalled - see DESIGN.md "Assumptions"):
//   Encoder   nf0 = LN(MLP(nf)), ef0 = LN(MLP(ef))
//   Processor m = LN(MLP_e([nf[s]; nf[r]; ef])); agg = scatter(+, m, r); n = LN(MLP_n([nf; agg]));
//             nf += n; ef += m        (x mps)
//   Decoder   out = MLP(nf)           (no LayerNorm)
//
// Every edge tensor inside the library lives in CSR order (sorted by receiver, stable), so the
// aggregation is a contiguous segmented sum and the receiver gather is near-sequential.
#include "common.cuh"
#include "tc.cuh"
#include "features.cuh"

namespace mgn {
namespace {

struct Bump {
  char* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<char*>(b)) {}
  float* f(size_t n) {
    const size_t bytes = (n * sizeof(float) + 255) & ~size_t(255);
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

struct MlpSaved {
  float* h[kMaxDense];  // post-ReLU outputs of Dense layers 0..L-2
  float* xhat = nullptr;
  float* rstd = nullptr;
};

struct Workspace {
  std::vector<float*> nf, ef;          // [mps+1] latents (training) or [2] ping-pong
  std::vector<float*> agg;             // [mps] or [1]
  std::vector<MlpSaved> saved;         // one per MLP (training) or one shared
  float* msg = nullptr;                // [E][D] LayerNorm output of the edge MLP
  // fused build_graph / inverse_data (mgn_forward_fused): the fp32 mode materialises the recipes with one launch each
  float *feat_n = nullptr, *feat_e = nullptr, *dout_f = nullptr;
  // backward scratch
  float *dzA = nullptr, *dzB = nullptr, *dxn = nullptr, *dxe = nullptr, *d_nf = nullptr,
        *d_ef = nullptr, *dw_partial = nullptr, *ln_partial = nullptr;
  size_t bytes = 0;
};

size_t max_dw_partial(const mgn_model* m, int64_t N, int64_t E) {
  size_t mx = 0;
  for (size_t i = 0; i < m->mlps.size(); ++i) {
    const MlpLayout& L = m->mlps[i];
    const bool edge_rows = (i == 1) || (i >= 2 && i + 1 < m->mlps.size() && ((i - 2) % 2 == 0));
    const int64_t M = edge_rows ? E : N;
    for (int l = 0; l < L.n_dense; ++l) mx = std::max(mx, dw_partial_floats(L.in[l], L.out[l], M));
  }
  return mx;
}

void layout(const mgn_model* m, const mgn_graph* g, bool training, void* base, Workspace& w) {
  const int64_t N = g->N, E = g->E;
  const int D = m->cfg.latent, mps = m->cfg.mps, L = m->n_dense();
  const int64_t R = std::max(N, E);
  Bump b(base);
  const int nlat = training ? mps + 1 : 2;
  w.nf.resize(nlat);
  w.ef.resize(nlat);
  for (int k = 0; k < nlat; ++k) {
    w.nf[k] = b.f((size_t)N * D);
    w.ef[k] = b.f((size_t)E * D);
  }
  w.agg.resize(training ? mps : 1);
  for (auto& a : w.agg) a = b.f((size_t)N * D);
  w.msg = b.f((size_t)E * D);
  w.feat_n = b.f((size_t)N * m->cfg.node_in);
  w.feat_e = b.f((size_t)std::max<int64_t>(E, 1) * m->cfg.edge_in);
  w.dout_f = b.f((size_t)N * m->cfg.out_dim);
  const size_t n_mlps = m->mlps.size();
  w.saved.resize(training ? n_mlps : 1);
  for (size_t i = 0; i < w.saved.size(); ++i) {
    int64_t rows = R;
    if (training) {
      const bool edge_rows = (i == 1) || (i >= 2 && i + 1 < n_mlps && ((i - 2) % 2 == 0));
      rows = edge_rows ? E : N;
    }
    for (int l = 0; l < L - 1; ++l) w.saved[i].h[l] = b.f((size_t)rows * D);
    if (training) {
      w.saved[i].xhat = b.f((size_t)rows * D);
      w.saved[i].rstd = b.f((size_t)rows);
    }
  }
  if (training) {
    w.dzA = b.f((size_t)R * D);
    w.dzB = b.f((size_t)R * D);
    w.dxn = b.f((size_t)N * 2 * D);
    w.dxe = b.f((size_t)E * 3 * D);
    w.d_nf = b.f((size_t)N * D);
    w.d_ef = b.f((size_t)E * D);
    w.dw_partial = b.f(max_dw_partial(m, N, E));
    w.ln_partial = b.f(ln_partial_floats(R, D));
  }
  w.bytes = b.off;
}

Operand op1(const float* p, const int32_t* idx, int width, int ld) {
  Operand o{};
  o.nseg = 1;
  o.s[0] = {p, idx, width, ld};
  return o;
}

// Dense chain of one MLP.  `final_out`: decoder output (no LayerNorm) or nullptr.
cudaError_t mlp_forward(const MlpLayout& L, const float* params, const Operand& in, int64_t M,
                        MlpSaved& sv, float ln_eps, float* ln_out, const float* resid_in,
                        float* resid_out, float* final_out, cudaStream_t st) {
  Operand x = in;
  for (int l = 0; l < L.n_dense; ++l) {
    const float* W = params + L.w_off[l];
    const float* bias = params + L.b_off[l];
    const bool last = (l == L.n_dense - 1);
    cudaError_t e;
    if (!last) {
      e = dense_forward(x, M, W, bias, L.out[l], true, sv.h[l], nullptr, st);
      x = op1(sv.h[l], nullptr, L.out[l], L.out[l]);
    } else if (L.layer_norm) {
      LnEpilogue ln{params + L.ln_scale_off, params + L.ln_bias_off, ln_eps, sv.xhat, sv.rstd,
                    ln_out, resid_in, resid_out};
      e = dense_forward(x, M, W, bias, L.out[l], false, nullptr, &ln, st);
    } else {
      e = dense_forward(x, M, W, bias, L.out[l], false, final_out, nullptr, st);
    }
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Reverse of mlp_forward.  dout = a[r] + b[bidx[r]] (either may be null).  dx_in (nullable)
// receives the gradient w.r.t. the MLP input, [M][in_dim].
cudaError_t mlp_backward(const MlpLayout& L, const float* params, float* grads, const Operand& in,
                         int64_t M, const MlpSaved& sv, const float* a, int lda, const float* b,
                         int ldb, const int32_t* bidx, Workspace& w, float* dx_in,
                         cudaStream_t st) {
  cudaError_t e;
  const float* dz = nullptr;
  float* nxt = w.dzA;
  if (L.layer_norm) {
    e = layernorm_backward(a, lda, b, ldb, bidx, sv.xhat, sv.rstd, params + L.ln_scale_off, M,
                           L.out_dim, w.dzA, w.ln_partial, grads + L.ln_scale_off,
                           grads + L.ln_bias_off, st);
    if (e != cudaSuccess) return e;
    dz = w.dzA;
    nxt = w.dzB;
  } else {
    dz = a;  // decoder: dout is dz of the last Dense layer
  }
  for (int l = L.n_dense - 1; l >= 0; --l) {
    const Operand x = l == 0 ? in : op1(sv.h[l - 1], nullptr, L.in[l], L.in[l]);
    e = dense_backward_dw(x, dz, M, L.out[l], w.dw_partial, grads + L.w_off[l], st);
    if (e != cudaSuccess) return e;
    if (l > 0) {
      e = dense_backward_dx(dz, M, L.out[l], params + L.w_off[l], L.in[l], sv.h[l - 1], nxt, st);
      if (e != cudaSuccess) return e;
      dz = nxt;
      nxt = (nxt == w.dzA) ? w.dzB : w.dzA;
    } else if (dx_in) {
      e = dense_backward_dx(dz, M, L.out[l], params + L.w_off[l], L.in[l], nullptr, dx_in, st);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

}  // namespace

int32_t workspace_bytes(const mgn_model* m, const mgn_graph* g, bool training, size_t* bytes) {
  if (m->cfg.compute_mode == MGN_COMPUTE_BF16) return tc_workspace_bytes(m, g, training, bytes);
  Workspace w;
  layout(m, g, training, nullptr, w);
  *bytes = w.bytes;
  return MGN_OK;
}

constexpr int kStageAll = -3;   // run every stage (mgn_forward / mgn_backward)

int32_t forward_stage(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                      const float* ef, float* out, void* ws, size_t ws_bytes, bool training, int stage,
                      cudaStream_t st, const FusedIo* io) {
  if (m->cfg.compute_mode == MGN_COMPUTE_BF16)
    return tc_forward_stage(m, g, params, nf, ef, out, ws, ws_bytes, training, stage, st, io);
  Workspace w;
  layout(m, g, training, ws, w);
  if (w.bytes > ws_bytes) return fail(MGN_ERR_WORKSPACE, "workspace too small for mgn_forward");
  const int64_t N = g->N, E = g->E;
  const int D = m->cfg.latent, mps = m->cfg.mps;
  const float eps = m->cfg.ln_eps;
  auto saved = [&](size_t i) -> MlpSaved& { return w.saved[training ? i : 0]; };
  auto lat = [&](int k) { return training ? k : (k & 1); };  // latent buffer read by MP step k
  const bool all = stage == kStageAll;

  if (all || stage == MGN_STAGE_ENCODE) {
    if (io) {  // normalise + concat (src/graph.jl:75-97) materialised by one launch per entity kind
      MGN_CUDA_TRY(build_features(io->node, N, w.feat_n, st));
      MGN_CUDA_TRY(build_features(io->edge, E, w.feat_e, st));
      nf = w.feat_n;
      ef = w.feat_e;
    }
    // Encoder.  Raw edge features arrive in original order: gather through perm.
    MGN_CUDA_TRY(mlp_forward(m->mlps[0], params, op1(nf, nullptr, m->cfg.node_in, m->cfg.node_in), N,
                             saved(0), eps, w.nf[0], nullptr, nullptr, nullptr, st));
    MGN_CUDA_TRY(mlp_forward(m->mlps[1], params, op1(ef, g->perm, m->cfg.edge_in, m->cfg.edge_in), E,
                             saved(1), eps, w.ef[0], nullptr, nullptr, nullptr, st));
  }
  for (int k = 0; k < mps; ++k) {
    if (!all && stage != k) continue;
    const int cur = lat(k), nxt = lat(k + 1);
    float* agg = w.agg[training ? k : 0];
    // edge update (a10) + residual (a12)
    Operand xe{};
    xe.nseg = 3;
    xe.s[0] = {w.nf[cur], g->send_csr, D, D};
    xe.s[1] = {w.nf[cur], g->recv_csr, D, D};
    xe.s[2] = {w.ef[cur], nullptr, D, D};
    MGN_CUDA_TRY(mlp_forward(m->mlps[2 + 2 * k], params, xe, E, saved(2 + 2 * k), eps, w.msg,
                             w.ef[cur], w.ef[nxt], nullptr, st));
    // aggregate the pre-residual messages (a11) - or, with aggregate_post_residual, the updated edge latent
    MGN_CUDA_TRY(segment_sum(m->cfg.aggregate_post_residual ? w.ef[nxt] : w.msg, g->row_ptr, N, D, agg, st));
    // node update (a12)
    Operand xn{};
    xn.nseg = 2;
    xn.s[0] = {w.nf[cur], nullptr, D, D};
    xn.s[1] = {agg, nullptr, D, D};
    MGN_CUDA_TRY(mlp_forward(m->mlps[3 + 2 * k], params, xn, N, saved(3 + 2 * k), eps, nullptr,
                             w.nf[cur], w.nf[nxt], nullptr, st));
  }
  if (all || stage == MGN_STAGE_DECODE) {  // Decoder (a13)
    const size_t di = m->mlps.size() - 1;
    MGN_CUDA_TRY(mlp_forward(m->mlps[di], params, op1(w.nf[lat(mps)], nullptr, D, D), N, saved(di), eps,
                             nullptr, nullptr, nullptr, out, st));
    if (io && (io->out.n > 0 || io->val_mask))  // inverse_data(...) .* val_mask (src/solve.jl:205-218)
      MGN_CUDA_TRY(finish_output(io->out, io->val_mask, N, m->cfg.out_dim, out, st));
  }
  return MGN_OK;
}

int32_t backward_stage(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                       const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                       size_t ws_bytes, int stage, cudaStream_t st, GradHook* hook, const FusedIo* io) {
  if (m->cfg.compute_mode == MGN_COMPUTE_BF16)
    return tc_backward_stage(m, g, params, nf, ef, dout, dparams, dnf, ws, ws_bytes, stage, st, hook, io);
  auto done = [&](size_t mi) -> int32_t { return hook ? hook->mlp_done(mi, st) : MGN_OK; };
  Workspace w;
  layout(m, g, true, ws, w);
  if (w.bytes > ws_bytes) return fail(MGN_ERR_WORKSPACE, "workspace too small for mgn_backward");
  const int64_t N = g->N, E = g->E;
  const int D = m->cfg.latent, mps = m->cfg.mps;
  const size_t di = m->mlps.size() - 1;
  const bool all = stage == kStageAll;

  if (io) {  // the features the matching forward materialised
    nf = w.feat_n;
    ef = w.feat_e;
  }
  if (all || stage == MGN_STAGE_DECODE) {
    if (io && (io->out.n > 0 || io->val_mask)) {  // pull the cotangent back through inverse_data(...) .* val_mask
      MGN_CUDA_TRY(prepare_dout(io->out, io->val_mask, dout, N, m->cfg.out_dim, w.dout_f, st));
      dout = w.dout_f;
    }
    // Decoder: d_nf = d(out)/d(nf[mps])
    MGN_CUDA_TRY(mlp_backward(m->mlps[di], params, dparams, op1(w.nf[mps], nullptr, D, D), N,
                              w.saved[di], dout, m->cfg.out_dim, nullptr, 0, nullptr, w, w.d_nf, st));
    MGN_TRY(done(di));
  }
  for (int k = mps - 1; k >= 0; --k) {
    if (!all && stage != k) continue;
    const bool d_ef_valid = k != mps - 1;  // the decoder does not read the edge latent
    // node update: nf[k+1] = nf[k] + LN(MLP_n([nf[k]; agg[k]]))
    Operand xn{};
    xn.nseg = 2;
    xn.s[0] = {w.nf[k], nullptr, D, D};
    xn.s[1] = {w.agg[k], nullptr, D, D};
    MGN_CUDA_TRY(mlp_backward(m->mlps[3 + 2 * k], params, dparams, xn, N, w.saved[3 + 2 * k], w.d_nf,
                              D, nullptr, 0, nullptr, w, w.dxn, st));
    MGN_TRY(done(3 + 2 * k));
    // edge update: ef[k+1] = ef[k] + m, agg = segsum(m)  =>  dm[j] = d_ef[j] + d_agg[recv[j]]
    Operand xe{};
    xe.nseg = 3;
    xe.s[0] = {w.nf[k], g->send_csr, D, D};
    xe.s[1] = {w.nf[k], g->recv_csr, D, D};
    xe.s[2] = {w.ef[k], nullptr, D, D};
    MGN_CUDA_TRY(mlp_backward(m->mlps[2 + 2 * k], params, dparams, xe, E, w.saved[2 + 2 * k],
                              d_ef_valid ? w.d_ef : nullptr, D, w.dxn + D, 2 * D, g->recv_csr, w,
                              w.dxe, st));
    MGN_TRY(done(2 + 2 * k));
    // d_nf[k] = d_nf[k+1] + d(node MLP)/d(nf) + gathers' adjoints (receiver: CSR, sender: CSC)
    MGN_CUDA_TRY(node_grad_gather(w.d_nf, w.dxn, 2 * D, w.dxe, g->row_ptr, g->col_ptr, g->csc_slot,
                                  N, D, w.d_nf, st));
    // residual path of the edge latent; aggregate_post_residual: agg = segsum(ef[k+1]), so d_agg[recv] rides along
    if (m->cfg.aggregate_post_residual)
      MGN_CUDA_TRY(add_cols(d_ef_valid ? w.d_ef : nullptr, w.dxe, 3 * D, 2 * D, E, D, w.d_ef, st, w.dxn, 2 * D, D,
                            g->recv_csr));
    else
      MGN_CUDA_TRY(add_cols(d_ef_valid ? w.d_ef : nullptr, w.dxe, 3 * D, 2 * D, E, D, w.d_ef, st));
  }
  if (all || stage == MGN_STAGE_ENCODE) {
    if (mps > 0) {
      MGN_CUDA_TRY(mlp_backward(m->mlps[1], params, dparams,
                                op1(ef, g->perm, m->cfg.edge_in, m->cfg.edge_in), E, w.saved[1], w.d_ef,
                                D, nullptr, 0, nullptr, w, nullptr, st));
    } else {
      const MlpLayout& L = m->mlps[1];
      const int64_t sz = m->mlps[2].w_off[0] - L.w_off[0];
      MGN_CUDA_TRY(cudaMemsetAsync(dparams + L.w_off[0], 0, sizeof(float) * sz, st));
    }
    MGN_TRY(done(1));
    MGN_CUDA_TRY(mlp_backward(m->mlps[0], params, dparams,
                              op1(nf, nullptr, m->cfg.node_in, m->cfg.node_in), N, w.saved[0], w.d_nf, D,
                              nullptr, 0, nullptr, w, dnf, st));
    if (io && dnf) MGN_CUDA_TRY(finish_dx(io->node, N, dnf, st));  // transposed normalisers: gradient w.r.t. the raw fields
    MGN_TRY(done(0));
  }
  return MGN_OK;
}

int32_t forward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                const float* ef, float* out, void* ws, size_t ws_bytes, bool training,
                cudaStream_t st, const FusedIo* io) {
  return forward_stage(m, g, params, nf, ef, out, ws, ws_bytes, training, kStageAll, st, io);
}

int32_t backward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                 const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                 size_t ws_bytes, cudaStream_t st, GradHook* hook, const FusedIo* io) {
  return backward_stage(m, g, params, nf, ef, dout, dparams, dnf, ws, ws_bytes, kStageAll, st, hook, io);
}

// Rows of the node latent / its gradient, for the halo exchange of graph-partitioned meshes.
int32_t halo_rows(const mgn_model* m, const mgn_graph* g, void* ws, size_t ws_bytes, bool training, int what,
                  int step, const int32_t* rows, int64_t n_rows, void* buf, int op, cudaStream_t st) {
  if (m->cfg.compute_mode == MGN_COMPUTE_BF16)
    return tc_halo_rows(m, g, ws, ws_bytes, training, what, step, rows, n_rows, buf, op, st);
  Workspace w;
  layout(m, g, training, ws, w);
  if (w.bytes > ws_bytes) return fail(MGN_ERR_WORKSPACE, "workspace too small for mgn_halo_rows");
  float* base = nullptr;
  if (what == MGN_HALO_LATENT) {
    if (step < 0 || step > m->cfg.mps) return fail(MGN_ERR_INVALID, "halo_rows: bad step");
    base = w.nf[training ? step : (step & 1)];
  } else if (what == MGN_HALO_GRAD) {
    if (!training) return fail(MGN_ERR_INVALID, "halo_rows: gradients need a training workspace");
    base = w.d_nf;
  } else {
    return fail(MGN_ERR_INVALID, "halo_rows: unknown tensor");
  }
  MGN_CUDA_TRY(rows_op(base, 4, m->cfg.latent, rows, n_rows, g->N, buf, op, st));
  return MGN_OK;
}

}  // namespace mgn
