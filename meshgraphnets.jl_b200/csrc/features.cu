// Fused build_graph / inverse_data entry points (include/mgn_b200.h: mgn_forward_fused, mgn_backward_fused,
// mgn_norm_online_update_multi) and the one-launch materialisation kernels the fp32 mode uses (features.cuh).
#include <algorithm>

#include "features.cuh"

namespace mgn {
namespace {

__global__ void __launch_bounds__(256) build_features_kernel(const FeatRecipe R, int64_t rows, float* __restrict__ y) {
  __shared__ FeatCol tab[kMaxFeat];
  feat_table(R, tab, threadIdx.x, 256);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * R.F) return;
  const int64_t r = i / R.F;
  y[i] = feat_eval(tab[(int)(i - r * R.F)], r);
}

// mode 0: out = inverse_data(out) .* val_mask in place ; mode 1: y = pullback of that map applied to dout
__global__ void __launch_bounds__(256) output_map_kernel(const FeatRecipe R, const float* __restrict__ val_mask,
                                                         const float* __restrict__ x, int64_t rows, int od, int mode,
                                                         float* __restrict__ y) {
  __shared__ FeatCol tab[16];
  if (R.n > 0) feat_table(R, tab, threadIdx.x, 256);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * od) return;
  const int j = (int)(i % od);
  float v = x[i];
  if (mode == 0) {
    if (R.n > 0) v = out_eval(tab[j], v);
    if (val_mask) v = v * val_mask[i];
  } else {
    if (val_mask) v = v * val_mask[i];
    if (R.n > 0) v = out_vjp(tab[j], v);
  }
  y[i] = v;
}

__global__ void __launch_bounds__(256) finish_dx_kernel(const FeatRecipe R, int64_t rows, float* __restrict__ dx) {
  __shared__ FeatCol tab[kMaxFeat];
  feat_table(R, tab, threadIdx.x, 256);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * R.F) return;
  dx[i] = feat_vjp(tab[(int)(i % R.F)], dx[i]);
}

// ---- all online-normaliser updates of a step in two launches -------------------------------------------------------
constexpr int kMultiMax = 8;
constexpr int kMultiBlocks = 64;  // row blocks per normaliser
struct NormJob {
  const float* x;
  int64_t rows;
  int ld, col, F;
  float* state;
  float max_acc;
};
struct NormJobs {
  NormJob j[kMultiMax];
  int n;
};
// grid (kMultiBlocks, n): block (b, i) reduces its fixed row range of normaliser i to 2F partial sums (fixed order)
__global__ void __launch_bounds__(256) norm_multi_partial_kernel(const NormJobs jobs, float* __restrict__ partial) {
  __shared__ float red[8][2 * kMaxFeat];
  const NormJob& J = jobs.j[blockIdx.y];
  if (J.state[2 * J.F + 1] >= J.max_acc) return;  // uniform per block
  const int64_t per = (J.rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(J.rows, r0 + per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int f = 0; f < J.F; ++f) {
    float s = 0.f, q = 0.f;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += 256) {
      const float v = J.x[r * J.ld + J.col + f];
      s += v;
      q = fmaf(v, v, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      red[warp][f] = s;
      red[warp][J.F + f] = q;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * J.F) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[((int64_t)blockIdx.y * kMultiBlocks + blockIdx.x) * 2 * kMaxFeat + threadIdx.x] = t;
  }
}
__global__ void __launch_bounds__(128) norm_multi_finish_kernel(const NormJobs jobs, const float* __restrict__ partial) {
  const NormJob& J = jobs.j[blockIdx.x];
  if (J.state[2 * J.F + 1] >= J.max_acc) return;
  if (threadIdx.x < 2 * J.F) {
    float t = 0.f;
    for (int b = 0; b < kMultiBlocks; ++b) t += partial[((int64_t)blockIdx.x * kMultiBlocks + b) * 2 * kMaxFeat + threadIdx.x];
    J.state[threadIdx.x] += t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    J.state[2 * J.F] += (float)J.rows;
    J.state[2 * J.F + 1] += 1.f;
  }
}

inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

// mgn_feature_seg[] -> FeatRecipe, validated
int32_t to_recipe(const mgn_feature_seg* segs, int n, int want_F, bool need_x, const char* what, FeatRecipe& R) {
  if (n < 0 || n > kMaxFeatSegs) return fail(MGN_ERR_INVALID, std::string(what) + ": between 0 and 8 segments");
  R = FeatRecipe{};
  R.n = n;
  for (int i = 0; i < n; ++i) {
    const mgn_feature_seg& s = segs[i];
    if (s.width <= 0 || (need_x && (!s.d_x || s.ld < s.col + s.width || s.col < 0)))
      return fail(MGN_ERR_INVALID, std::string(what) + ": bad segment geometry");
    if (s.kind != MGN_FEAT_AFFINE && s.kind != MGN_FEAT_ONLINE)
      return fail(MGN_ERR_INVALID, std::string(what) + ": unknown segment kind");
    if (s.kind == MGN_FEAT_ONLINE && !s.d_state) return fail(MGN_ERR_INVALID, std::string(what) + ": online segment without a state");
    R.s[i] = {s.d_x, s.ld, s.col, s.width, s.kind, s.scale, s.shift, s.d_state, s.std_eps};
    R.F += s.width;
  }
  if (R.F != want_F)
    return fail(MGN_ERR_INVALID, std::string(what) + ": segment widths sum to " + std::to_string(R.F) + ", the model needs " +
                                     std::to_string(want_F));
  return MGN_OK;
}

int32_t to_io(const mgn_model* m, const mgn_graph* g, const mgn_fused_io* io, FusedIo& f) {
  MGN_REQUIRE(io, "fused: null io description");
  MGN_TRY(to_recipe(io->node, io->n_node_segs, m->cfg.node_in, true, "fused node features", f.node));
  if (g->E > 0 || io->n_edge_segs > 0) MGN_TRY(to_recipe(io->edge, io->n_edge_segs, m->cfg.edge_in, g->E > 0, "fused edge features", f.edge));
  else {
    f.edge = FeatRecipe{};
    f.edge.F = m->cfg.edge_in;
  }
  if (io->n_out_segs > 0) MGN_TRY(to_recipe(io->out, io->n_out_segs, m->cfg.out_dim, false, "fused outputs", f.out));
  else f.out = FeatRecipe{};
  f.val_mask = io->d_val_mask;
  return MGN_OK;
}

}  // namespace

cudaError_t build_features(const FeatRecipe& R, int64_t rows, float* y, cudaStream_t st) {
  if (rows == 0 || R.F == 0) return cudaSuccess;
  ProfScope ps(TAG_NORM, st);
  build_features_kernel<<<nblk(rows * R.F), 256, 0, st>>>(R, rows, y);
  return cudaGetLastError();
}
cudaError_t finish_output(const FeatRecipe& R, const float* val_mask, int64_t rows, int od, float* out, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  ProfScope ps(TAG_NORM, st);
  output_map_kernel<<<nblk(rows * od), 256, 0, st>>>(R, val_mask, out, rows, od, 0, out);
  return cudaGetLastError();
}
cudaError_t prepare_dout(const FeatRecipe& R, const float* val_mask, const float* dout, int64_t rows, int od, float* y,
                         cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  ProfScope ps(TAG_NORM, st);
  output_map_kernel<<<nblk(rows * od), 256, 0, st>>>(R, val_mask, dout, rows, od, 1, y);
  return cudaGetLastError();
}
cudaError_t finish_dx(const FeatRecipe& R, int64_t rows, float* dx, cudaStream_t st) {
  if (rows == 0 || R.F == 0) return cudaSuccess;
  ProfScope ps(TAG_NORM, st);
  finish_dx_kernel<<<nblk(rows * R.F), 256, 0, st>>>(R, rows, dx);
  return cudaGetLastError();
}

}  // namespace mgn

using namespace mgn;

extern "C" {

int32_t mgn_forward_fused(const mgn_model* m, const mgn_graph* g, const float* d_params, const mgn_fused_io* io,
                          float* d_out, void* d_workspace, size_t workspace_bytes, int32_t training, void* stream) {
  MGN_REQUIRE(m && g && d_params && d_out && d_workspace, "forward_fused: null argument");
  FusedIo f;
  MGN_TRY(to_io(m, g, io, f));
  return forward(m, g, d_params, nullptr, nullptr, d_out, d_workspace, workspace_bytes, training != 0,
                 static_cast<cudaStream_t>(stream), &f);
}

int32_t mgn_backward_fused(const mgn_model* m, const mgn_graph* g, const float* d_params, const mgn_fused_io* io,
                           const float* d_dout, float* d_dparams, float* d_dx, void* d_workspace, size_t workspace_bytes,
                           void* stream) {
  MGN_REQUIRE(m && g && d_params && d_dout && d_dparams && d_workspace, "backward_fused: null argument");
  FusedIo f;
  MGN_TRY(to_io(m, g, io, f));
  return backward(m, g, d_params, nullptr, nullptr, d_dout, d_dparams, d_dx, d_workspace, workspace_bytes,
                  static_cast<cudaStream_t>(stream), nullptr, &f);
}

int32_t mgn_norm_online_update_multi(const mgn_norm_update* h_jobs, int32_t n_jobs, void* stream) {
  MGN_REQUIRE(n_jobs >= 0 && n_jobs <= kMultiMax && (h_jobs || n_jobs == 0), "norm_online_update_multi: 0..8 jobs");
  if (n_jobs == 0) return MGN_OK;
  NormJobs jobs{};
  jobs.n = n_jobs;
  for (int i = 0; i < n_jobs; ++i) {
    const mgn_norm_update& u = h_jobs[i];
    MGN_REQUIRE(u.d_x && u.d_state && u.rows >= 0 && u.features > 0 && u.features <= kMaxFeat && u.col >= 0 &&
                    u.ld >= u.col + u.features,
                "norm_online_update_multi: bad job");
    jobs.j[i] = {u.d_x, u.rows, u.ld, u.col, u.features, u.d_state, u.max_acc};
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* scratch = nullptr;
  MGN_CUDA_TRY(stream_scratch(st, SCRATCH_NORM_MULTI, sizeof(float) * kMultiMax * kMultiBlocks * 2 * kMaxFeat,
                              reinterpret_cast<void**>(&scratch)));
  { ProfScope ps(TAG_NORM, st);
    norm_multi_partial_kernel<<<dim3(kMultiBlocks, n_jobs), 256, 0, st>>>(jobs, scratch); }
  { ProfScope ps(TAG_NORM, st);
    norm_multi_finish_kernel<<<n_jobs, 128, 0, st>>>(jobs, scratch); }
  MGN_CUDA_TRY(cudaGetLastError());
  return MGN_OK;
}

}  // extern "C"
