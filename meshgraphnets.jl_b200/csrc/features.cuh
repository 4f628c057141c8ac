// build_graph (src/graph.jl:75-97) and the RHS post-processing of ode_step (src/solve.jl:205-218) as RECIPES that the
// model kernels evaluate on the fly: the node-feature matrix `vcat(n_norm[f](data[f]) ..., n_norm[
// "node_type"](onehot))` and `e_norm(edge_features)` are never materialised in the tensor-core mode - the first Dense
// layer of the encoders reads the raw fields and normalises while it stages its operand; `inverse_data(o_norm, out) .*
// val_mask` runs in the decoder's last epilogue; the pullbacks apply the transposed maps where they read / write.
#pragma once
#include "common.cuh"

namespace mgn {

constexpr int kMaxFeatSegs = 8;
constexpr int kMaxFeat = 64;  // raw features per entity (tensor-core path limit)

// One block of columns: `width` columns starting at column `col` of the row-major matrix x [rows][ld].
struct FeatSeg {
  const float* x;
  int ld, col, width;
  int kind;            // MGN_FEAT_AFFINE: y = x * scale + shift ; MGN_FEAT_ONLINE: y = (x - mean) / std from `state`
  float scale, shift;
  const float* state;  // NormaliserOnline state [sum | sum_sq | count | num_acc], 2 * width + 2 floats
  float eps;
};
struct FeatRecipe {
  FeatSeg s[kMaxFeatSegs];
  int n;   // segments
  int F;   // total width
};
struct FusedIo {        // what the pipelines receive; all-identity when the caller used mgn_forward / mgn_backward
  FeatRecipe node, edge;
  FeatRecipe out;       // out.n == 0: the network output is returned as is
  const float* val_mask;  // [N][out_dim] or nullptr
};

inline FeatRecipe identity_recipe(const float* x, int F) {
  FeatRecipe r{};
  r.n = 1;
  r.F = F;
  r.s[0] = {x, F, 0, F, MGN_FEAT_AFFINE, 1.f, 0.f, nullptr, 0.f};
  return r;
}

// Per-column form, built once per CTA in shared memory.
struct FeatCol {
  const float* x;
  int ld, col, kind;
  float a, b;  // AFFINE: scale, shift ; ONLINE: mean, std
};

__device__ __forceinline__ void online_mean_std(const float* state, int F, int f, float std_eps, float& mean, float& sd) {
  const float cnt = fmaxf(state[2 * F], 1.f);
  mean = state[f] / cnt;
  const float var = state[F + f] / cnt - mean * mean;
  float s = sqrtf(var);
  if (!(s == s)) s = std_eps;  // NaN from a slightly negative variance
  sd = fmaxf(s, std_eps);
}

// Threads tid, tid + nthreads, ... each fill one column of the table (tab has R.F entries).
__device__ __forceinline__ void feat_table(const FeatRecipe& R, FeatCol* tab, int tid, int nthreads) {
  for (int f = tid; f < R.F; f += nthreads) {
    int k = 0, f0 = 0;
    while (k < R.n - 1 && f >= f0 + R.s[k].width) {
      f0 += R.s[k].width;
      ++k;
    }
    const FeatSeg& s = R.s[k];
    FeatCol c;
    c.x = s.x;
    c.ld = s.ld;
    c.col = s.col + (f - f0);
    c.kind = s.kind;
    if (s.kind == MGN_FEAT_ONLINE) online_mean_std(s.state, s.width, f - f0, s.eps, c.a, c.b);
    else {
      c.a = s.scale;
      c.b = s.shift;
    }
    tab[f] = c;
  }
}
// forward map of column c at `row` - the very expressions of norm_apply_kernel / affine_kernel
__device__ __forceinline__ float feat_eval(const FeatCol& c, int64_t row) {
  const float v = c.x[row * c.ld + c.col];
  return c.kind == MGN_FEAT_ONLINE ? (v - c.a) / c.b : v * c.a + c.b;
}
// transposed Jacobian of the forward map (dy -> dx): dy / std or dy * scale
__device__ __forceinline__ float feat_vjp(const FeatCol& c, float dy) {
  return c.kind == MGN_FEAT_ONLINE ? dy / c.b : dy * c.a;
}
// inverse map on an output column (inverse_data): y * std + mean, or the caller's inverse affine
__device__ __forceinline__ float out_eval(const FeatCol& c, float v) {
  return c.kind == MGN_FEAT_ONLINE ? v * c.b + c.a : v * c.a + c.b;
}
__device__ __forceinline__ float out_vjp(const FeatCol& c, float dy) {
  return c.kind == MGN_FEAT_ONLINE ? dy * c.b : dy * c.a;
}

// fp32 mode: the recipes are materialised by ONE launch each (features.cu)
cudaError_t build_features(const FeatRecipe& R, int64_t rows, float* y, cudaStream_t st);
cudaError_t finish_output(const FeatRecipe& R, const float* val_mask, int64_t rows, int od, float* out, cudaStream_t st);
cudaError_t prepare_dout(const FeatRecipe& R, const float* val_mask, const float* dout, int64_t rows, int od, float* y,
                         cudaStream_t st);
cudaError_t finish_dx(const FeatRecipe& R, int64_t rows, float* dx, cudaStream_t st);

}  // namespace mgn
