// Kernels for the NeuralODE callers of the hot path (src/solve.jl ode_func_train / ode_step and the
// SolverStrategy losses of src/strategies.jl): explicit Runge-Kutta stage combinations, the inflow
// overwrite, the val_mask product, strided normaliser maps with their transposed Jacobians, and the
// shooting losses with a fixed summation order.  All of them are HBM-bound elementwise / reduction
// passes over [K*N][S] state matrices (S = sum of the target feature dims, 2 for CylinderFlow).
#include <algorithm>

#include "common.cuh"

namespace mgn {

static inline int blocks_for(int64_t n, int per) { return (int)std::max<int64_t>(1, (n + per - 1) / per); }

// y = x + sum_j c[j] * k[j]: every product and sum rounded separately (no FMA contraction), terms in
// ascending j - the same arithmetic as `x .+ (dt*a1) .* k1 .+ ...` evaluated left to right.
struct LinComb {
  const float* k[kOdeMaxTerms];
  float c[kOdeMaxTerms];
  int n;
};
__global__ void __launch_bounds__(256)
lincomb_kernel(const float* x, LinComb lc, int64_t n, float* y) {  // y may alias x or a term: no __restrict__
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = x ? x[i] : 0.f;
#pragma unroll
    for (int j = 0; j < kOdeMaxTerms; ++j)
      if (j < lc.n) acc = __fadd_rn(acc, __fmul_rn(lc.c[j], lc.k[j][i]));
    y[i] = acc;
  }
}

// y = mask ? (src ? src : 0) : x   (src/solve.jl:104-107,151; with src == nullptr its transposed Jacobian)
__global__ void __launch_bounds__(256)
overwrite_kernel(const float* __restrict__ x, const float* __restrict__ src, const uint8_t* __restrict__ mask,
                 int64_t n, float* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = mask[i] ? (src ? src[i] : 0.f) : x[i];
}

__global__ void __launch_bounds__(256)
mul_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a[i] * b[i];
}

// mode 0: (x - mean)/sd   1: x*sd + mean   2: x/sd (transposed Jacobian of 0)   3: x*sd (of 1)
__global__ void __launch_bounds__(256)
norm_apply_ld_kernel(const float* __restrict__ x, int ld_x, int col_x, int64_t rows, int F,
                     const float* __restrict__ state, float std_eps, int mode, float* __restrict__ y, int ld_y,
                     int col_y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  const int64_t r = i / F;
  const int f = (int)(i - r * F);
  const float cnt = fmaxf(state[2 * F], 1.f);
  const float mean = state[f] / cnt;
  float s = sqrtf(state[F + f] / cnt - mean * mean);
  if (!(s == s)) s = std_eps;  // NaN from a slightly negative variance
  const float sd = fmaxf(s, std_eps);
  const float v = x[r * ld_x + col_x + f];
  float o;
  switch (mode) {
    case 0: o = (v - mean) / sd; break;
    case 1: o = v * sd + mean; break;
    case 2: o = v / sd; break;
    default: o = v * sd; break;
  }
  y[r * ld_y + col_y + f] = o;
}

__global__ void __launch_bounds__(256)
affine_ld_kernel(const float* __restrict__ x, int ld_x, int col_x, int64_t rows, int F, float scale, float shift,
                 float* __restrict__ y, int ld_y, int col_y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  const int64_t r = i / F;
  const int f = (int)(i - r * F);
  y[r * ld_y + col_y + f] = x[r * ld_x + col_x + f] * scale + shift;
}

// Shooting losses.  Stage 1: block b reduces the contiguous element range [b*per, (b+1)*per) in a fixed order and
// writes the gradient; stage 2: one thread block adds the partials in block order.  Deterministic.
//   kind 0: sum (gt - pred)^2 * vm[i % period],  dpred  = -2 w (gt - pred) vm      (strategies.jl:253-286, :367-378)
//   kind 1: sum |a - b|,                          da    += w sign(a - b)            (strategies.jl:380-383)
constexpr int kLossBlocks = 256;
template <int KIND>
__global__ void __launch_bounds__(256)
shoot_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ vm,
                     int64_t period, int64_t n, float w, float* __restrict__ dpred, float* __restrict__ partial) {
  __shared__ float red[8];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * per, i1 = min(n, i0 + per);
  float s = 0.f;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += 256) {
    if (KIND == 0) {
      const float m = vm[i % period];
      const float d = gt[i] - pred[i];
      s = fmaf(d * d, m, s);
      dpred[i] = -2.f * w * d * m;
    } else {
      const float d = pred[i] - gt[i];
      s += fabsf(d);
      dpred[i] += d > 0.f ? w : (d < 0.f ? -w : 0.f);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k];
    partial[blockIdx.x] = t;
  }
}

__global__ void shoot_finish_kernel(const float* __restrict__ partial, int nblk, float w, int accumulate,
                                    float* __restrict__ loss) {
  float t = 0.f;
  for (int b = 0; b < nblk; ++b) t += partial[b];
  loss[0] = (accumulate ? loss[0] : 0.f) + w * t;
}

static cudaError_t loss_scratch(cudaStream_t st, float** out) {
  return stream_scratch(st, SCRATCH_LOSS, sizeof(float) * kLossBlocks, reinterpret_cast<void**>(out));
}

static int stream_grid(int64_t n) { return (int)std::min<int64_t>((int64_t)device_sm_count() * 8, blocks_for(n, 256)); }

cudaError_t ode_lincomb(const float* x, const float* const* k, const float* coef, int n_terms, int64_t n, float* y,
                        cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  LinComb lc;
  lc.n = n_terms;
  for (int j = 0; j < kOdeMaxTerms; ++j) {
    lc.k[j] = j < n_terms ? k[j] : nullptr;
    lc.c[j] = j < n_terms ? coef[j] : 0.f;
  }
  { ProfScope ps(TAG_SOLVER, st);
  lincomb_kernel<<<stream_grid(n), 256, 0, st>>>(x, lc, n, y); }
  return cudaGetLastError();
}

cudaError_t masked_overwrite(const float* x, const float* src, const uint8_t* mask, int64_t n, float* y,
                             cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  { ProfScope ps(TAG_SOLVER, st);
  overwrite_kernel<<<stream_grid(n), 256, 0, st>>>(x, src, mask, n, y); }
  return cudaGetLastError();
}

cudaError_t vec_mul(const float* a, const float* b, int64_t n, float* y, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  { ProfScope ps(TAG_SOLVER, st);
  mul_kernel<<<stream_grid(n), 256, 0, st>>>(a, b, n, y); }
  return cudaGetLastError();
}

cudaError_t norm_online_apply_ld(const float* x, int ld_x, int col_x, int64_t rows, int F, const float* state,
                                 float std_eps, int mode, float* y, int ld_y, int col_y, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  { ProfScope ps(TAG_NORM, st);
  norm_apply_ld_kernel<<<blocks_for(rows * F, 256), 256, 0, st>>>(x, ld_x, col_x, rows, F, state, std_eps, mode, y,
                                                                  ld_y, col_y); }
  return cudaGetLastError();
}

cudaError_t affine_apply_ld(const float* x, int ld_x, int col_x, int64_t rows, int F, float scale, float shift,
                            float* y, int ld_y, int col_y, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  { ProfScope ps(TAG_NORM, st);
  affine_ld_kernel<<<blocks_for(rows * F, 256), 256, 0, st>>>(x, ld_x, col_x, rows, F, scale, shift, y, ld_y, col_y); }
  return cudaGetLastError();
}

cudaError_t shooting_mse(const float* pred, const float* gt, const float* vm, int64_t period, int64_t n, float w,
                         int accumulate, float* loss, float* dpred, cudaStream_t st) {
  float* partial = nullptr;
  cudaError_t e = loss_scratch(st, &partial);
  if (e != cudaSuccess) return e;
  const int nblk = (int)std::min<int64_t>(kLossBlocks, blocks_for(n, 2048));
  { ProfScope ps(TAG_LOSS, st);
  shoot_partial_kernel<0><<<nblk, 256, 0, st>>>(pred, gt, vm, period, n, w, dpred, partial); }
  { ProfScope ps(TAG_LOSS, st);
  shoot_finish_kernel<<<1, 1, 0, st>>>(partial, nblk, w, accumulate, loss); }
  return cudaGetLastError();
}

cudaError_t shooting_continuity(const float* a, const float* b, int64_t n, float w, float* loss, float* da,
                                cudaStream_t st) {
  float* partial = nullptr;
  cudaError_t e = loss_scratch(st, &partial);
  if (e != cudaSuccess) return e;
  const int nblk = (int)std::min<int64_t>(kLossBlocks, blocks_for(n, 2048));
  { ProfScope ps(TAG_LOSS, st);
  shoot_partial_kernel<1><<<nblk, 256, 0, st>>>(a, b, nullptr, 1, n, w, da, partial); }
  { ProfScope ps(TAG_LOSS, st);
  shoot_finish_kernel<<<1, 1, 0, st>>>(partial, nblk, w, 1, loss); }
  return cudaGetLastError();
}

}  // namespace mgn
