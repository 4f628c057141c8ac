// fp32 CUDA-core kernels of the Encode-Process-Decode path (MGN_COMPUTE_FP32, the parity mode)
// plus the memory-bound kernels every mode shares (segmented sum, LayerNorm backward, gradient
// gather, loss, Adam, normalisers).  sm_100a only.
//
// Reference semantics (MeshGraphNets.jl v0.4.1 call sites; GraphNetCore.jl / Lux 0.5 / NNlib):
//   Dense        y = W x + b, relu on hidden layers              SURVEY 8 a14
//   LayerNorm    biased variance over the feature dim, eps       SURVEY 8 a14
//   scatter(+)   sequential sum in ascending edge id             SURVEY 8 a11 (here: CSR segment)
//   step! loss   mean(sum_rows((target-out)^2)[mask])            src/strategies.jl:421
//   Adam         Optimisers.update                               src/MeshGraphNets.jl:374-378
#include <algorithm>

#include "common.cuh"
#include "features.cuh"

namespace mgn {

namespace {

constexpr int BM = 64;   // rows per block
constexpr int BN = 128;  // output columns per block
constexpr int BK = 16;   // reduction chunk
constexpr int NT = 256;  // threads: 16 (cols) x 16 (rows)

struct SegDev {
  const float* base;
  const int32_t* idx;
  int width;
  int ld;
};
struct OperandDev {
  SegDev s[3];
  int nseg;
  int K;
  int chunk_aligned;  // all widths % 16 == 0, ld % 4 == 0, 16B aligned bases
};

struct LnDev {
  const float* scale;
  const float* bias;
  float eps;
  float* xhat;
  float* rstd;
  float* out;
  const float* resid_in;
  float* resid_out;
};

__device__ __forceinline__ float fetch_elem(const OperandDev& x, int64_t r, int k) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s < x.nseg) {
      if (k < x.s[s].width) {
        int64_t row = x.s[s].idx ? (int64_t)x.s[s].idx[r] : r;
        return x.s[s].base[row * x.s[s].ld + k];
      }
      k -= x.s[s].width;
    }
  }
  return 0.f;
}

// Pointer to 4 consecutive elements starting at column k of logical row r (aligned operands).
__device__ __forceinline__ const float* fetch_ptr(const OperandDev& x, int64_t r, int k) {
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (s < x.nseg) {
      if (k < x.s[s].width) {
        int64_t row = x.s[s].idx ? (int64_t)x.s[s].idx[r] : r;
        return x.s[s].base + row * x.s[s].ld + k;
      }
      k -= x.s[s].width;
    }
  }
  return nullptr;
}

// ---------------------------------------------------------------------------------------------
// Y = epilogue(X B).  B_TRANS == false: B[k][c] = W[k*n + c]  (forward, W is [K][n]).
//                     B_TRANS == true : B[k][c] = W[c*ldw + k] (dX = dZ W^T, W is [cols][K]).
// Epilogues: bias (+relu) ; bias -> LayerNorm (-> residual) ; relu mask.
// ---------------------------------------------------------------------------------------------
template <bool B_TRANS>
__global__ void __launch_bounds__(NT)
gemm_kernel(OperandDev x, int64_t M, const float* __restrict__ W, int ldw, int ncols,
            const float* __restrict__ bias, int relu, const float* __restrict__ mask_src,
            float* __restrict__ Y, int ldy, LnDev ln, int has_ln) {
  __shared__ __align__(16) float Xs[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int K = x.K;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool w_al16 = (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  const bool b_aligned = w_al16 && (B_TRANS ? ((ldw & 3) == 0 && (K & 15) == 0)
                                            : ((ldw & 3) == 0 && (ncols - col0) >= BN));
  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- X chunk [BM][BK] -> Xs[k][r]
    {
      const int r = tid >> 2, kq = (tid & 3) * 4;
      const int64_t gr = row0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < M) {
        if (x.chunk_aligned) {
          v = *reinterpret_cast<const float4*>(fetch_ptr(x, gr, k0 + kq));
        } else {
          float t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] = (k0 + kq + j < K) ? fetch_elem(x, gr, k0 + kq + j) : 0.f;
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      Xs[kq + 0][r] = v.x;
      Xs[kq + 1][r] = v.y;
      Xs[kq + 2][r] = v.z;
      Xs[kq + 3][r] = v.w;
    }
    // ---- B chunk [BK][BN]
    if (!B_TRANS) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int kk = (tid >> 5) + h * 8, c = (tid & 31) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + kk < K) {
          const float* src = W + (int64_t)(k0 + kk) * ldw + col0 + c;
          if (b_aligned) {
            v = *reinterpret_cast<const float4*>(src);
          } else {
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = (col0 + c + j < ncols) ? src[j] : 0.f;
            v = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
        *reinterpret_cast<float4*>(&Bs[kk][c]) = v;
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = (tid >> 2) + h * 64, kq = (tid & 3) * 4;
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        if (col0 + c < ncols) {
          const float* src = W + (int64_t)(col0 + c) * ldw + k0 + kq;
          if (b_aligned) {
            float4 v = *reinterpret_cast<const float4*>(src);
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = (k0 + kq + j < K) ? src[j] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[kq + j][c] = t[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue.  Thread owns rows ty*4+i and columns {tx*4+j, 64+tx*4+j}.
  int cidx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cidx[j] = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
  if (bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b = cidx[j] < ncols ? bias[cidx[j]] : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][j] += b;
    }
  }
  if (!has_ln) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t gr = row0 + ty * 4 + i;
      if (gr >= M) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (cidx[j] >= ncols) continue;
        float v = acc[i][j];
        if (relu) v = fmaxf(v, 0.f);
        if (mask_src) v = mask_src[gr * ldy + cidx[j]] > 0.f ? v : 0.f;
        Y[gr * ldy + cidx[j]] = v;
      }
    }
    return;
  }
  // LayerNorm over the ncols (<= BN) columns of each row: the 16 threads sharing a row are the
  // 16 consecutive lanes of a half warp.
  const float inv_n = 1.f / (float)ncols;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty * 4 + i;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += cidx[j] < ncols ? acc[i][j] : 0.f;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * inv_n;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = cidx[j] < ncols ? acc[i][j] - mu : 0.f;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = 1.f / sqrtf(q * inv_n + ln.eps);
    if (gr >= M) continue;
    if (ln.rstd && tx == 0) ln.rstd[gr] = rs;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (cidx[j] >= ncols) continue;
      const int64_t o = gr * ncols + cidx[j];
      const float xh = (acc[i][j] - mu) * rs;
      const float y = fmaf(xh, ln.scale[cidx[j]], ln.bias[cidx[j]]);
      if (ln.xhat) ln.xhat[o] = xh;
      if (ln.out) ln.out[o] = y;
      if (ln.resid_out) ln.resid_out[o] = ln.resid_in[o] + y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient: G[i][o] = sum_r X[r][i] dZ[r][o], i in [0, K]; row i == K is the bias
// gradient (X extended by a column of ones).  Rows are split over gridDim.y; every split writes
// its own partial (no atomics), reduce_partials_kernel sums them in a fixed order.
// ---------------------------------------------------------------------------------------------
constexpr int DW_BI = 64;
__global__ void __launch_bounds__(NT)
dw_kernel(OperandDev x, const float* __restrict__ dZ, int64_t M, int n, int64_t rows_per_split,
          float* __restrict__ partial) {
  __shared__ __align__(16) float Xs[BK][DW_BI + 4];
  __shared__ __align__(16) float Zs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int K = x.K;
  const int i0 = blockIdx.x * DW_BI;
  const int o0 = blockIdx.z * BN;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(M, r_begin + rows_per_split);
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const bool x_fast = x.chunk_aligned && (i0 + DW_BI <= K);
  const bool z_fast = ((n & 3) == 0) && (o0 + BN <= n) && (reinterpret_cast<uintptr_t>(dZ) & 15) == 0;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
    {  // X chunk: 16 rows x 64 cols, one float4 per thread
      const int rr = tid >> 4, c = (tid & 15) * 4;
      const int64_t gr = r0 + rr;
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      if (gr < r_end) {
        if (x_fast) {
          float4 v = *reinterpret_cast<const float4*>(fetch_ptr(x, gr, i0 + c));
          t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = i0 + c + j;
            t[j] = i < K ? fetch_elem(x, gr, i) : (i == K ? 1.f : 0.f);
          }
        }
      }
      *reinterpret_cast<float4*>(&Xs[rr][c]) = make_float4(t[0], t[1], t[2], t[3]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // dZ chunk: 16 rows x 128 cols
      const int rr = (tid >> 5) + h * 8, c = (tid & 31) * 4;
      const int64_t gr = r0 + rr;
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      if (gr < r_end) {
        const float* src = dZ + gr * n + o0 + c;
        if (z_fast) {
          float4 v = *reinterpret_cast<const float4*>(src);
          t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] = (o0 + c + j < n) ? src[j] : 0.f;
        }
      }
      *reinterpret_cast<float4*>(&Zs[rr][c]) = make_float4(t[0], t[1], t[2], t[3]);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Zs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Zs[kk][64 + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = partial + (int64_t)blockIdx.y * (int64_t)(K + 1) * n;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = i0 + ty * 4 + i;
    if (gi > K) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int go = o0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (go < n) dst[(int64_t)gi * n + go] = acc[i][j];
    }
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nsplit,
                                       int64_t count, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int k = 0; k < nsplit; ++k) s += partial[(int64_t)k * count + i];
  out[i] = s;
}

// ---------------------------------------------------------------------------------------------
// Segmented sum over CSR rows: one warp per node, each lane owns a strided set of columns.
// Replaces NNlib.scatter(+) (atomic adds on the GPU) with a deterministic sum in ascending
// original edge id - the CPU reference's order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ m, const int32_t* __restrict__ row_ptr, int64_t N,
                   int D, float* __restrict__ agg) {
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= N) return;
  const int lane = threadIdx.x & 31;
  const int b = row_ptr[v], e = row_ptr[v + 1];
  if ((D & 127) == 0) {
    for (int c = lane * 4; c < D; c += 128) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = b; j < e; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(m + (int64_t)j * D + c);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      *reinterpret_cast<float4*>(agg + v * D + c) = s;
    }
  } else {
    for (int c = lane; c < D; c += 32) {
      float s = 0.f;
      for (int j = b; j < e; ++j) s += m[(int64_t)j * D + c];
      agg[v * D + c] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  dy[r] = a[r] + b[bidx[r]] ; dz = rstd (dxh - mean(dxh) - xhat mean(dxh xhat))
// with dxh = dy * scale.  Column sums of dy (-> g_bias) and dy*xhat (-> g_scale) are produced as
// per-block partials and reduced in a fixed order.  One warp per row; D <= 128.
// ---------------------------------------------------------------------------------------------
constexpr int LN_ROWS_PER_BLOCK = 64;
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
              const int32_t* __restrict__ bidx, const float* __restrict__ xhat,
              const float* __restrict__ rstd, const float* __restrict__ scale, int64_t M, int D,
              float* __restrict__ dz, float* __restrict__ partial) {
  __shared__ float red[8][2][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * LN_ROWS_PER_BLOCK;
  float sc[4], gs[4], gb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane + 32 * j;
    sc[j] = c < D ? scale[c] : 0.f;
    gs[j] = 0.f;
    gb[j] = 0.f;
  }
  const float inv_d = 1.f / (float)D;
  for (int rr = warp; rr < LN_ROWS_PER_BLOCK; rr += 8) {
    const int64_t r = row0 + rr;
    if (r >= M) break;
    const int64_t br = b ? (bidx ? (int64_t)bidx[r] : r) : 0;
    float dy[4], xh[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      float v = 0.f, h = 0.f;
      if (c < D) {
        if (a) v += a[r * lda + c];
        if (b) v += b[br * ldb + c];
        h = xhat[r * D + c];
      }
      dy[j] = v;
      xh[j] = h;
      const float dxh = v * sc[j];
      s1 += dxh;
      s2 = fmaf(dxh, h, s2);
      gb[j] += v;
      gs[j] = fmaf(v, h, gs[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float rs = rstd[r];
    const float m1 = s1 * inv_d, m2 = s2 * inv_d;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      if (c < D) dz[r * D + c] = rs * (dy[j] * sc[j] - m1 - xh[j] * m2);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[warp][0][lane + 32 * j] = gs[j];
    red[warp][1][lane + 32 * j] = gb[j];
  }
  __syncthreads();
  const int t = threadIdx.x;  // 256 threads: [0,128) -> g_scale, [128,256) -> g_bias
  const int which = t >> 7, c = t & 127;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][which][c];
  partial[(int64_t)blockIdx.x * 256 + t] = s;
}

// out[0:D] = g_scale, out2[0:D] = g_bias from nblk partials of 256 floats.
__global__ void ln_reduce_kernel(const float* __restrict__ partial, int64_t nblk, int D,
                                 float* __restrict__ g_scale, float* __restrict__ g_bias) {
  const int t = threadIdx.x;
  const int which = t >> 7, c = t & 127;
  if (c >= D) return;
  float s = 0.f;
  for (int64_t k = 0; k < nblk; ++k) s += partial[k * 256 + t];
  (which ? g_bias : g_scale)[c] = s;
}

// ---------------------------------------------------------------------------------------------
// d_nf[v] = base[v] + add[v] + sum_{j in CSR row v} dxe[j][D:2D] + sum_{j in CSC row v} dxe[slot_j][0:D]
// (gradient of the two gathers nf[:, receivers], nf[:, senders] - deterministic, no atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
node_grad_gather_kernel(const float* __restrict__ base, const float* __restrict__ add, int ld_add,
                        const float* __restrict__ dxe, const int32_t* __restrict__ row_ptr,
                        const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ csc_slot,
                        int64_t N, int D, float* __restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= N) return;
  const int lane = threadIdx.x & 31;
  const int rb = row_ptr[v], re = row_ptr[v + 1];
  const int cb = col_ptr[v], ce = col_ptr[v + 1];
  const int ldx = 3 * D;
  for (int c = lane; c < D; c += 32) {
    float s = base ? base[v * D + c] : 0.f;
    if (add) s += add[v * ld_add + c];
    for (int j = rb; j < re; ++j) s += dxe[(int64_t)j * ldx + D + c];
    for (int j = cb; j < ce; ++j) s += dxe[(int64_t)csc_slot[j] * ldx + c];
    out[v * D + c] = s;
  }
}

// out = a + b[:, col_b : col_b + D] (+ g[idx[r], col_g : col_g + D] when g != nullptr: the gathered third term is the
// aggregation adjoint on the residual path of aggregate_post_residual = 1)
__global__ void add_cols_kernel(const float* __restrict__ a, const float* __restrict__ b, int ldb,
                                int col_b, int64_t M, int D, float* __restrict__ out, const float* __restrict__ g,
                                int ldg, int col_g, const int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  const int64_t r = i / D;
  const int c = (int)(i - r * D);
  float v = (a ? a[i] : 0.f) + b[r * ldb + col_b + c];
  if (g) v += g[(int64_t)idx[r] * ldg + col_g + c];
  out[i] = v;
}

// ---------------------------------------------------------------------------------------------
// Loss: single block, fixed summation order (deterministic scalar).
// ---------------------------------------------------------------------------------------------
__global__ void zero_kernel(float* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

__global__ void __launch_bounds__(1024)
loss_kernel(const float* __restrict__ out, const float* __restrict__ target, int out_dim,
            const int32_t* __restrict__ mask, int64_t n_mask, int base, float* __restrict__ loss,
            float* __restrict__ dout) {
  __shared__ float red[32];
  float s = 0.f;
  const float inv = 1.f / (float)n_mask;
  for (int64_t k = threadIdx.x; k < n_mask; k += blockDim.x) {
    const int64_t v = (int64_t)mask[k] - base;
    for (int c = 0; c < out_dim; ++c) {
      const float d = out[v * out_dim + c] - target[v * out_dim + c];
      s = fmaf(d, d, s);
      dout[v * out_dim + c] = 2.f * d * inv;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) loss[0] = t * inv;
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                            float b1, float b2, float eps, float c1, float c2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] = p[i] - (mi / c1) / (sqrtf(vi / c2) + eps) * lr;  // Optimisers.apply! operation order
}

// ---------------------------------------------------------------------------------------------
// Normalisers
// ---------------------------------------------------------------------------------------------
// state = [sum[F] | sumsq[F] | count | num_acc].  Two stages, both with a fixed summation order (deterministic):
// every block reduces a fixed, contiguous range of rows to 2F partial sums; one block then adds the partials in
// block order and updates the state.  F <= 64.
constexpr int kNormBlocks = 128;
constexpr int kNormMaxF = 64;
__global__ void __launch_bounds__(256)
norm_partial_kernel(const float* __restrict__ x, int64_t rows, int F, const float* __restrict__ state, float max_acc,
                    float* __restrict__ partial) {
  __shared__ float red[8][2 * kNormMaxF];
  if (state[2 * F + 1] >= max_acc) return;  // uniform
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // thread t walks the flattened [rows][F] range with stride 256: element e belongs to feature e % F
  for (int f0 = 0; f0 < F; f0 += 1) {
    float s = 0.f, q = 0.f;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += 256) {
      const float v = x[r * F + f0];
      s += v;
      q = fmaf(v, v, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      red[warp][f0] = s;
      red[warp][F + f0] = q;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * F) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[(int64_t)blockIdx.x * 2 * kNormMaxF + threadIdx.x] = t;
  }
}

__global__ void __launch_bounds__(128)
norm_finish_kernel(const float* __restrict__ partial, int nblk, int64_t rows, int F, float* __restrict__ state,
                   float max_acc) {
  if (state[2 * F + 1] >= max_acc) return;
  if (threadIdx.x < 2 * F) {
    float t = 0.f;
    for (int b = 0; b < nblk; ++b) t += partial[(int64_t)b * 2 * kNormMaxF + threadIdx.x];
    state[threadIdx.x] += t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    state[2 * F] += (float)rows;
    state[2 * F + 1] += 1.f;
  }
}

__global__ void norm_apply_kernel(const float* __restrict__ x, int64_t rows, int F,
                                  const float* __restrict__ state, float std_eps, int inverse,
                                  float* __restrict__ y, int ld_y, int col_y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  const int64_t r = i / F;
  const int f = (int)(i - r * F);
  float mean, sd;
  online_mean_std(state, F, f, std_eps, mean, sd);
  const float v = x[i];
  y[r * ld_y + col_y + f] = inverse ? v * sd + mean : (v - mean) / sd;
}

__global__ void affine_kernel(const float* __restrict__ x, int64_t rows, int F, float scale,
                              float shift, float* __restrict__ y, int ld_y, int col_y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  const int64_t r = i / F;
  const int f = (int)(i - r * F);
  y[r * ld_y + col_y + f] = x[i] * scale + shift;
}

struct AdamState {
  long long step;
  float c1, c2;
};
__global__ void adam_tick_kernel(AdamState* s, float b1, float b2) {
  const long long t = s->step + 1;
  s->step = t;
  s->c1 = (float)(1.0 - pow((double)b1, (double)t));
  s->c2 = (float)(1.0 - pow((double)b2, (double)t));
}
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                float b1, float b2, float eps, const AdamState* __restrict__ s) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c1 = s->c1, c2 = s->c2;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] = p[i] - (mi / c1) / (sqrtf(vi / c2) + eps) * lr;
}

// One warp per row; rows are copied as 32-bit words (row bytes are a multiple of 4).
__global__ void __launch_bounds__(256)
rows_op_kernel(uint32_t* __restrict__ base, int row_words, const int32_t* __restrict__ rows, int64_t n_rows,
               int64_t N, uint32_t* __restrict__ buf, int op) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n_rows) return;
  const int64_t r = rows[i];
  if (r < 0 || r >= N) return;  // ids are validated on the host side of the exchange plan
  uint32_t* t = base + r * row_words;
  uint32_t* b = buf + i * row_words;
  for (int c = threadIdx.x & 31; c < row_words; c += 32) {
    if (op == MGN_ROWS_PACK) {
      b[c] = t[c];
    } else if (op == MGN_ROWS_UNPACK) {
      t[c] = b[c];
    } else if (op == MGN_ROWS_ADD) {
      t[c] = __float_as_uint(__uint_as_float(t[c]) + __uint_as_float(b[c]));
    } else {
      b[c] = t[c];
      t[c] = 0u;
    }
  }
}

OperandDev to_dev(const Operand& x) {
  OperandDev d{};
  d.nseg = x.nseg;
  d.K = 0;
  d.chunk_aligned = 1;
  for (int i = 0; i < x.nseg; ++i) {
    d.s[i] = {x.s[i].base, x.s[i].idx, x.s[i].width, x.s[i].ld};
    d.K += x.s[i].width;
    if (x.s[i].width % 16 || x.s[i].ld % 4 || (reinterpret_cast<uintptr_t>(x.s[i].base) & 15))
      d.chunk_aligned = 0;
  }
  return d;
}

inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
cudaError_t dense_forward(const Operand& x, int64_t M, const float* W, const float* bias, int n,
                          bool relu, float* Y, const LnEpilogue* ln, cudaStream_t st) {
  if (M == 0) return cudaSuccess;
  OperandDev xd = to_dev(x);
  LnDev l{};
  if (ln) {
    if (n > BN) return cudaErrorInvalidValue;
    l = {ln->ln_scale, ln->ln_bias, ln->eps, ln->xhat, ln->rstd, ln->out, ln->resid_in, ln->resid_out};
  }
  dim3 grid(blocks_for(M, BM), blocks_for(n, BN));
  { ProfScope ps(TAG_SIMT_GEMM_FWD, st);
  gemm_kernel<false><<<grid, NT, 0, st>>>(xd, M, W, n, n, bias, relu ? 1 : 0, nullptr, Y, n, l,
                                          ln ? 1 : 0); }
  return cudaGetLastError();
}

cudaError_t dense_backward_dx(const float* dZ, int64_t M, int n, const float* W, int K,
                              const float* relu_src, float* dX, cudaStream_t st) {
  if (M == 0) return cudaSuccess;
  Operand x{};
  x.nseg = 1;
  x.s[0] = {dZ, nullptr, n, n};
  OperandDev xd = to_dev(x);
  LnDev l{};
  dim3 grid(blocks_for(M, BM), blocks_for(K, BN));
  // reduction dim = n, output columns = K, B[k][c] = W[c*n + k]
  { ProfScope ps(TAG_SIMT_GEMM_DX, st);
  gemm_kernel<true><<<grid, NT, 0, st>>>(xd, M, W, n, K, nullptr, 0, relu_src, dX, K, l, 0); }
  return cudaGetLastError();
}

static int64_t dw_rows_per_split(int64_t M) {
  int64_t rps = (M + 63) / 64;
  if (rps < 512) rps = 512;
  return (rps + BK - 1) / BK * BK;
}
static int dw_nsplit(int64_t M) {
  const int64_t rps = dw_rows_per_split(M);
  return (int)((M + rps - 1) / rps);
}
size_t dw_partial_floats(int K, int n, int64_t M) {
  return (size_t)dw_nsplit(M) * (size_t)(K + 1) * (size_t)n;
}

cudaError_t dense_backward_dw(const Operand& x, const float* dZ, int64_t M, int n, float* partial,
                              float* g_w_and_b, cudaStream_t st) {
  OperandDev xd = to_dev(x);
  const int K = xd.K;
  const int64_t count = (int64_t)(K + 1) * n;
  if (M == 0) {
    zero_kernel<<<blocks_for(count, 256), 256, 0, st>>>(g_w_and_b, count);
    return cudaGetLastError();
  }
  const int nsplit = dw_nsplit(M);
  dim3 grid(blocks_for(K + 1, DW_BI), nsplit, blocks_for(n, BN));
  { ProfScope ps(TAG_SIMT_DW, st);
  dw_kernel<<<grid, NT, 0, st>>>(xd, dZ, M, n, dw_rows_per_split(M), partial); }
  { ProfScope ps(TAG_REDUCE_PARTIALS, st);
  reduce_partials_kernel<<<blocks_for(count, 256), 256, 0, st>>>(partial, nsplit, count, g_w_and_b); }
  return cudaGetLastError();
}

cudaError_t segment_sum(const float* m, const int32_t* row_ptr, int64_t N, int D, float* agg,
                        cudaStream_t st) {
  if (N == 0) return cudaSuccess;
  { ProfScope ps(TAG_SEGMENT_SUM, st);
  segment_sum_kernel<<<blocks_for(N, 8), 256, 0, st>>>(m, row_ptr, N, D, agg); }
  return cudaGetLastError();
}

size_t ln_partial_floats(int64_t M, int D) {
  (void)D;
  return (size_t)blocks_for(M > 0 ? M : 1, LN_ROWS_PER_BLOCK) * 256;
}

cudaError_t layernorm_backward(const float* a, int lda, const float* b, int ldb,
                               const int32_t* bidx, const float* xhat, const float* rstd,
                               const float* scale, int64_t M, int D, float* dz, float* partial,
                               float* g_scale, float* g_bias, cudaStream_t st) {
  if (D > 128) return cudaErrorInvalidValue;
  const int64_t nblk = M > 0 ? blocks_for(M, LN_ROWS_PER_BLOCK) : 0;
  if (nblk) {
    ProfScope ps(TAG_LN_BWD, st);
    ln_bwd_kernel<<<(unsigned)nblk, 256, 0, st>>>(a, lda, b, ldb, bidx, xhat, rstd, scale, M, D, dz,
                                                  partial);
  }
  { ProfScope ps(TAG_LN_REDUCE, st);
  ln_reduce_kernel<<<1, 256, 0, st>>>(partial, nblk, D, g_scale, g_bias); }
  return cudaGetLastError();
}

cudaError_t node_grad_gather(const float* base, const float* add, int ld_add, const float* dxe,
                             const int32_t* row_ptr, const int32_t* col_ptr,
                             const int32_t* csc_slot, int64_t N, int D, float* out,
                             cudaStream_t st) {
  if (N == 0) return cudaSuccess;
  { ProfScope ps(TAG_NODE_GRAD_GATHER, st);
  node_grad_gather_kernel<<<blocks_for(N, 8), 256, 0, st>>>(base, add, ld_add, dxe, row_ptr, col_ptr,
                                                            csc_slot, N, D, out); }
  return cudaGetLastError();
}

cudaError_t add_cols(const float* a, const float* b, int ldb, int col_b, int64_t M, int D,
                     float* out, cudaStream_t st, const float* g, int ldg, int col_g, const int32_t* idx) {
  if (M == 0) return cudaSuccess;
  { ProfScope ps(TAG_ADD_COLS, st);
  add_cols_kernel<<<blocks_for(M * D, 256), 256, 0, st>>>(a, b, ldb, col_b, M, D, out, g, ldg, col_g, idx); }
  return cudaGetLastError();
}

cudaError_t loss_mse_masked(const float* out, const float* target, int64_t N, int out_dim,
                            const int32_t* mask, int64_t n_mask, int base, float* loss,
                            float* dout, cudaStream_t st) {
  ProfScope ps(TAG_LOSS, st);
  zero_kernel<<<blocks_for(N * out_dim, 256), 256, 0, st>>>(dout, N * out_dim);
  loss_kernel<<<1, 1024, 0, st>>>(out, target, out_dim, mask, n_mask, base, loss, dout);
  return cudaGetLastError();
}

cudaError_t rows_op(void* base, int elem_bytes, int row_elems, const int32_t* rows, int64_t n_rows, int64_t N,
                    void* buf, int op, cudaStream_t st) {
  if (n_rows == 0) return cudaSuccess;
  if ((row_elems * elem_bytes) % 4 != 0 || op < MGN_ROWS_PACK || op > MGN_ROWS_PACK_ZERO) return cudaErrorInvalidValue;
  if (op == MGN_ROWS_ADD && elem_bytes != 4) return cudaErrorInvalidValue;
  ProfScope ps(TAG_TC_MISC, st);
  rows_op_kernel<<<blocks_for(n_rows, 8), 256, 0, st>>>(static_cast<uint32_t*>(base), row_elems * elem_bytes / 4, rows,
                                                        n_rows, N, static_cast<uint32_t*>(buf), op);
  return cudaGetLastError();
}

cudaError_t adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1,
                      float b2, float eps, int64_t t, cudaStream_t st) {
  // bias corrections in double, rounded once (Optimisers keeps beta^t as Float32 products; the
  // difference is below 1 ulp of the step for t < 1e6)
  const float c1 = (float)(1.0 - pow((double)b1, (double)t));
  const float c2 = (float)(1.0 - pow((double)b2, (double)t));
  { ProfScope ps(TAG_ADAM, st);
  adam_kernel<<<blocks_for(n, 256), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, c1, c2); }
  return cudaGetLastError();
}

cudaError_t adam_step_device(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                             float b1, float b2, float eps, void* state16, cudaStream_t st) {
  ProfScope ps(TAG_ADAM, st);
  AdamState* s = static_cast<AdamState*>(state16);
  adam_tick_kernel<<<1, 1, 0, st>>>(s, b1, b2);
  adam_dev_kernel<<<blocks_for(n, 256), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, s);
  return cudaGetLastError();
}

cudaError_t adam_tick(void* state16, float b1, float b2, cudaStream_t st) {
  ProfScope ps(TAG_ADAM, st);
  adam_tick_kernel<<<1, 1, 0, st>>>(static_cast<AdamState*>(state16), b1, b2);
  return cudaGetLastError();
}

cudaError_t adam_apply_range(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
                             float eps, const void* state16, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  ProfScope ps(TAG_ADAM, st);
  adam_dev_kernel<<<blocks_for(n, 256), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, static_cast<const AdamState*>(state16));
  return cudaGetLastError();
}

cudaError_t norm_online_update(const float* x, int64_t rows, int F, float* state, float max_acc,
                               cudaStream_t st) {
  if (F > kNormMaxF) return cudaErrorInvalidValue;
  // scratch for the block partials: private to (device, stream), see stream_scratch()
  float* scratch = nullptr;
  cudaError_t se = stream_scratch(st, SCRATCH_NORM, sizeof(float) * kNormBlocks * 2 * kNormMaxF,
                                  reinterpret_cast<void**>(&scratch));
  if (se != cudaSuccess) return se;
  const int nblk = (int)std::min<int64_t>(kNormBlocks, std::max<int64_t>(1, (rows + 2047) / 2048));
  { ProfScope ps(TAG_NORM, st);
  norm_partial_kernel<<<nblk, 256, 0, st>>>(x, rows, F, state, max_acc, scratch); }
  { ProfScope ps(TAG_NORM, st);
  norm_finish_kernel<<<1, 128, 0, st>>>(scratch, nblk, rows, F, state, max_acc); }
  return cudaGetLastError();
}

cudaError_t norm_online_apply(const float* x, int64_t rows, int F, const float* state,
                              float std_eps, int inverse, float* y, int ld_y, int col_y,
                              cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  { ProfScope ps(TAG_NORM, st);
  norm_apply_kernel<<<blocks_for(rows * F, 256), 256, 0, st>>>(x, rows, F, state, std_eps, inverse, y,
                                                               ld_y, col_y); }
  return cudaGetLastError();
}

cudaError_t affine_apply(const float* x, int64_t rows, int F, float scale, float shift, float* y,
                         int ld_y, int col_y, cudaStream_t st) {
  if (rows == 0) return cudaSuccess;
  { ProfScope ps(TAG_NORM, st);
  affine_kernel<<<blocks_for(rows * F, 256), 256, 0, st>>>(x, rows, F, scale, shift, y, ld_y, col_y); }
  return cudaGetLastError();
}

}  // namespace mgn
