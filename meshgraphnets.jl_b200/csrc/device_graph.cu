// create_base_graph on the device (src/graph.jl:25-55): one-hot node types, triangles -> unique
// two-way edges in first-occurrence order, the 0 -> 1 based shift and the edge features [rel ; ||rel||], with the
// inputs already resident in HBM.  Integer results are BIT-EXACT equal to the host path of abi.cu (and to the oracle):
//
//   triangles_to_edges   raw edge e = side * C + face (the order of `[f0f1 ... ; f1f2 ... ; f2f0 ...]`), key = (max, min).
//                        An open-addressing hash set keyed by the 64-bit pair records, with atomicMin, the SMALLEST raw
//                        index of every key; raw edge e survives iff it is that minimum (= first occurrence); an
//                        exclusive scan of the survivor flags gives its position.  The result does not depend on the
//                        order in which threads reach the table: deterministic.
//   edge features        rel in fp32, ||rel|| accumulated in double and rounded once (LinearAlgebra.norm); the products
//                        of two floats are exact in double, so a contracted fma gives the same bits as the host loop.
#include "common.cuh"

namespace mgn {
namespace {

constexpr unsigned long long kEmpty = ~0ull;

__device__ __forceinline__ uint32_t hash64(unsigned long long k) {  // splitmix64 finaliser
  k ^= k >> 30;
  k *= 0xbf58476d1ce4e5b9ull;
  k ^= k >> 27;
  k *= 0x94d049bb133111ebull;
  k ^= k >> 31;
  return (uint32_t)k;
}

__device__ __forceinline__ unsigned long long raw_edge_key(const int32_t* __restrict__ cells, int64_t C, int64_t e) {
  const int side = (int)(e / C);
  const int64_t face = e - (int64_t)side * C;
  const int32_t a = cells[face * 3 + side], b = cells[face * 3 + (side == 2 ? 0 : side + 1)];
  const int32_t mx = max(a, b), mn = min(a, b);
  return ((unsigned long long)(uint32_t)mx << 32) | (uint32_t)mn;
}

__global__ void hash_insert_kernel(const int32_t* __restrict__ cells, int64_t C, unsigned long long* __restrict__ keys,
                                   int32_t* __restrict__ first, uint32_t mask) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * C) return;
  const unsigned long long key = raw_edge_key(cells, C, e);
  uint32_t slot = hash64(key) & mask;
  while (true) {
    const unsigned long long old = atomicCAS(&keys[slot], kEmpty, key);
    if (old == kEmpty || old == key) {
      atomicMin(&first[slot], (int32_t)e);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__global__ void hash_flag_kernel(const int32_t* __restrict__ cells, int64_t C, const unsigned long long* __restrict__ keys,
                                 const int32_t* __restrict__ first, uint32_t mask, int32_t* __restrict__ flags) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * C) return;
  const unsigned long long key = raw_edge_key(cells, C, e);
  uint32_t slot = hash64(key) & mask;
  while (keys[slot] != key) slot = (slot + 1) & mask;
  flags[e] = first[slot] == (int32_t)e ? 1 : 0;
}

// Block-local inclusive scan of 1024-element chunks + per-chunk totals; a second pass adds the chunk offsets.
__global__ void __launch_bounds__(1024) scan_chunks_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ excl,
                                                           int32_t* __restrict__ chunk_sum) {
  __shared__ int32_t warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  const int32_t v = i < n ? in[i] : 0;
  int32_t s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o) s += t;
  }
  if (lane == 31) warp_sums[warp] = s;
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  const int32_t incl = (warp ? warp_sums[warp - 1] : 0) + s;
  if (i < n) excl[i] = incl - v;
  if (threadIdx.x == 1023) chunk_sum[blockIdx.x] = incl;
}

// One block: exclusive scan of the chunk totals in place (n_chunks is small: 3C / 1024), total -> chunk_sum[n_chunks].
__global__ void __launch_bounds__(1024) scan_totals_kernel(int32_t* __restrict__ chunk_sum, int64_t n_chunks) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n_chunks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int32_t v = i < n_chunks ? chunk_sum[i] : 0;
    int32_t s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int32_t prefix = carry + (warp ? warp_sums[warp - 1] : 0) + s - v;
    if (i < n_chunks) chunk_sum[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) chunk_sum[n_chunks] = carry;
}

__global__ void emit_edges_kernel(const int32_t* __restrict__ cells, int64_t C, const int32_t* __restrict__ flags,
                                  const int32_t* __restrict__ excl, const int32_t* __restrict__ chunk_off, int64_t n_chunks,
                                  int32_t* __restrict__ senders, int32_t* __restrict__ receivers) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * C || !flags[e]) return;
  const int64_t U = chunk_off[n_chunks];
  const int64_t p = (int64_t)chunk_off[e >> 10] + excl[e];
  const unsigned long long key = raw_edge_key(cells, C, e);
  const int32_t mx = (int32_t)(key >> 32), mn = (int32_t)(key & 0xffffffffu);
  senders[p] = mx;
  senders[U + p] = mn;
  receivers[p] = mn;
  receivers[U + p] = mx;
}

__global__ void parse_edges_kernel(const int32_t* __restrict__ edges, int64_t n, int32_t* __restrict__ senders,
                                   int32_t* __restrict__ receivers) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t s = edges[2 * i], r = edges[2 * i + 1];
  senders[i] = s;
  senders[n + i] = r;
  receivers[i] = r;
  receivers[n + i] = s;
}

__global__ void any_zero_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n,
                                int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (a[i] == 0 || b[i] == 0)) *flag = 1;  // benign race: every writer stores 1
}
__global__ void shift_if_kernel(int32_t* __restrict__ a, int32_t* __restrict__ b, int64_t n, const int32_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && *flag) {
    a[i] += 1;
    b[i] += 1;
  }
}

__global__ void one_hot_kernel(const int32_t* __restrict__ v, int64_t n, int depth, int offset, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * depth) return;
  const int64_t row = i / depth;
  const int j = (int)(i - row * depth);
  out[i] = ((int64_t)v[row] + offset - 1 == j) ? 1.0f : 0.0f;
}

__global__ void edge_features_kernel(const float* __restrict__ pos, int64_t N, int dim, const int32_t* __restrict__ senders,
                                     const int32_t* __restrict__ receivers, int64_t E, int base, float* __restrict__ out,
                                     int32_t* __restrict__ err) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s = (int64_t)senders[e] - base, r = (int64_t)receivers[e] - base;
  if (s < 0 || s >= N || r < 0 || r >= N) {
    *err = 1;
    return;
  }
  double acc = 0.0;
  for (int d = 0; d < dim; ++d) {
    const float rel = __fsub_rn(pos[s * dim + d], pos[r * dim + d]);
    out[e * (dim + 1) + d] = rel;
    acc += (double)rel * (double)rel;
  }
  out[e * (dim + 1) + dim] = (float)sqrt(acc);
}

inline unsigned nblk(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

struct DevBuf {  // scratch of one call (graph construction may allocate and synchronise, like mgn_graph_create)
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 4); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

}  // namespace
}  // namespace mgn

using namespace mgn;

extern "C" {

int32_t mgn_one_hot_device(const int32_t* d_v, int64_t n, int32_t depth, int32_t offset, float* d_out, void* stream) {
  MGN_REQUIRE(n >= 0 && depth > 0, "one_hot_device: bad sizes");
  MGN_REQUIRE((d_v && d_out) || n == 0, "one_hot_device: null pointer");
  if (n == 0) return MGN_OK;
  one_hot_kernel<<<nblk(n * depth), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_v, n, depth, offset, d_out);
  MGN_CUDA_TRY(cudaGetLastError());
  return MGN_OK;
}

int32_t mgn_triangles_to_edges_device(const int32_t* d_cells, int64_t n_cells, int32_t* d_senders, int32_t* d_receivers,
                                      int64_t* h_n_edges, void* stream) {
  MGN_REQUIRE(n_cells >= 0 && h_n_edges, "triangles_to_edges_device: bad arguments");
  MGN_REQUIRE((d_cells && d_senders && d_receivers) || n_cells == 0, "triangles_to_edges_device: null pointer");
  MGN_REQUIRE(3 * n_cells < ((int64_t)1 << 30), "triangles_to_edges_device: too many cells for Int32 edge ids");
  *h_n_edges = 0;
  if (n_cells == 0) return MGN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t R = 3 * n_cells, n_chunks = (R + 1023) / 1024;
  uint32_t cap = 1024;
  while ((int64_t)cap < 2 * R) cap <<= 1;
  DevBuf keys, first, flags, excl, chunks;
  MGN_CUDA_TRY(keys.alloc(sizeof(unsigned long long) * cap));
  MGN_CUDA_TRY(first.alloc(sizeof(int32_t) * cap));
  MGN_CUDA_TRY(flags.alloc(sizeof(int32_t) * R));
  MGN_CUDA_TRY(excl.alloc(sizeof(int32_t) * R));
  MGN_CUDA_TRY(chunks.alloc(sizeof(int32_t) * (n_chunks + 1)));
  MGN_CUDA_TRY(cudaMemsetAsync(keys.p, 0xff, sizeof(unsigned long long) * cap, st));
  MGN_CUDA_TRY(cudaMemsetAsync(first.p, 0x7f, sizeof(int32_t) * cap, st));  // 0x7f7f7f7f > any raw edge id
  hash_insert_kernel<<<nblk(R), 256, 0, st>>>(d_cells, n_cells, keys.as<unsigned long long>(), first.as<int32_t>(), cap - 1);
  hash_flag_kernel<<<nblk(R), 256, 0, st>>>(d_cells, n_cells, keys.as<unsigned long long>(), first.as<int32_t>(), cap - 1,
                                            flags.as<int32_t>());
  scan_chunks_kernel<<<(unsigned)n_chunks, 1024, 0, st>>>(flags.as<int32_t>(), R, excl.as<int32_t>(), chunks.as<int32_t>());
  scan_totals_kernel<<<1, 1024, 0, st>>>(chunks.as<int32_t>(), n_chunks);
  emit_edges_kernel<<<nblk(R), 256, 0, st>>>(d_cells, n_cells, flags.as<int32_t>(), excl.as<int32_t>(), chunks.as<int32_t>(),
                                             n_chunks, d_senders, d_receivers);
  MGN_CUDA_TRY(cudaGetLastError());
  int32_t U = 0;
  MGN_CUDA_TRY(cudaMemcpyAsync(&U, chunks.as<int32_t>() + n_chunks, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MGN_CUDA_TRY(cudaStreamSynchronize(st));
  *h_n_edges = 2 * (int64_t)U;
  return MGN_OK;
}

int32_t mgn_parse_edges_device(const int32_t* d_edges, int64_t n_pairs, int32_t* d_senders, int32_t* d_receivers,
                               void* stream) {
  MGN_REQUIRE(n_pairs >= 0, "parse_edges_device: bad size");
  MGN_REQUIRE((d_edges && d_senders && d_receivers) || n_pairs == 0, "parse_edges_device: null pointer");
  if (n_pairs == 0) return MGN_OK;
  parse_edges_kernel<<<nblk(n_pairs), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_edges, n_pairs, d_senders, d_receivers);
  MGN_CUDA_TRY(cudaGetLastError());
  return MGN_OK;
}

int32_t mgn_shift_one_based_device(int32_t* d_senders, int32_t* d_receivers, int64_t n_edges, int32_t* h_shifted,
                                   void* stream) {
  MGN_REQUIRE(n_edges >= 0, "shift_one_based_device: bad size");
  MGN_REQUIRE((d_senders && d_receivers) || n_edges == 0, "shift_one_based_device: null pointer");
  if (h_shifted) *h_shifted = 0;
  if (n_edges == 0) return MGN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DevBuf flag;
  MGN_CUDA_TRY(flag.alloc(sizeof(int32_t)));
  MGN_CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(int32_t), st));
  any_zero_kernel<<<nblk(n_edges), 256, 0, st>>>(d_senders, d_receivers, n_edges, flag.as<int32_t>());
  shift_if_kernel<<<nblk(n_edges), 256, 0, st>>>(d_senders, d_receivers, n_edges, flag.as<int32_t>());
  MGN_CUDA_TRY(cudaGetLastError());
  int32_t f = 0;
  MGN_CUDA_TRY(cudaMemcpyAsync(&f, flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MGN_CUDA_TRY(cudaStreamSynchronize(st));  // the flag buffer is freed on return
  if (h_shifted) *h_shifted = f;
  return MGN_OK;
}

int32_t mgn_edge_features_device(const float* d_pos, int64_t n_nodes, int32_t dim, const int32_t* d_senders,
                                 const int32_t* d_receivers, int64_t n_edges, int32_t index_base, float* d_out,
                                 void* stream) {
  MGN_REQUIRE(n_nodes >= 0 && n_edges >= 0 && dim > 0, "edge_features_device: bad sizes");
  MGN_REQUIRE((d_pos && d_senders && d_receivers && d_out) || n_edges == 0, "edge_features_device: null pointer");
  if (n_edges == 0) return MGN_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DevBuf err;
  MGN_CUDA_TRY(err.alloc(sizeof(int32_t)));
  MGN_CUDA_TRY(cudaMemsetAsync(err.p, 0, sizeof(int32_t), st));
  edge_features_kernel<<<nblk(n_edges), 256, 0, st>>>(d_pos, n_nodes, dim, d_senders, d_receivers, n_edges, index_base,
                                                      d_out, err.as<int32_t>());
  MGN_CUDA_TRY(cudaGetLastError());
  int32_t bad = 0;
  MGN_CUDA_TRY(cudaMemcpyAsync(&bad, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MGN_CUDA_TRY(cudaStreamSynchronize(st));
  if (bad) return fail(MGN_ERR_INDEX, "edge_features_device: node id out of range");
  return MGN_OK;
}

}  // extern "C"
