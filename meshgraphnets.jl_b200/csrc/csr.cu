// Device-side graph indexing (NEW - no reference counterpart): the stable sort of
// edge ids by receiver (CSR) and by sender (CSC).  Integer work, bit-exact against
// oracle.build_csr: inside a segment the order is the ascending original edge id, i.e. the
// summation order of the sequential CPU NNlib.scatter(+) behind GraphNetCore's aggregation.
//
// Algorithm: histogram -> exclusive scan -> unordered placement with an integer cursor ->
// per-segment sort of the edge ids.  The multiset of ids in a segment does not depend on the
// placement order, so sorting it makes the result deterministic and equal to a stable sort.
#include "common.cuh"

namespace mgn {
namespace {

__global__ void validate_hist_kernel(const int32_t* __restrict__ keys, int64_t E, int64_t N,
                                     int base, int32_t* __restrict__ counts,
                                     int32_t* __restrict__ err) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t k = (int64_t)keys[e] - base;
  if (k < 0 || k >= N) {
    atomicExch(err, 1);
    return;
  }
  atomicAdd(&counts[k], 1);
}

// Single-block exclusive scan (chunks of 1024 with a running carry); out has n+1 entries.
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out,
                      int32_t* __restrict__ max_out) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry;
  __shared__ int32_t mx[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t local_max = 0;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int32_t v = i < n ? in[i] : 0;
    local_max = max(local_max, v);
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
    if (i < n) out[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
  if (max_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if (lane == 0) mx[warp] = local_max;
    __syncthreads();
    if (threadIdx.x == 0) {
      int32_t m = 0;
      for (int w = 0; w < 32; ++w) m = max(m, mx[w]);
      *max_out = m;
    }
  }
}

__global__ void place_kernel(const int32_t* __restrict__ keys, int64_t E, int base,
                             const int32_t* __restrict__ ptr, int32_t* __restrict__ cursor,
                             int32_t* __restrict__ perm) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t k = keys[e] - base;
  const int32_t pos = ptr[k] + atomicAdd(&cursor[k], 1);
  perm[pos] = (int32_t)e;
}

__device__ void sift_down(int32_t* a, int start, int end) {
  int root = start;
  while (2 * root + 1 <= end) {
    int child = 2 * root + 1;
    if (child + 1 <= end && a[child] < a[child + 1]) ++child;
    if (a[root] < a[child]) {
      const int32_t t = a[root];
      a[root] = a[child];
      a[child] = t;
      root = child;
    } else {
      return;
    }
  }
}

// One thread per segment: insertion sort for mesh-sized degrees, heapsort for hubs.
__global__ void segment_sort_kernel(const int32_t* __restrict__ ptr, int64_t N,
                                    int32_t* __restrict__ perm) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  int32_t* a = perm + ptr[v];
  const int n = ptr[v + 1] - ptr[v];
  if (n <= 48) {
    for (int i = 1; i < n; ++i) {
      const int32_t key = a[i];
      int j = i - 1;
      while (j >= 0 && a[j] > key) {
        a[j + 1] = a[j];
        --j;
      }
      a[j + 1] = key;
    }
  } else {
    for (int s = (n - 2) / 2; s >= 0; --s) sift_down(a, s, n - 1);
    for (int end = n - 1; end > 0; --end) {
      const int32_t t = a[end];
      a[end] = a[0];
      a[0] = t;
      sift_down(a, 0, end - 1);
    }
  }
}

__global__ void fill_csr_kernel(const int32_t* __restrict__ senders,
                                const int32_t* __restrict__ receivers, int64_t E, int base,
                                const int32_t* __restrict__ perm, int32_t* __restrict__ send_csr,
                                int32_t* __restrict__ recv_csr, int32_t* __restrict__ inv_perm) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= E) return;
  const int32_t e = perm[j];
  send_csr[j] = senders[e] - base;
  recv_csr[j] = receivers[e] - base;
  inv_perm[e] = (int32_t)j;
}

__global__ void fill_csc_kernel(const int32_t* __restrict__ perm_sender,
                                const int32_t* __restrict__ inv_perm, int64_t E,
                                int32_t* __restrict__ csc_slot) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= E) return;
  csc_slot[j] = inv_perm[perm_sender[j]];
}

inline unsigned nblk(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace

int32_t build_graph_index(mgn_graph* g, const int32_t* d_senders, const int32_t* d_receivers,
                          cudaStream_t st) {
  const int64_t N = g->N, E = g->E;
  const int base = g->index_base;
  MGN_CUDA_TRY(cudaMalloc(&g->row_ptr, sizeof(int32_t) * (N + 1)));
  MGN_CUDA_TRY(cudaMalloc(&g->col_ptr, sizeof(int32_t) * (N + 1)));
  const size_t eb = sizeof(int32_t) * (size_t)(E > 0 ? E : 1);
  MGN_CUDA_TRY(cudaMalloc(&g->perm, eb));
  MGN_CUDA_TRY(cudaMalloc(&g->send_csr, eb));
  MGN_CUDA_TRY(cudaMalloc(&g->recv_csr, eb));
  MGN_CUDA_TRY(cudaMalloc(&g->perm_sender, eb));
  MGN_CUDA_TRY(cudaMalloc(&g->csc_slot, eb));
  int32_t *counts = nullptr, *inv_perm = nullptr, *flags = nullptr;
  MGN_CUDA_TRY(cudaMalloc(&counts, sizeof(int32_t) * (N + 1)));
  MGN_CUDA_TRY(cudaMalloc(&inv_perm, eb));
  MGN_CUDA_TRY(cudaMalloc(&flags, sizeof(int32_t) * 2));
  MGN_CUDA_TRY(cudaMemsetAsync(flags, 0, sizeof(int32_t) * 2, st));

  for (int pass = 0; pass < 2; ++pass) {
    const int32_t* keys = pass == 0 ? d_receivers : d_senders;
    int32_t* ptr = pass == 0 ? g->row_ptr : g->col_ptr;
    int32_t* perm = pass == 0 ? g->perm : g->perm_sender;
    MGN_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (N + 1), st));
    if (E > 0) validate_hist_kernel<<<nblk(E), 256, 0, st>>>(keys, E, N, base, counts, flags);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, N, ptr, pass == 0 ? flags + 1 : nullptr);
    MGN_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (N + 1), st));
    if (pass == 0) {
      // stop before placement if an id is out of range
      int32_t h[2];
      MGN_CUDA_TRY(cudaMemcpyAsync(h, flags, sizeof(h), cudaMemcpyDeviceToHost, st));
      MGN_CUDA_TRY(cudaStreamSynchronize(st));
      if (h[0]) {
        cudaFree(counts); cudaFree(inv_perm); cudaFree(flags);
        return fail(MGN_ERR_INDEX, "receiver id outside [index_base, index_base + n_nodes)");
      }
      g->max_in_degree = h[1];
    } else {
      int32_t h = 0;
      MGN_CUDA_TRY(cudaMemcpyAsync(&h, flags, sizeof(h), cudaMemcpyDeviceToHost, st));
      MGN_CUDA_TRY(cudaStreamSynchronize(st));
      if (h) {
        cudaFree(counts); cudaFree(inv_perm); cudaFree(flags);
        return fail(MGN_ERR_INDEX, "sender id outside [index_base, index_base + n_nodes)");
      }
    }
    if (E > 0) {
      place_kernel<<<nblk(E), 256, 0, st>>>(keys, E, base, ptr, counts, perm);
      segment_sort_kernel<<<nblk(N, 128), 128, 0, st>>>(ptr, N, perm);
    }
  }
  if (E > 0) {
    fill_csr_kernel<<<nblk(E), 256, 0, st>>>(d_senders, d_receivers, E, base, g->perm, g->send_csr,
                                             g->recv_csr, inv_perm);
    fill_csc_kernel<<<nblk(E), 256, 0, st>>>(g->perm_sender, inv_perm, E, g->csc_slot);
  }
  MGN_CUDA_TRY(cudaGetLastError());
  MGN_CUDA_TRY(cudaStreamSynchronize(st));
  cudaFree(counts);
  cudaFree(inv_perm);
  cudaFree(flags);

  // Node-aligned edge tiles (host greedy packing of whole CSR segments into <= 128 rows / nodes).
  {
    std::vector<int32_t> rp((size_t)N + 1);
    MGN_CUDA_TRY(cudaMemcpy(rp.data(), g->row_ptr, sizeof(int32_t) * (N + 1), cudaMemcpyDeviceToHost));
    std::vector<int32_t> trow{0}, tnode{0};
    bool ok = true;
    int rows = 0, nodes = 0;
    for (int64_t v = 0; v < N; ++v) {
      const int deg = rp[v + 1] - rp[v];
      if (deg > 128) ok = false;
      if (rows + deg > 128 || nodes == 128) {
        trow.push_back(rp[v]);
        tnode.push_back((int32_t)v);
        rows = 0;
        nodes = 0;
      }
      rows += deg;
      ++nodes;
    }
    trow.push_back(rp[N]);
    tnode.push_back((int32_t)N);
    g->tiles_ok = ok;
    g->n_edge_tiles = (int32_t)trow.size() - 1;
    MGN_CUDA_TRY(cudaMalloc(&g->tile_row_start, sizeof(int32_t) * trow.size()));
    MGN_CUDA_TRY(cudaMalloc(&g->tile_node_start, sizeof(int32_t) * tnode.size()));
    MGN_CUDA_TRY(cudaMemcpy(g->tile_row_start, trow.data(), sizeof(int32_t) * trow.size(), cudaMemcpyHostToDevice));
    MGN_CUDA_TRY(cudaMemcpy(g->tile_node_start, tnode.data(), sizeof(int32_t) * tnode.size(), cudaMemcpyHostToDevice));
    // CSC slot -> row in tile-image space (tensors stored as [tile][128 rows] images are gathered through this)
    if (E > 0) {
      std::vector<int32_t> slot((size_t)E), pos((size_t)E), tile_of((size_t)E);
      MGN_CUDA_TRY(cudaMemcpy(slot.data(), g->csc_slot, sizeof(int32_t) * E, cudaMemcpyDeviceToHost));
      for (size_t t = 0; t + 1 < trow.size(); ++t)
        for (int32_t j = trow[t]; j < trow[t + 1]; ++j) tile_of[(size_t)j] = (int32_t)t;
      for (int64_t j = 0; j < E; ++j) {
        const int32_t sl = slot[(size_t)j], t = tile_of[(size_t)sl];
        pos[(size_t)j] = t * 128 + (sl - trow[(size_t)t]);
      }
      MGN_CUDA_TRY(cudaMalloc(&g->csc_pos, sizeof(int32_t) * E));
      MGN_CUDA_TRY(cudaMemcpy(g->csc_pos, pos.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice));
    }
  }
  return MGN_OK;
}

}  // namespace mgn
