// Multi-GPU transport behind the C ABI (include/mgn_b200.h, "multi-GPU transport"): NCCL loaded at run time, the
// data-parallel gradient all-reduce, the online-normaliser merge, the halo exchange of partitioned meshes and
// mgn_backward_dp = backward + bucketed all-reduce + Adam, overlapped with the tail of the backward pass.
//
// The reference has no counterpart (single device, src/MeshGraphNets.jl:257); the semantics are those of SURVEY 8e:
// P-way data parallelism = the batch-P SGD that `batchsize` (src/MeshGraphNets.jl:224) leaves unimplemented.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace mgn {
namespace {

// ---- the few NCCL declarations used (stable since NCCL 2.10; resolved with dlsym, so no link-time dependency) ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclFloat32 = 7 };
enum { ncclSum = 0, ncclAvg = 4 };

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};

Nccl* nccl() {
  static std::once_flag once;
  static Nccl n;
  std::call_once(once, [] {
    // an already loaded libnccl.so.2 (e.g. the one a host framework brought) is reused by soname
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.lib) break;
    }
    if (!n.lib) {
      n.error = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    auto sym = [&](const char* s) -> void* {
      void* p = dlsym(n.lib, s);
      if (!p && n.error.empty()) n.error = std::string("libnccl lacks ") + s;
      return p;
    };
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
    n.Send = reinterpret_cast<decltype(n.Send)>(sym("ncclSend"));
    n.Recv = reinterpret_cast<decltype(n.Recv)>(sym("ncclRecv"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &n;
}

int32_t nccl_ready(Nccl** out) {
  Nccl* n = nccl();
  if (!n->error.empty()) return fail(MGN_ERR_UNSUPPORTED, n->error);
  *out = n;
  return MGN_OK;
}

#define MGN_NCCL_TRY(n, expr)                                                                         \
  do {                                                                                                \
    int _r = (expr);                                                                                  \
    if (_r != ncclSuccess) return ::mgn::fail(MGN_ERR_CUDA, std::string(#expr) + ": " + (n)->GetErrorString(_r)); \
  } while (0)

__global__ void norm_delta_kernel(float* __restrict__ state, const float* __restrict__ prev, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) state[i] = state[i] - prev[i];
}
__global__ void norm_merge_kernel(float* __restrict__ state, const float* __restrict__ prev, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) state[i] = prev[i] + state[i];
}

}  // namespace

}  // namespace mgn

struct mgn_comm {
  mgn::ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t side = nullptr;             // the stream bucketed collectives / Adam updates run on
  cudaEvent_t ev_ready[64] = {};           // bucket b's gradients are final on the caller's stream
  cudaEvent_t ev_join = nullptr;
};

namespace mgn {
namespace {

// Side stream + events for mgn_backward_dp without a communicator (single GPU, Adam only): per device.
struct LocalSide {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_ready[64] = {};
  cudaEvent_t ev_join = nullptr;
};
LocalSide* local_side() {
  static PerDeviceOnce once;
  static LocalSide sides[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  cudaError_t e = once.run([&](int d) {
    cudaError_t r = cudaStreamCreateWithFlags(&sides[d].side, cudaStreamNonBlocking);
    for (int i = 0; i < 64 && r == cudaSuccess; ++i) r = cudaEventCreateWithFlags(&sides[d].ev_ready[i], cudaEventDisableTiming);
    if (r == cudaSuccess) r = cudaEventCreateWithFlags(&sides[d].ev_join, cudaEventDisableTiming);
    return r;
  });
  return e == cudaSuccess ? &sides[dev] : nullptr;
}

struct BucketHook : GradHook {
  const mgn_model* m;
  float* params;
  float* grads;
  mgn_comm* comm;
  const mgn_adam_config* adam;
  Nccl* n = nullptr;
  cudaStream_t st, side;
  cudaEvent_t* ev_ready;
  std::vector<size_t> bucket_lo;  // bucket b covers MLPs [bucket_lo[b], bucket_lo[b+1]); b ascending = parameter order
  int next_bucket;                // buckets complete from the last one down
  bool ticked = false;

  int64_t mlp_begin(size_t mi) const { return mi < m->mlps.size() ? m->mlps[mi].w_off[0] : m->n_params; }

  int32_t mlp_done(size_t mi, cudaStream_t where) override {
    while (next_bucket >= 0 && mi <= bucket_lo[next_bucket]) {
      const int b = next_bucket--;
      const int64_t lo = mlp_begin(bucket_lo[b]), hi = mlp_begin(bucket_lo[b + 1]);
      MGN_CUDA_TRY(cudaEventRecord(ev_ready[b], where));
      MGN_CUDA_TRY(cudaStreamWaitEvent(side, ev_ready[b], 0));
      if (comm && comm->world > 1)
        MGN_NCCL_TRY(n, n->AllReduce(grads + lo, grads + lo, (size_t)(hi - lo), ncclFloat32, ncclAvg, comm->comm, side));
      if (adam) {
        if (!ticked) {
          MGN_CUDA_TRY(adam_tick(adam->d_state16, adam->beta1, adam->beta2, side));
          ticked = true;
        }
        MGN_CUDA_TRY(adam_apply_range(params + lo, grads + lo, adam->d_m + lo, adam->d_v + lo, hi - lo, adam->lr,
                                      adam->beta1, adam->beta2, adam->eps, adam->d_state16, side));
      }
    }
    return MGN_OK;
  }
};

}  // namespace
}  // namespace mgn

using namespace mgn;

extern "C" {

int32_t mgn_dp_unique_id(void* h_id128) {
  MGN_REQUIRE(h_id128, "dp_unique_id: null buffer");
  Nccl* n = nullptr;
  MGN_TRY(nccl_ready(&n));
  ncclUniqueId id;
  MGN_NCCL_TRY(n, n->GetUniqueId(&id));
  static_assert(sizeof(id) == MGN_DP_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  std::memcpy(h_id128, &id, sizeof(id));
  return MGN_OK;
}

int32_t mgn_dp_init(const void* h_id128, int32_t rank, int32_t world, mgn_comm** out) {
  MGN_REQUIRE(h_id128 && out, "dp_init: null argument");
  *out = nullptr;
  MGN_REQUIRE(world >= 1 && rank >= 0 && rank < world, "dp_init: bad rank / world");
  Nccl* n = nullptr;
  MGN_TRY(nccl_ready(&n));
  mgn_comm* c = new mgn_comm();
  c->rank = rank;
  c->world = world;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
  for (int i = 0; i < 64 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->ev_ready[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    mgn_dp_finalize(c);
    return fail(MGN_ERR_CUDA, std::string("dp_init: ") + cudaGetErrorString(e));
  }
  ncclUniqueId id;
  std::memcpy(&id, h_id128, sizeof(id));
  const int r = n->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    c->comm = nullptr;
    mgn_dp_finalize(c);
    return fail(MGN_ERR_CUDA, std::string("ncclCommInitRank: ") + n->GetErrorString(r));
  }
  *out = c;
  return MGN_OK;
}

int32_t mgn_dp_finalize(mgn_comm* c) {
  if (!c) return MGN_OK;
  Nccl* n = nccl();
  if (c->comm && n->CommDestroy) n->CommDestroy(c->comm);
  for (auto& e : c->ev_ready)
    if (e) cudaEventDestroy(e);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->side) cudaStreamDestroy(c->side);
  delete c;
  return MGN_OK;
}

int32_t mgn_dp_rank(const mgn_comm* c, int32_t* rank, int32_t* world) {
  MGN_REQUIRE(c, "dp_rank: null communicator");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return MGN_OK;
}

int32_t mgn_dp_allreduce(mgn_comm* c, float* d_buf, int64_t n_elems, int32_t op, void* stream) {
  MGN_REQUIRE(c && (d_buf || n_elems == 0) && n_elems >= 0, "dp_allreduce: bad argument");
  MGN_REQUIRE(op == MGN_DP_SUM || op == MGN_DP_MEAN, "dp_allreduce: unknown op");
  if (n_elems == 0 || c->world == 1) return MGN_OK;
  Nccl* n = nullptr;
  MGN_TRY(nccl_ready(&n));
  MGN_NCCL_TRY(n, n->AllReduce(d_buf, d_buf, (size_t)n_elems, ncclFloat32, op == MGN_DP_MEAN ? ncclAvg : ncclSum,
                               c->comm, static_cast<cudaStream_t>(stream)));
  return MGN_OK;
}

int32_t mgn_dp_allreduce_normaliser(mgn_comm* c, float* d_state, const float* d_prev, int32_t n_floats, void* stream) {
  MGN_REQUIRE(c && d_state && d_prev && n_floats > 0, "dp_allreduce_normaliser: bad argument");
  if (c->world == 1) return MGN_OK;
  Nccl* n = nullptr;
  MGN_TRY(nccl_ready(&n));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (n_floats + 127) / 128;
  norm_delta_kernel<<<blocks, 128, 0, st>>>(d_state, d_prev, n_floats);
  MGN_NCCL_TRY(n, n->AllReduce(d_state, d_state, (size_t)n_floats, ncclFloat32, ncclSum, c->comm, st));
  norm_merge_kernel<<<blocks, 128, 0, st>>>(d_state, d_prev, n_floats);
  MGN_CUDA_TRY(cudaGetLastError());
  return MGN_OK;
}

int32_t mgn_halo_exchange(mgn_comm* c, const void* d_send, const int64_t* h_send_rows, void* d_recv,
                          const int64_t* h_recv_rows, int64_t row_bytes, void* stream) {
  MGN_REQUIRE(c && h_send_rows && h_recv_rows && row_bytes > 0, "halo_exchange: bad argument");
  Nccl* n = nullptr;
  MGN_TRY(nccl_ready(&n));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* sp = static_cast<const char*>(d_send);
  char* rp = static_cast<char*>(d_recv);
  int64_t total = 0;
  for (int p = 0; p < c->world; ++p) {
    MGN_REQUIRE(h_send_rows[p] >= 0 && h_recv_rows[p] >= 0, "halo_exchange: negative row count");
    total += h_send_rows[p] + h_recv_rows[p];
  }
  if (total == 0) return MGN_OK;
  MGN_NCCL_TRY(n, n->GroupStart());
  int rc = ncclSuccess;
  for (int p = 0; p < c->world && rc == ncclSuccess; ++p) {
    const size_t sb = (size_t)h_send_rows[p] * (size_t)row_bytes, rb = (size_t)h_recv_rows[p] * (size_t)row_bytes;
    if (p == c->rank) {  // rows a rank "sends to itself" (none in a partition plan) are a plain copy
      if (sb) cudaMemcpyAsync(rp, sp, std::min(sb, rb), cudaMemcpyDeviceToDevice, st);
    } else {
      if (sb) rc = n->Send(sp, sb, ncclUint8, p, c->comm, st);
      if (rb && rc == ncclSuccess) rc = n->Recv(rp, rb, ncclUint8, p, c->comm, st);
    }
    sp += sb;
    rp += rb;
  }
  const int rc2 = n->GroupEnd();
  MGN_NCCL_TRY(n, rc);
  MGN_NCCL_TRY(n, rc2);
  return MGN_OK;
}

int32_t mgn_backward_dp(const mgn_model* m, const mgn_graph* g, float* d_params, const float* d_nf, const float* d_ef,
                        const float* d_dout, float* d_dparams, float* d_dnf, void* d_workspace, size_t workspace_bytes,
                        mgn_comm* comm, const mgn_adam_config* adam, int32_t n_buckets, void* stream) {
  MGN_REQUIRE(m && g && d_params && d_nf && d_dout && d_dparams && d_workspace, "backward_dp: null argument");
  MGN_REQUIRE(g->E == 0 || d_ef, "backward_dp: null edge features");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool dist = comm != nullptr && comm->world > 1;
  if (!dist && !adam)
    return backward(m, g, d_params, d_nf, d_ef, d_dout, d_dparams, d_dnf, d_workspace, workspace_bytes, st);
  if (adam) MGN_REQUIRE(adam->d_m && adam->d_v && adam->d_state16, "backward_dp: incomplete Adam configuration");
  BucketHook h;
  h.m = m;
  h.params = d_params;
  h.grads = d_dparams;
  h.comm = dist ? comm : nullptr;
  h.adam = adam;
  h.st = st;
  cudaEvent_t ev_join;
  if (comm) {
    int dev = 0;
    MGN_CUDA_TRY(cudaGetDevice(&dev));
    MGN_REQUIRE(dev == comm->device, "backward_dp: the communicator belongs to another device");
    h.side = comm->side;
    h.ev_ready = comm->ev_ready;
    ev_join = comm->ev_join;
  } else {
    LocalSide* ls = local_side();
    if (!ls) return fail(MGN_ERR_CUDA, "backward_dp: cannot create the side stream");
    h.side = ls->side;
    h.ev_ready = ls->ev_ready;
    ev_join = ls->ev_join;
  }
  if (dist) MGN_TRY(nccl_ready(&h.n));
  // buckets of consecutive MLPs with about equal parameter counts (at most 64)
  const size_t n_mlp = m->mlps.size();
  const int nb = (int)std::max<size_t>(1, std::min<size_t>({(size_t)std::max(n_buckets, 1), n_mlp, (size_t)64}));
  h.bucket_lo.assign(1, 0);
  for (int b = 1; b < nb; ++b) {
    const int64_t want = m->n_params * b / nb;
    size_t mi = h.bucket_lo.back() + 1;
    while (mi < n_mlp - (size_t)(nb - b) && m->mlps[mi].w_off[0] < want) ++mi;
    h.bucket_lo.push_back(mi);
  }
  h.bucket_lo.push_back(n_mlp);
  h.next_bucket = nb - 1;
  MGN_TRY(backward(m, g, d_params, d_nf, d_ef, d_dout, d_dparams, d_dnf, d_workspace, workspace_bytes, st, &h));
  MGN_TRY(h.mlp_done(0, st));  // whatever is left (nothing, unless a path skipped a notification)
  MGN_CUDA_TRY(cudaEventRecord(ev_join, h.side));
  MGN_CUDA_TRY(cudaStreamWaitEvent(st, ev_join, 0));
  return MGN_OK;
}

int32_t mgn_library_release(void) {
  release_device_state();
  return MGN_OK;
}

}  // extern "C"
