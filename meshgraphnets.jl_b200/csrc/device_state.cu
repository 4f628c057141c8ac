// Per-device / per-stream state of libmgn_b200.so (see common.cuh): everything the launch path needs that is not
// owned by a caller-visible handle lives here, behind call_once / a mutex.
#include <cstdlib>
#include <map>
#include <tuple>

#include "common.cuh"

namespace mgn {

int device_sm_count() {
  static PerDeviceOnce once;
  static int n_sm[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  once.run([&](int d) {
    int n = 0;
    cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    n_sm[d] = (e == cudaSuccess && n > 0) ? n : 148;
    return cudaSuccess;
  });
  return n_sm[dev];
}

namespace {
std::mutex g_scratch_mu;
std::map<std::tuple<int, cudaStream_t, int>, std::pair<void*, size_t>> g_scratch;
std::vector<std::pair<int, void*>> g_retired;  // outgrown buffers: an enqueued kernel or a captured graph may still use them
}  // namespace

cudaError_t stream_scratch(cudaStream_t st, int kind, size_t bytes, void** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  auto key = std::make_tuple(dev, st, kind);
  auto it = g_scratch.find(key);
  if (it != g_scratch.end() && it->second.second >= bytes) {
    *out = it->second.first;
    return cudaSuccess;
  }
  // cudaMalloc is not allowed while a capture in global / thread-local mode is under way on this thread: relax the
  // mode for the allocation only (the buffer is a plain device pointer baked into the captured nodes)
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  cudaThreadExchangeStreamCaptureMode(&mode);
  void* p = nullptr;
  e = cudaMalloc(&p, bytes);
  cudaThreadExchangeStreamCaptureMode(&mode);
  if (e != cudaSuccess) return e;
  if (it != g_scratch.end()) g_retired.push_back({dev, it->second.first});
  g_scratch[key] = {p, bytes};
  *out = p;
  return cudaSuccess;
}

void release_device_state() {
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& kv : g_scratch) {
    cudaSetDevice(std::get<0>(kv.first));
    cudaFree(kv.second.first);
  }
  for (auto& r : g_retired) {
    cudaSetDevice(r.first);
    cudaFree(r.second);
  }
  g_scratch.clear();
  g_retired.clear();
  cudaSetDevice(cur);
}

ReduceLane* reduce_lane() {
  static PerDeviceOnce once;
  static ReduceLane lanes[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  cudaError_t e = once.run([&](int d) {
    ReduceLane& L = lanes[d];
    cudaError_t r = cudaStreamCreateWithFlags(&L.side, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && r == cudaSuccess; ++i) {
      r = cudaEventCreateWithFlags(&L.fork[i], cudaEventDisableTiming);
      if (r == cudaSuccess) r = cudaEventCreateWithFlags(&L.done[i], cudaEventDisableTiming);
    }
    if (r == cudaSuccess) r = cudaEventCreateWithFlags(&L.join, cudaEventDisableTiming);
    return r;
  });
  return e == cudaSuccess ? &lanes[dev] : nullptr;
}

TuneKnobs read_tune_knobs() {
  TuneKnobs k;
  if (const char* e = getenv("MGN_FWD_EPI_WARPS")) k.fwd_epi_warps = atoi(e) == 4 ? 4 : 8;
  if (const char* e = getenv("MGN_FWD_STAGGER_NS")) k.fwd_stagger_ns = atoi(e) > 0 ? atoi(e) : 0;
  if (const char* e = getenv("MGN_FWD_DEEP_RING")) k.fwd_deep_ring = atoi(e) == 0 ? 0 : 1;
  if (const char* e = getenv("MGN_FWD_PERSIST")) k.fwd_persist = atoi(e) == 0 ? 0 : (atoi(e) == 2 ? 2 : 1);
  if (const char* e = getenv("MGN_RECOMPUTE")) k.recompute = atoi(e) == 1 ? 1 : 0;
  if (const char* e = getenv("MGN_REDUCE_LANE")) k.reduce_lane = atoi(e) == 0 ? 0 : 1;
  if (const char* e = getenv("MGN_PDL")) k.pdl = atoi(e) == 1 ? 1 : 0;
  return k;
}

}  // namespace mgn
