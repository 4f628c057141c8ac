// Test-only probe of the tcgen05 building blocks (tc_ptx.cuh): one CTA computes
//   mode 0: D[128][128] = A[128][K] * B[128][K]^T      (A, B row-major, both K-major operands)
//   mode 1: D[128][128] = A[K][128]^T * B[K][128]      (A, B row-major, both MN-major operands,
//                                                        the weight-gradient form dW = dZ^T H)
// with K a multiple of 64 (<= 256).  Exercised by tests/test_gpu_tc_probe.py against torch.
// NOT part of the product library: built on its own as libmgn_b200_probe.so (build.py), loaded by that test only.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc.cuh"

namespace mgn {
// the probe library is self-contained: its own copies of the two error helpers of abi.cu
static thread_local std::string g_probe_error;
void set_error(const std::string& msg) { g_probe_error = msg; }
int32_t fail(int32_t code, const std::string& msg) {
  g_probe_error = msg;
  return code;
}
namespace {
using namespace tc;

__global__ void __launch_bounds__(128) umma_probe_kernel(const __nv_bfloat16* __restrict__ A,
                                                         const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                                                         int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nkb = K / 64;
  // mode 0: tile kb of A = rows 0..127, cols kb*64..+63 of A[128][K]   (nkb tiles per operand)
  // mode 1: tile (kb, slab) of A = rows kb*128.. hmm K rows; here K <= 256 rows are split into
  //         tiles of 128 K-rows x 64 columns: tile index = kr * 2 + slab, kr = K-row block of 128.
  const uint32_t sA = smem_u32(smem), sB = sA + 4 * kTileBytes;
  if (mode == 0) {
    for (int i = tid; i < nkb * 128 * 8; i += 128) {
      const int kb = i / 1024, row = (i / 8) % 128, chunk = i % 8;
      const uint4 va = *reinterpret_cast<const uint4*>(A + (size_t)row * K + kb * 64 + chunk * 8);
      const uint4 vb = *reinterpret_cast<const uint4*>(B + (size_t)row * K + kb * 64 + chunk * 8);
      st_shared_v4(sA + kb * kTileBytes + t128_off(row, chunk), va.x, va.y, va.z, va.w);
      st_shared_v4(sB + kb * kTileBytes + t128_off(row, chunk), vb.x, vb.y, vb.z, vb.w);
    }
  } else {
    const int nkr = (K + 127) / 128;  // K-row blocks of 128 rows (zero padded)
    for (int i = tid; i < nkr * 2 * 128 * 8; i += 128) {
      const int t = i / 1024, row = (i / 8) % 128, chunk = i % 8;
      const int kr = t / 2, slab = t % 2;
      const int krow = kr * 128 + row;
      uint4 va = make_uint4(0, 0, 0, 0), vb = va;
      if (krow < K) {
        va = *reinterpret_cast<const uint4*>(A + (size_t)krow * 128 + slab * 64 + chunk * 8);
        vb = *reinterpret_cast<const uint4*>(B + (size_t)krow * 128 + slab * 64 + chunk * 8);
      }
      st_shared_v4(sA + t * kTileBytes + t128_off(row, chunk), va.x, va.y, va.z, va.w);
      st_shared_v4(sB + t * kTileBytes + t128_off(row, chunk), vb.x, vb.y, vb.z, vb.w);
    }
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 128);
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    if (mode == 0) {
      const uint32_t idesc = umma_idesc(128, 128, false, false);
      for (int kb = 0; kb < nkb; ++kb)
        for (int k = 0; k < 4; ++k)
          umma(tmem, desc_kmajor(sA + kb * kTileBytes, k), desc_kmajor(sB + kb * kTileBytes, k), idesc,
               (kb | k) != 0);
    } else {
      const uint32_t idesc = umma_idesc(128, 128, true, true);
      const int nsteps = K / 16;
      for (int s = 0; s < nsteps; ++s) {
        const int kr = s / 8, ks = s % 8;  // 8 steps of 16 K-rows per 128-row tile
        umma(tmem, desc_mnmajor(sA + kr * 2 * kTileBytes, kTileBytes, ks),
             desc_mnmajor(sB + kr * 2 * kTileBytes, kTileBytes, ks), idesc, s != 0);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int row = tid;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)row * 128 + c * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace
}  // namespace mgn

extern "C" int32_t mgn_debug_umma_probe(const void* d_a, const void* d_b, float* d_out, int32_t K, int32_t mode,
                                        void* stream) {
  using namespace mgn;
  MGN_REQUIRE(d_a && d_b && d_out, "umma_probe: null argument");
  MGN_REQUIRE(K > 0 && K <= 256 && (mode == 0 ? K % 64 == 0 : K % 16 == 0), "umma_probe: bad K");
  const size_t smem = 8 * tc::kTileBytes + 1024;
  MGN_CUDA_TRY(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(d_a), static_cast<const __nv_bfloat16*>(d_b), d_out, K, mode);
  MGN_CUDA_TRY(cudaGetLastError());
  return MGN_OK;
}

