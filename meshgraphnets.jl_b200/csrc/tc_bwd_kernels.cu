// tcgen05 kernels of the backward pass (MGN_COMPUTE_BF16), sm_100a only.
//
// mlp_bwd_chain_kernel   one launch = the Dense layers top .. 1 of one GraphNetCore MLP, for every 128-row
//                        tile the CTA owns.  Per tile: the head warps do the LayerNorm backward straight from
//                        the fp32 gradient of the MLP output (gathering the aggregation adjoint d_agg[recv],
//                        SURVEY 8 a11) and write dZ_top as a 128B-swizzled bf16 operand tile; then for each
//                        layer one MMA thread issues  dX = dZ W^T  (K-major operands) into a TMEM
//                        accumulator and  dW += H^T dZ  (the SAME shared-memory bytes read as MN-major
//                        operands) into per-layer TMEM accumulators that live across all tiles of the CTA -
//                        weight gradients are reduced on chip and leave the SM once per launch.  The
//                        epilogue warps apply the ReLU mask and write the next dZ tile back to shared memory;
//                        hidden gradients never touch HBM.
// mlp_bwd_input_kernel   first Dense layer: dW_0 += X^T dZ_0 with X gathered exactly like the forward operand
//                        (sender rows, receiver rows, edge latent) and dX = dZ_0 W_0^T scattered by fused sinks:
//                        bf16 rows for the sender adjoint, a deterministic CSR segmented sum for the receiver
//                        adjoint, an in-place fp32 add for the edge-latent gradient.
// Cross-CTA reduction of the weight-gradient partials is a fixed-order sum (reduce_pieces): deterministic.
#include "tc.cuh"
#include "tc_ptx.cuh"

namespace mgn {
namespace tc {
namespace {

constexpr uint32_t kImg = 2 * kTileB;  // one 128 x 128 bf16 operand (two T128 tiles)

__device__ __forceinline__ void tile_rows(const int32_t* trs, int64_t M, int tile, int64_t& row0, int& cnt) {
  if (trs) {
    row0 = trs[tile];
    cnt = trs[tile + 1] - (int)row0;
  } else {
    row0 = (int64_t)tile * kTile;
    cnt = (int)min((int64_t)kTile, M - row0);
  }
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }

// A warp's 32 rows x 32 fp32 columns of a TMEM accumulator (thread == row: v[32]) -> global rows of 128 floats, through 4 KB of
// warp-private shared memory `scr` (16-byte chunks XOR-swizzled by the row: conflict free both ways).  Stored straight from the
// registers every instruction would touch 32 different lines (one 16-byte piece of each lane's row); read back transposed, eight
// lanes write the 128 contiguous bytes of a row: 4 lines per instruction.  `g`: element (first row of the warp, first column).
__device__ __forceinline__ void store_acc_chunk(uint32_t scr, const float (&v)[32], float* g, int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    st_shared_v4(scr + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4), __float_as_uint(v[4 * q]),
                 __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
  __syncwarp();
  const int ch = lane & 7, sub = lane >> 3;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + sub;
    const uint4 x = ld_shared_v4(scr + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4));
    *reinterpret_cast<uint4*>(g + (size_t)r * 128 + ch * 4) = x;
  }
  __syncwarp();
}

// ======================================================================================================
// Chain kernel
// ======================================================================================================
namespace chain {
// Warps 0-7 epilogue (two threads per tile row: thread (row, half) owns 64 accumulator columns), 8-15 head (LayerNorm
// backward), 16 producer, 17 MMA issue, 18-19 idle (they complete the fifth warpgroup).  The launch allocates 96 registers
// per thread; setmaxnreg then moves them where they are needed: the head's three-deep load pipeline gets 128, the
// epilogue keeps 88, the producer / MMA warpgroup 40.
constexpr int kThreads = 640;
constexpr int kEpiThreads = 256;
constexpr int kHeadThreads = 256;
constexpr int kWarpH = 8, kWarpP = 16, kWarpM = 17;
constexpr int kRegsHead = 128, kRegsEpi = 88, kRegsMisc = 40;   // 256 * 128 + 256 * 88 + 128 * 40 <= 640 * 96
// dZ slots: 0-1 hold the head's output (top dZ of a tile; the head may run a whole tile ahead of the MMAs),
// 2-3 the epilogue's outputs (dZ of the layers below, alternating).  H and W^T images stream through 2-slot rings.
constexpr int kZ = 4, kH = 2, kW = 2;
constexpr uint32_t kSmemZ = 0;
constexpr uint32_t kSmemH = kSmemZ + kZ * kImg;
constexpr uint32_t kSmemW = kSmemH + kH * kImg;
constexpr uint32_t kSmemScale = kSmemW + kW * kTileB;         // ln scale [128]
constexpr uint32_t kSmemIdx = kSmemScale + 512;               // gather rows of dy_b of this and of the next tile, 2 x 128 ints
constexpr uint32_t kSmemRstd = kSmemIdx + 1024;               // rstd of the rows of this and of the next tile, 2 x 128 floats
constexpr uint32_t kSmemTrs = kSmemRstd + 1024;                // head: (first row, row count) of its next 16 tiles, int2 ring
constexpr uint32_t kSmemBar = kSmemTrs + 128;                 // (the head's final [8][3][128] reduction reuses a dZ slot)
constexpr uint32_t kNumBar = 2 * kZ + 2 * kH + 2 * kW + 4;
constexpr uint32_t kSmemTmem = kSmemBar + 8 * kNumBar;
constexpr uint32_t kSmemTotal = kSmemTmem + 16;
// Everything the SM has (227 KB): the dynamic region starts 1024-byte aligned in practice (no static shared memory), so the
// alignment slack is the 336 bytes that are left; the kernel traps if the aligned layout does not fit.
constexpr uint32_t kSmemLaunch = 232448;
static_assert(kSmemTotal <= kSmemLaunch, "chain kernel shared memory");

// kDyImg: the fp32 part of dy comes as bf16 tile images (edge MLPs) - kept as raw bits until it is consumed, which frees
// 16 registers of the head's double-buffered loads compared with the fp32 row-major form (node MLPs, encoders).
template <bool kDyImg>
__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_chain_kernel(const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s_base = smem_u32(smem);
  if (s_base + kSmemTotal > smem_u32(smem_raw) + kSmemLaunch) __trap();
  float* scale_s = reinterpret_cast<float*>(smem + kSmemScale);
  int* idx_s = reinterpret_cast<int*>(smem + kSmemIdx);
  float* rstd_s = reinterpret_cast<float*>(smem + kSmemRstd);
  const uint32_t bar0 = s_base + kSmemBar;
  auto z_full = [&](int s) { return bar0 + 8u * s; };
  auto z_empty = [&](int s) { return bar0 + 8u * (kZ + s); };
  auto h_full = [&](int s) { return bar0 + 8u * (2 * kZ + s); };
  auto h_empty = [&](int s) { return bar0 + 8u * (2 * kZ + kH + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (2 * kZ + 2 * kH + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (2 * kZ + 2 * kH + kW + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * kZ + 2 * kH + 2 * kW);
  const uint32_t acc_empty = acc_full + 8, done_bar = acc_full + 16;
  auto z_slot = [&](int s) { return s_base + kSmemZ + (uint32_t)s * kImg; };
  auto h_slot = [&](int s) { return s_base + kSmemH + (uint32_t)s * kImg; };
  auto w_slot = [&](int s) { return s_base + kSmemW + (uint32_t)s * (uint32_t)kTileB; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemTmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ns = p.nsteps;
  float* my_partial = p.partial + (size_t)blockIdx.x * chain_partial_floats(ns);

  pdl_trigger();   // programmatic dependent launch: see common.cuh
  if (tid == 0) {
    for (int s = 0; s < kZ; ++s) {
      mbar_init(z_full(s), 1);
      mbar_init(z_empty(s), 2);
    }
    for (int s = 0; s < kH; ++s) {
      mbar_init(h_full(s), 1);
      mbar_init(h_empty(s), 2);
    }
    for (int s = 0; s < kW; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpiThreads);
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == kWarpM) tmem_alloc(smem_u32(tmem_slot), 512);
  pdl_wait();      // no global memory access above this line
  if (p.head_mode == HEAD_LN)
    for (int i = tid; i < 128; i += kThreads) scale_s[i] = p.ln_scale[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= kWarpP) {
   reg_dealloc<kRegsMisc>();                              // the whole warpgroup 16-19
   if (warp == kWarpP) {
    // ================================ producer (one thread: bulk copies only) ========================
    if (lane == 0) {
      uint32_t hc = 0, wc = 0, t_local = 0;
      int tn = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t_local) {
        if (p.head_mode == HEAD_IMAGE) {
          const uint32_t s = t_local & 1;
          mbar_wait(z_empty(s), ((t_local >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(z_full(s), kImg);
          bulk_g2s(z_slot(s), reinterpret_cast<const uint8_t*>(p.z_top) + (size_t)tile * kImg, kImg, z_full(s));
          mbar_arrive(z_empty(s));  // stands in for the head warps' "column sums done" arrival
        }
        for (int j = 0; j < ns; ++j) {
          for (int kb = 0; kb < 2; ++kb, ++wc) {
            const uint32_t s = wc % kW;
            mbar_wait(w_empty(s), ((wc / kW) & 1) ^ 1);
            mbar_arrive_expect_tx(w_full(s), (uint32_t)kTileB);
            bulk_g2s(w_slot(s), reinterpret_cast<const uint8_t*>(p.wt_img[j]) + (size_t)kb * kTileB, (uint32_t)kTileB,
                     w_full(s));
          }
          const uint32_t s = hc % kH;
          mbar_wait(h_empty(s), ((hc / kH) & 1) ^ 1);
          trace_ev(p.trace, 3, tn);  // P: H slot free -> load issued
          mbar_arrive_expect_tx(h_full(s), kImg);
          bulk_g2s(h_slot(s), reinterpret_cast<const uint8_t*>(p.h_img[j]) + (size_t)tile * kImg, kImg, h_full(s));
          ++hc;
        }
      }
    }
   } else if (warp == kWarpM) {
    // ================================ MMA issue ================================
    if (lane == 0) {
      const uint32_t idesc_k = umma_idesc(128, 128, false, false);
      const uint32_t idesc_mn = umma_idesc(128, 128, true, true);
      uint32_t hc = 0, wc = 0, t_local = 0, acc_par = 0;
      bool first = true;
      int tn = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t_local) {
        for (int j = 0; j < ns; ++j) {
          // source dZ of this step: the head's output (j == 0) or the epilogue's output of step j - 1
          const uint32_t ec = t_local * ns + j - 1;
          const uint32_t zs = j == 0 ? (t_local & 1) : 2 + (ec & 1);
          const uint32_t zpar = j == 0 ? ((t_local >> 1) & 1) : ((ec >> 1) & 1);
          trace_ev(p.trace, 1, tn);  // M0: step start
          mbar_wait(z_full(zs), zpar);
          trace_ev(p.trace, 1, tn);  // M1: dZ ready
          if (!first) {  // the epilogue has drained the accumulator of the previous step
            mbar_wait(acc_empty, acc_par);
            acc_par ^= 1;
          }
          first = false;
          trace_ev(p.trace, 1, tn);  // M2: accumulator free
          fence_proxy_async();
          tc_fence_after();
          // dX = dZ W^T : A = dZ (K-major over the layer's output features), B = W^T image tile kb
          for (int kb = 0; kb < 2; ++kb, ++wc) {
            const uint32_t ws = wc % kW;
            mbar_wait(w_full(ws), (wc / kW) & 1);
            tc_fence_after();
            const uint32_t a_lo = kdesc_lo(z_slot(zs) + kb * kTileB), b_lo = kdesc_lo(w_slot(ws));
            umma_lo(tmem, a_lo, b_lo, idesc_k, kb != 0);
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_lo(tmem, a_lo + 2 * k, b_lo + 2 * k, idesc_k, true);
            umma_commit(w_empty(ws));
          }
          umma_commit(acc_full);
          trace_ev(p.trace, 1, tn);  // M3: dX issued (weights were there)
          // dW += H^T dZ : both operands read MN-major (rows of the tile are the reduction index)
          const uint32_t hs = hc % kH;
          mbar_wait(h_full(hs), (hc / kH) & 1);
          trace_ev(p.trace, 1, tn);  // M4: H ready
          tc_fence_after();
          const uint32_t d_w = tmem + 128u * (1 + j);
          {
            const uint32_t a_lo = mndesc_lo(h_slot(hs), (uint32_t)kTileB), b_lo = mndesc_lo(z_slot(zs), (uint32_t)kTileB);
            umma_lo(d_w, a_lo, b_lo, idesc_mn, t_local != 0);
#pragma unroll
            for (int ks = 1; ks < 8; ++ks) umma_lo(d_w, a_lo + 128 * ks, b_lo + 128 * ks, idesc_mn, true);
          }
          umma_commit(h_empty(hs));
          umma_commit(z_empty(zs));
          ++hc;
        }
      }
      umma_commit(done_bar);
    }
   }   // warps 18-19: idle
  } else if (warp < kWarpH) {
    // ================================ epilogue (thread == (row, column half)) ================================
    reg_dealloc<kRegsEpi>();                              // warpgroups 0-3, 4-7
    const int row = tid & 127, half = tid >> 7;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t hc = 0, t_local = 0, acc_par = 0;
    int tn = 0;
    // bias-gradient partials: this thread sums 8 columns (one 16-byte chunk) over one sixteenth of the rows, per step
    float db0[8], db1[8], db2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) db0[e] = db1[e] = db2[e] = 0.f;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t_local) {
      int64_t row0;
      int cnt;
      tile_rows(p.tile_row_start, p.M, tile, row0, cnt);
#pragma unroll 1
      for (int j = 0; j < ns; ++j) {
        const uint32_t ec = t_local * ns + j, zs = 2 + (ec & 1), hs = hc % kH;
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E0: step start
        mbar_wait(acc_full, acc_par);
        acc_par ^= 1;
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E1: accumulator full
        mbar_wait(h_full(hs), (hc / kH) & 1);
        mbar_wait(z_empty(zs), ((ec >> 1) & 1) ^ 1);
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E2: H there, Z slot free
        tc_fence_after();
#pragma unroll 1
        for (int c = 2 * half; c < 2 * half + 2; ++c) {
          float v[32];
          tmem_ld32(t_lane + c * 32, v);
          const uint32_t hb = h_slot(hs) + (c >> 1) * kTileB, zb = z_slot(zs) + (c >> 1) * kTileB;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t off = t128_off(row, (c & 1) * 4 + q4);
            const uint4 hq = ld_shared_v4(hb + off);
            const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = (hw[e] & 0x00007fffu) ? v[q4 * 8 + 2 * e] : 0.f;      // ReLU mask: H > 0
              const float b = (hw[e] & 0x7fff0000u) ? v[q4 * 8 + 2 * e + 1] : 0.f;
              w[e] = pack_bf16x2(a, b);
            }
            st_shared_v4(zb + off, w[0], w[1], w[2], w[3]);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(acc_empty);
        named_bar_sync(1, kEpiThreads);
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E3: masked dZ written
        const bool last = j == ns - 1;
        if (tid == 0) {
          mbar_arrive(z_full(zs));
          mbar_arrive(h_empty(hs));
          if (last) {
            bulk_s2g(reinterpret_cast<uint8_t*>(p.dz_out) + (size_t)tile * kImg, z_slot(zs), kImg);
            bulk_commit();
          }
        }
        // bias gradient: column sums of the bf16 dZ just written.  128-bit shared loads (16 per thread instead of
        // 128 narrow ones: the shared-memory pipe is the scarce resource); rows >= cnt hold zeros.
        {
          const uint32_t chunk = (uint32_t)tid & 7u, tsel = ((uint32_t)tid >> 3) & 1u, grp = (uint32_t)tid >> 4;
          const uint32_t zb2 = z_slot(zs) + tsel * (uint32_t)kTileB;
          float t8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t8[e] = 0.f;
#pragma unroll 4
          for (uint32_t rr = 0; rr < 8; ++rr) {
            const uint32_t jr = grp * 8u + rr;
            const uint4 q = ld_shared_v4(zb2 + jr * 128u + ((chunk ^ (jr & 7u)) << 4));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              t8[2 * e] += __uint_as_float(w[e] << 16);
              t8[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (j == 0) db0[e] += t8[e];
            else if (j == 1) db1[e] += t8[e];
            else db2[e] += t8[e];
          }
        }
        named_bar_sync(1, kEpiThreads);
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E4: column sums done
        if (tid == 0) {
          if (last) {
            bulk_wait_read0();
            mbar_arrive(z_empty(zs));  // stands in for the MMA commit: nobody multiplies the last dZ here
          }
          mbar_arrive(z_empty(zs));
        }
        ++hc;
      }
    }
    // ---- weight-gradient partials: TMEM -> global, once per launch
    mbar_wait(done_bar, 0);
    tc_fence_after();
    named_bar_sync(1, kEpiThreads);   // thread 0 has waited for the last dZ bulk store's read: the dZ slots are free
    for (int j = 0; j < ns; ++j) {
#pragma unroll 1
      for (int c = 2 * half; c < 2 * half + 2; ++c) {
        float v[32];
        tmem_ld32(t_lane + 128u * (1 + j) + c * 32, v);
        // staging: 4 KB per epilogue warp in dZ slot 3 (every MMA is done; the bias-gradient reduction below uses slot 2,
        // the head's final reduction slot 0 or 1)
        store_acc_chunk(s_base + kSmemZ + 3u * kImg + (uint32_t)warp * 4096u, v,
                        my_partial + (size_t)j * 16384 + (size_t)((warp & 3) * 32) * 128 + c * 32, lane);
      }
    }
    // bias gradients: [16 row groups][3 steps][128] through an epilogue dZ slot (all MMAs and the last dZ bulk
    // store are done), then a fixed-order sum over the row groups
    named_bar_sync(1, kEpiThreads);
    {
      float* red = reinterpret_cast<float*>(smem + kSmemZ + 2 * kImg);
      const int col0 = (int)((((uint32_t)tid >> 3) & 1u) * 64u + ((uint32_t)tid & 7u) * 8u), grp = tid >> 4;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red[(grp * 3 + 0) * 128 + col0 + e] = db0[e];
        red[(grp * 3 + 1) * 128 + col0 + e] = db1[e];
        red[(grp * 3 + 2) * 128 + col0 + e] = db2[e];
      }
      named_bar_sync(1, kEpiThreads);
      if (tid < 128)
        for (int j = 0; j < ns; ++j) {
          float t = 0.f;
#pragma unroll
          for (int g = 0; g < 16; ++g) t += red[(g * 3 + j) * 128 + tid];
          my_partial[(size_t)ns * 16384 + (size_t)(j + 1) * 128 + tid] = t;
        }
    }
    if (tid == 0) bulk_wait0();
  } else {
    // ================================ head: LayerNorm backward (16 lanes per row, 2 x 32 rows in flight) =====
    reg_alloc<kRegsHead>();                               // warpgroups 8-11, 12-15
    const int lt = tid - 32 * kWarpH, cc = lt & 15, rg = lt >> 4;  // rg in [0, 16)
    float gs[8], gb[8], dbt[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) gs[e] = gb[e] = dbt[e] = 0.f;
    if (p.head_mode == HEAD_LN) {
      const float* sc = scale_s + cc * 8;   // LayerNorm scale of this thread's 8 columns (shared memory: registers are
                                            // the scarce resource of the head, see kDepth below)
      uint32_t t_local = 0;
      int tn = 0;
      // gather rows of dy_b16, one coalesced load per tile, fetched ONE TILE AHEAD into the other half of idx_s so that
      // no batch ever waits on a dependent index load (the barrier that ends a tile publishes the next tile's rows)
      // (the load is issued at the top of a tile and its value stored at the bottom: a store right behind the load would
      // stall the thread for a whole memory latency in front of the tile's first batch)
      auto load_idx = [&](int tile) -> int {   // kernel start only: reads tile_row_start itself
        int v = 0;
        if (p.dy_b16 && lt < kTile && tile < p.n_tiles) {
          int64_t r0;
          int n;
          tile_rows(p.tile_row_start, p.M, tile, r0, n);
          if (lt < n) v = p.b_idx ? p.b_idx[r0 + lt] : (int)(r0 + lt);
        }
        return v;
      };
      // rstd of a tile's rows travels the same way (one load per row and tile instead of one per row and 16-lane group)
      auto load_rstd = [&](int tile) -> float {
        float v = 0.f;
        if (lt < kTile && tile < p.n_tiles) {
          int64_t r0;
          int n;
          tile_rows(p.tile_row_start, p.M, tile, r0, n);
          if (lt < n) v = p.rstd[r0 + lt];
        }
        return v;
      };
      if (lt < kTile) {
        idx_s[lt] = load_idx(blockIdx.x);
        rstd_s[lt] = load_rstd(blockIdx.x);
      }
      // Row ranges of the CTA's tiles come from a 16-entry ring in shared memory that one thread refills 15 tiles ahead:
      // tile_row_start is a dependent global load (~2 us under load) that used to stand in front of every tile TWICE - at
      // its top, and again in front of the index / rstd loads of the next tile.
      int2* trs_s = reinterpret_cast<int2*>(smem + kSmemTrs);
      auto tile_range = [&](int tile) -> int2 {
        int64_t r0;
        int n;
        tile_rows(p.tile_row_start, p.M, tile, r0, n);
        return make_int2((int)r0, n);
      };
      if (lt < 16 && blockIdx.x + lt * (int)gridDim.x < p.n_tiles) trs_s[lt] = tile_range(blockIdx.x + lt * gridDim.x);
      named_bar_sync(2, kHeadThreads);
      // Batches of 32 rows (2 per thread: i = 32 b + 2 rg + u) flow through a register pipeline kDepth batches deep: the
      // loads of batch b + kDepth - 1 are issued before batch b is computed.  In the image form everything arrives as
      // 16-byte bf16 chunks kept as raw bits (edge MLPs: the gradient image, the gathered d_agg row, xhat: 12 registers
      // per row).  Two stages: a third one (round 2a) measured the same on the device - the head does not wait for want of
      // bytes in flight - and pushed the accumulators into local memory.
      constexpr int kDepth = 2;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t_local) {
        const int2 rc = trs_s[t_local & 15];
        const int64_t row0 = rc.x;
        const int cnt = rc.y;
        const uint32_t zs = t_local & 1;
        const int* idx_c = idx_s + (t_local & 1) * kTile;
        const float* rstd_c = rstd_s + (t_local & 1) * kTile;
        if (lt == 0) trace_ev(p.trace, 0, tn);  // L0: tile start
        const uint8_t* ximg = reinterpret_cast<const uint8_t*>(p.xhat) + (size_t)tile * kImg + (cc >> 3) * kTileB;
        const uint8_t* dimg = reinterpret_cast<const uint8_t*>(p.dy_a_img) + (size_t)tile * kImg + (cc >> 3) * kTileB;
        const uint32_t zb = z_slot(zs) + (cc >> 3) * kTileB;
        float4 a0[kDyImg ? 1 : 2 * kDepth], a1[kDyImg ? 1 : 2 * kDepth];  // [stage h][row u] at 2 h + u
        uint4 aq[kDyImg ? 2 * kDepth : 1], cq[kDyImg ? 2 * kDepth : 1];
        uint4 xq[2 * kDepth];
        uint32_t same = 0;  // bit k: row k shares its receiver with row k - 1 (its gather is the previous row's)
        auto issue = [&](int b, int h) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = 32 * b + rg * 2 + u, k = 2 * h + u;
            if constexpr (kDyImg) {
              aq[k] = make_uint4(0u, 0u, 0u, 0u);
              cq[k] = aq[k];
            } else {
              a0[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              a1[k] = a0[k];
            }
            xq[k] = make_uint4(0u, 0u, 0u, 0u);
            if (i < cnt) {
              const int64_t r = row0 + i;
              if constexpr (kDyImg) {  // 8 bf16 of the tile's own gradient image (same offset as the xhat chunk below)
                if (p.dy_a_img) aq[k] = *reinterpret_cast<const uint4*>(dimg + t128_off(i, cc & 7));
                if (p.dy_b16) {
                  // consecutive CSR rows share their receiver: reuse the previous row's gather instead of asking L2
                  // again (the max-carveout shared memory leaves no L1).  Only a FLAG is set here: copying the register
                  // now would make the issue of every later load wait for the previous row's load to land.
                  const int br = idx_c[i];
                  if (u > 0 && br == idx_c[i - 1]) same |= 1u << k;
                  else {
                    same &= ~(1u << k);
                    cq[k] = *reinterpret_cast<const uint4*>(p.dy_b16 + (int64_t)br * 128 + cc * 8);
                  }
                }
              } else if (p.dy_a) {
                a0[k] = *reinterpret_cast<const float4*>(p.dy_a + r * 128 + cc * 8);
                a1[k] = *reinterpret_cast<const float4*>(p.dy_a + r * 128 + cc * 8 + 4);
              }
              xq[k] = *reinterpret_cast<const uint4*>(ximg + t128_off(i, cc & 7));
            }
          }
        };
        auto compute = [&](int b, int h) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = 32 * b + rg * 2 + u, k = 2 * h + u;
            float dy[8];
            if constexpr (kDyImg) {
              const uint32_t aw[4] = {aq[k].x, aq[k].y, aq[k].z, aq[k].w};
              const uint4 cv = (u > 0 && ((same >> k) & 1u)) ? cq[k - 1] : cq[k];
              const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                dy[2 * e] = __uint_as_float(aw[e] << 16) + __uint_as_float(cw[e] << 16);
                dy[2 * e + 1] = __uint_as_float(aw[e] & 0xffff0000u) + __uint_as_float(cw[e] & 0xffff0000u);
              }
              // aggregate_post_residual: the residual path of the edge latent carries d_ef + d_agg[recv] too - written back
              // over the gradient image (same thread, same 16 bytes it read), where the input kernel's sink picks it up
              if (p.dy_out_img != nullptr && i < cnt)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.dy_out_img) + (size_t)tile * kImg + (cc >> 3) * kTileB +
                                          t128_off(i, cc & 7)) =
                    make_uint4(pack_bf16x2(dy[0], dy[1]), pack_bf16x2(dy[2], dy[3]), pack_bf16x2(dy[4], dy[5]),
                               pack_bf16x2(dy[6], dy[7]));
            } else {
              dy[0] = a0[k].x; dy[1] = a0[k].y; dy[2] = a0[k].z; dy[3] = a0[k].w;
              dy[4] = a1[k].x; dy[5] = a1[k].y; dy[6] = a1[k].z; dy[7] = a1[k].w;
            }
            const uint32_t xw[4] = {xq[k].x, xq[k].y, xq[k].z, xq[k].w};
            float xh[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              xh[2 * e] = bf16_bits_to_float(xw[e] & 0xffffu);
              xh[2 * e + 1] = __uint_as_float(xw[e] & 0xffff0000u);
            }
            float s1 = 0.f, s2 = 0.f, dxh[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              dxh[e] = dy[e] * sc[e];
              s1 += dxh[e];
              s2 = fmaf(dxh[e], xh[e], s2);
              gb[e] += dy[e];
              gs[e] = fmaf(dy[e], xh[e], gs[e]);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
              s1 += __shfl_xor_sync(0xffffffffu, s1, o);
              s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            const float m1 = s1 * (1.f / 128.f), m2 = s2 * (1.f / 128.f);
            const float rs = rstd_c[i];   // rows >= cnt: 0 (and every other factor is 0 too)
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = rs * (dxh[2 * e] - m1 - xh[2 * e] * m2);
              const float b2 = rs * (dxh[2 * e + 1] - m1 - xh[2 * e + 1] * m2);
              w[e] = pack_bf16x2(a, b2);
              dbt[2 * e] += bf16_bits_to_float(w[e] & 0xffffu);
              dbt[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
            }
            st_shared_v4(zb + t128_off(i, cc & 7), w[0], w[1], w[2], w[3]);
          }
        };
        // the first kDepth - 1 batches are in flight while the dZ slot of this tile is still being read
#pragma unroll
        for (int b = 0; b < kDepth - 1; ++b) issue(b, b);
        // next tile: gather rows and rstd (row range from the ring: no dependent global load); ring refill 15 tiles ahead
        int idx_next = 0;
        float rstd_next = 0.f;
        if (lt < kTile && tile + (int)gridDim.x < p.n_tiles) {
          const int2 rn = trs_s[(t_local + 1) & 15];
          if (lt < rn.y) {
            if (p.dy_b16) idx_next = p.b_idx ? p.b_idx[(int64_t)rn.x + lt] : rn.x + lt;
            rstd_next = p.rstd[(int64_t)rn.x + lt];
          }
        }
        const bool refill = lt == kTile && (int64_t)tile + 15 * (int64_t)gridDim.x < p.n_tiles;
        int2 rc_refill = make_int2(0, 0);
        if (refill) rc_refill = tile_range(tile + 15 * gridDim.x);
        mbar_wait(z_empty(zs), ((t_local >> 1) & 1) ^ 1);
        if (lt == 0) trace_ev(p.trace, 0, tn);  // L1: may write
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (b + kDepth - 1 < 4) issue(b + kDepth - 1, (b + kDepth - 1) % kDepth);
          compute(b, b % kDepth);
          if (lt == 0) trace_ev(p.trace, 0, tn);  // Lb: one batch of rows done
        }
        if (lt < kTile) {
          idx_s[((t_local & 1) ^ 1) * kTile + lt] = idx_next;
          rstd_s[((t_local & 1) ^ 1) * kTile + lt] = rstd_next;
        }
        if (refill) trs_s[(t_local + 15) & 15] = rc_refill;
        fence_proxy_async();
        named_bar_sync(2, kHeadThreads);
        if (lt == 0) trace_ev(p.trace, 0, tn);  // L2: top dZ written
        if (lt == 0) {
          mbar_arrive(z_full(zs));
          mbar_arrive(z_empty(zs));  // "column sums done": the head warps keep theirs in registers
        }
      }
    }
    // ---- reduce the per-row-group column sums: the two row groups of a warp by shuffle, then [8][3][128] -> [3][128]
    //      through the head slot the next tile would have used (wait until its last reader is done)
    if (p.head_mode == HEAD_LN) {
      uint32_t n_local = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) ++n_local;
      const uint32_t zs = n_local & 1;
      mbar_wait(z_empty(zs), ((n_local >> 1) & 1) ^ 1);
      float* red_s = reinterpret_cast<float*>(smem + kSmemZ + zs * kImg);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dbt[e] += __shfl_xor_sync(0xffffffffu, dbt[e], 16);
        gs[e] += __shfl_xor_sync(0xffffffffu, gs[e], 16);
        gb[e] += __shfl_xor_sync(0xffffffffu, gb[e], 16);
      }
      if ((lt & 16) == 0) {
        const int wg = lt >> 5;  // head warp 0..7
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          red_s[(wg * 3 + 0) * 128 + cc * 8 + e] = dbt[e];
          red_s[(wg * 3 + 1) * 128 + cc * 8 + e] = gs[e];
          red_s[(wg * 3 + 2) * 128 + cc * 8 + e] = gb[e];
        }
      }
      named_bar_sync(2, kHeadThreads);
      if (lt < 128) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          t0 += red_s[(g * 3 + 0) * 128 + lt];
          t1 += red_s[(g * 3 + 1) * 128 + lt];
          t2 += red_s[(g * 3 + 2) * 128 + lt];
        }
        float* tail = my_partial + (size_t)ns * 16384;
        tail[lt] = t0;                                 // db of the top layer
        tail[(size_t)(ns + 1) * 128 + lt] = t1;        // g_scale
        tail[(size_t)(ns + 1) * 128 + 128 + lt] = t2;  // g_bias
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpM) tmem_dealloc(tmem, 512);
}
}  // namespace chain

// ======================================================================================================
// Input kernel
// ======================================================================================================
namespace input {
constexpr int kEW = 8;                       // epilogue warps: two threads per tile row, each owns 64 accumulator columns
constexpr int kEpi = 32 * kEW;
constexpr int kNP = 2;                       // producer warps: each stages 64 of the 128 rows of a gathered block
constexpr int kWarpP = kEW, kWarpM = kEW + kNP;
constexpr int kThreads = 32 * (kWarpM + 1);  // warps 0..kEW-1 epilogue, then kNP producers, then the MMA warp
constexpr int kZ = 2, kX = 2, kW = 3;
constexpr uint32_t kSmemZ = 0;
constexpr uint32_t kSmemX = kSmemZ + kZ * kImg;
constexpr uint32_t kSmemW = kSmemX + kX * kImg;
constexpr uint32_t kSmemStage = kSmemW + kW * kTileB;
constexpr uint32_t kSmemRp = kSmemStage + kImg;               // tile-local CSR row pointer, 132 ints
constexpr uint32_t kSmemBar = kSmemRp + 132 * 4;
constexpr uint32_t kNumBar = 2 * kZ + 2 * kX + 2 * kW + 3;
constexpr uint32_t kSmemTmem = kSmemBar + 8 * kNumBar;
constexpr uint32_t kSmemTotal = kSmemTmem + 16;
constexpr uint32_t kSmemLaunch = kSmemTotal + 1024;

__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_input_kernel(const InputParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + kSmemBar;
  auto z_full = [&](int s) { return bar0 + 8u * s; };
  auto z_empty = [&](int s) { return bar0 + 8u * (kZ + s); };
  auto x_full = [&](int s) { return bar0 + 8u * (2 * kZ + s); };
  auto x_empty = [&](int s) { return bar0 + 8u * (2 * kZ + kX + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (2 * kZ + 2 * kX + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (2 * kZ + 2 * kX + kW + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * kZ + 2 * kX + 2 * kW);
  const uint32_t acc_empty = acc_full + 8, done_bar = acc_full + 16;
  auto z_slot = [&](int s) { return s_base + kSmemZ + (uint32_t)s * kImg; };
  auto x_slot = [&](int s) { return s_base + kSmemX + (uint32_t)s * kImg; };
  auto w_slot = [&](int s) { return s_base + kSmemW + (uint32_t)s * (uint32_t)kTileB; };
  const uint32_t s_stage = s_base + kSmemStage;
  int* rp_s = reinterpret_cast<int*>(smem + kSmemRp);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemTmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = p.nblk;
  float* my_partial = p.partial + (size_t)blockIdx.x * (size_t)nblk * 16384;

  pdl_trigger();
  if (tid == 0) {
    for (int s = 0; s < kZ; ++s) {
      mbar_init(z_full(s), 1);
      mbar_init(z_empty(s), 1);
    }
    for (int s = 0; s < kX; ++s) {
      mbar_init(x_full(s), 32 * kNP);
      mbar_init(x_empty(s), 1);
    }
    for (int s = 0; s < kW; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kEpi);
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == kWarpM) tmem_alloc(smem_u32(tmem_slot), 512);
  pdl_wait();      // no global memory access above this line
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= kWarpP && warp < kWarpM) {
    // ================================ producer ================================
    uint32_t xc = 0, wc = 0, t_local = 0;
    int tn = 0;
    const int pw = warp - kWarpP;
    const bool lead = lane == 0 && pw == 0;  // issues every bulk copy
    constexpr int RPW = 4 / kNP;             // 32-row groups per producer warp
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t_local) {
      int64_t row0;
      int cnt;
      tile_rows(p.tile_row_start, p.M, tile, row0, cnt);
      if (lead) {
        const uint32_t zs = t_local % kZ;
        mbar_wait(z_empty(zs), ((t_local / kZ) & 1) ^ 1);
        mbar_arrive_expect_tx(z_full(zs), kImg);
        bulk_g2s(z_slot(zs), reinterpret_cast<const uint8_t*>(p.dz0) + (size_t)tile * kImg, kImg, z_full(zs));
      }
      __syncwarp();
      for (int b = 0; b < nblk; ++b) {
        if (lead) {
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t c = wc + kb, s = c % kW;
            mbar_wait(w_empty(s), ((c / kW) & 1) ^ 1);
            mbar_arrive_expect_tx(w_full(s), (uint32_t)kTileB);
            bulk_g2s(w_slot(s), reinterpret_cast<const uint8_t*>(p.wt_img) + (size_t)(b * 2 + kb) * kTileB,
                     (uint32_t)kTileB, w_full(s));
          }
        }
        wc += 2;
        __syncwarp();
        const uint32_t xs = xc % kX;
        if (lead) trace_ev(p.trace, 3, tn);  // P0: before X slot wait
        mbar_wait(x_empty(xs), ((xc / kX) & 1) ^ 1);
        if (lead) trace_ev(p.trace, 3, tn);  // P1: slot free, gather starts
        const __nv_bfloat16* src_base = p.x[b];
        const int32_t* idx = p.idx[b];
        const uint32_t dst = x_slot(xs);
        if (p.x_is_img[b]) {  // the tile's own rows, stored as a tile image: one bulk copy
          if (lead) {
            mbar_arrive_expect_tx(x_full(xs), kImg);
            bulk_g2s(dst, reinterpret_cast<const uint8_t*>(src_base) + (size_t)tile * kImg, kImg, x_full(xs));
          } else {
            mbar_arrive(x_full(xs));
          }
          if (lead) trace_ev(p.trace, 3, tn);  // P2: image copy issued
          ++xc;
          continue;
        }
        int32_t srow[RPW];
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {  // the index loads of this lane are in flight together
          const int r = lane + 32 * (pw * RPW + rr);
          srow[rr] = r < cnt ? (idx ? idx[row0 + r] : (int32_t)(row0 + r)) : 0;
        }
        // One cp.async instruction covers TWO source rows with sixteen lanes each (a 256-byte row = two lines): four lines
        // per instruction instead of the 32 of a lane-per-row mapping (the load/store unit processes one line per cycle
        // and is shared with the epilogue's shared-memory traffic).  The row index is fetched from its lane by a shuffle.
        const int sub = lane >> 4, ch = lane & 15;
        const uint32_t dst_c = dst + (uint32_t)(ch >> 3) * (uint32_t)kTileB;
        const uint8_t* src_c = reinterpret_cast<const uint8_t*>(src_base) + ch * 16;
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
          const int g32 = 32 * (pw * RPW + rr);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int rl = 2 * i + sub;
            const int r = g32 + rl;
            const int64_t src_row = __shfl_sync(0xffffffffu, srow[rr], rl);
            cp_async16(dst_c + t128_off(r, ch & 7), src_c + src_row * 256, r < cnt ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(x_full(xs));
        if (lead) trace_ev(p.trace, 3, tn);  // P2: gather issued
        ++xc;
      }
    }
  } else if (warp == kWarpM) {
    // ================================ MMA issue ================================
    if (lane == 0) {
      const uint32_t idesc_k = umma_idesc(128, 128, false, false);
      const uint32_t idesc_mn = umma_idesc(128, 128, true, true);
      uint32_t xc = 0, wc = 0, t_local = 0, acc_par = 0;
      bool first = true;
      int tn = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++t_local) {
        const uint32_t zs = t_local % kZ;
        trace_ev(p.trace, 1, tn);  // M0: tile start
        mbar_wait(z_full(zs), (t_local / kZ) & 1);
        trace_ev(p.trace, 1, tn);  // M1: dZ0 there
        tc_fence_after();
        // dX of block b is issued BEFORE the weight-gradient MMAs of block b - 1: the epilogue gets its accumulator one
        // gather latency earlier and the X block of b - 1 has the whole dX(b) time more to arrive.
        auto issue_dw = [&](int b) {
          const uint32_t xs = xc % kX;
          mbar_wait(x_full(xs), (xc / kX) & 1);
          trace_ev(p.trace, 1, tn);  // M3: X block there
          fence_proxy_async();
          tc_fence_after();
          const uint32_t d_w = tmem + 128u * (1 + b);
          {
            const uint32_t a_lo = mndesc_lo(x_slot(xs), (uint32_t)kTileB), b_lo = mndesc_lo(z_slot(zs), (uint32_t)kTileB);
            umma_lo(d_w, a_lo, b_lo, idesc_mn, t_local != 0);
#pragma unroll
            for (int ks = 1; ks < 8; ++ks) umma_lo(d_w, a_lo + 128 * ks, b_lo + 128 * ks, idesc_mn, true);
          }
          umma_commit(x_empty(xs));
          ++xc;
        };
        for (int b = 0; b < nblk; ++b) {
          if (!first) {
            mbar_wait(acc_empty, acc_par);
            acc_par ^= 1;
          }
          first = false;
          tc_fence_after();
          for (int kb = 0; kb < 2; ++kb, ++wc) {
            const uint32_t ws = wc % kW;
            mbar_wait(w_full(ws), (wc / kW) & 1);
            tc_fence_after();
            const uint32_t a_lo = kdesc_lo(z_slot(zs) + kb * kTileB), b_lo = kdesc_lo(w_slot(ws));
            umma_lo(tmem, a_lo, b_lo, idesc_k, kb != 0);
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_lo(tmem, a_lo + 2 * k, b_lo + 2 * k, idesc_k, true);
            umma_commit(w_empty(ws));
          }
          umma_commit(acc_full);
          trace_ev(p.trace, 1, tn);  // M2: dX issued
          if (b > 0) issue_dw(b - 1);
        }
        issue_dw(nblk - 1);
        umma_commit(z_empty(zs));
      }
      umma_commit(done_bar);
    }
  } else {
    // ================================ epilogue ================================
    const int row = tid & 127, half = tid >> 7;
    const int c_lo = 2 * half, c_hi = c_lo + 2;  // 32-column accumulator chunks of this thread
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc_par = 0;
    int tn = 0;
    bool stage_store_pending = false;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      int64_t row0;
      int cnt;
      tile_rows(p.tile_row_start, p.M, tile, row0, cnt);
#pragma unroll 1
      for (int b = 0; b < nblk; ++b) {
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E0: block start
        // SINK_ADD_F32 with a source: the tile's fp32 rows (8 per thread) are requested NOW, before the wait for the
        // accumulator and the staging pass, so that their latency is off the critical path (1 CTA/SM: registers abound)
        float4 pr0[8], pr1[8];
        const bool prefetch = p.sink[b] == SINK_ADD_F32 && p.f32_src[b] != nullptr;
        const bool prefetch_img = p.sink[b] == SINK_ADD_IMG && p.img_src[b] != nullptr;
        if (prefetch_img) {  // the tile's own rows of the source image: 8 x 16 bytes per thread, kept as raw bits in pr0
          const int pcc = tid & 15, prg = tid >> 4;
          const uint8_t* src = reinterpret_cast<const uint8_t*>(p.img_src[b]) + (size_t)tile * kImg + (pcc >> 3) * kTileB;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = prg + 16 * u;
            pr0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < cnt) pr0[u] = *reinterpret_cast<const float4*>(src + t128_off(i, pcc & 7));
          }
        }
        if (prefetch) {
          const int pcc = tid & 15, prg = tid >> 4;
          const float* src = p.f32_src[b];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = prg + 16 * u;
            pr0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            pr1[u] = pr0[u];
            if (i < cnt) {
              const int64_t o = (row0 + i) * 128 + pcc * 8;
              pr0[u] = *reinterpret_cast<const float4*>(src + o);
              pr1[u] = *reinterpret_cast<const float4*>(src + o + 4);
            }
          }
        }
        mbar_wait(acc_full, acc_par);
        acc_par ^= 1;
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E1: accumulator full
        tc_fence_after();
        if (tid == 0 && stage_store_pending) {  // a bulk store of the staged tile must have read it
          bulk_wait_read0();
          stage_store_pending = false;
        }
        named_bar_sync(1, kEpi);  // the previous copy-out has finished reading the staging tile
        if (p.sink[b] == SINK_SEGSUM_F32) {  // tile-local CSR row pointer (visible after the next barrier)
          const int n0 = p.tile_node_start[tile], nn = p.tile_node_start[tile + 1] - n0;
          if (tid <= nn) rp_s[tid] = p.row_ptr[n0 + tid] - (int)row0;
          if (tid == 0 && nn == 128) rp_s[128] = p.row_ptr[n0 + 128] - (int)row0;
        }
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          float v[32];
          tmem_ld32(t_lane + c * 32, v);
          const uint32_t sb = s_stage + (c >> 1) * kTileB;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            st_shared_v4(sb + t128_off(row, (c & 1) * 4 + q4), pack_bf16x2(v[q4 * 8], v[q4 * 8 + 1]),
                         pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]), pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]),
                         pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]));
        }
        tc_fence_before();
        mbar_arrive(acc_empty);
        if (p.sink[b] == SINK_STORE_IMG || (p.sink[b] == SINK_ADD_IMG && !prefetch_img)) fence_proxy_async();
        named_bar_sync(1, kEpi);
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E2: staged
        int sink = p.sink[b];
        if (sink == SINK_ADD_IMG) {
          if (prefetch_img) {
            // dst = bf16(src + dX) in the staged tile, in place (rows >= cnt stay zero), then one bulk store
            const int cc = tid & 15, rg = tid >> 4;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int i = rg + 16 * u;
              if (i >= cnt) continue;
              const uint32_t sa = s_stage + (cc >> 3) * kTileB + t128_off(i, cc & 7);
              const uint4 q = ld_shared_v4(sa);
              const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
              const uint32_t sw[4] = {__float_as_uint(pr0[u].x), __float_as_uint(pr0[u].y), __float_as_uint(pr0[u].z),
                                      __float_as_uint(pr0[u].w)};
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o[e] = pack_bf16x2(bf16_bits_to_float(qw[e] & 0xffffu) + bf16_bits_to_float(sw[e] & 0xffffu),
                                   __uint_as_float(qw[e] & 0xffff0000u) + __uint_as_float(sw[e] & 0xffff0000u));
              st_shared_v4(sa, o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async();
            named_bar_sync(1, kEpi);
          }
          sink = SINK_STORE_IMG;
        }
        if (sink == SINK_STORE_IMG) {
          // the staged tile already has the image layout (rows >= cnt are zero): it leaves as one bulk store
          if (tid == 0) {
            bulk_s2g(reinterpret_cast<uint8_t*>(p.bf16_dst[b]) + (size_t)tile * kImg, s_stage, kImg);
            bulk_commit();
            stage_store_pending = true;
          }
        } else if (sink == SINK_STORE_BF16) {
          const int cc = tid & 15, rg = tid >> 4;
#pragma unroll 4
          for (int i = rg; i < cnt; i += kEpi / 16) {
            const uint4 q = ld_shared_v4(s_stage + (cc >> 3) * kTileB + t128_off(i, cc & 7));
            *reinterpret_cast<uint4*>(p.bf16_dst[b] + (row0 + i) * 128 + cc * 8) = q;
          }
        } else if (sink == SINK_ADD_F32) {
          // dst = src + dX; the source rows were prefetched at the top of the block
          const int cc = tid & 15, rg = tid >> 4;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i = rg + 16 * u;
            if (i >= cnt) continue;
            const uint4 q = ld_shared_v4(s_stage + (cc >> 3) * kTileB + t128_off(i, cc & 7));
            const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
            float m[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              m[2 * e] = bf16_bits_to_float(qw[e] & 0xffffu);
              m[2 * e + 1] = __uint_as_float(qw[e] & 0xffff0000u);
            }
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            if (prefetch) {
              a0 = pr0[u];
              a1 = pr1[u];
            }
            const int64_t o = (row0 + i) * 128 + cc * 8;
            *reinterpret_cast<float4*>(p.f32_dst[b] + o) = make_float4(m[0] + a0.x, m[1] + a0.y, m[2] + a0.z, m[3] + a0.w);
            *reinterpret_cast<float4*>(p.f32_dst[b] + o + 4) = make_float4(m[4] + a1.x, m[5] + a1.y, m[6] + a1.z, m[7] + a1.w);
          }
        } else if (sink == SINK_SEGSUM_F32) {
          // adjoint of the receiver gather: deterministic segmented sum over the tile's CSR rows.  In place
          // (src == dst) the flush is a fire-and-forget RED: every address is touched by exactly one thread of one
          // tile, so the result is deterministic and nothing waits on a load.
          const int n0 = p.tile_node_start[tile], nn = p.tile_node_start[tile + 1] - n0;
          const bool in_place = p.f32_src[b] != nullptr;
          float* dst = p.f32_dst[b] + (int64_t)n0 * 128;
          segsum_tile<false, 4>(s_stage, s_base + kSmemRp, nn, tid, nullptr, nullptr,
                             [&](int v, int col0, const float (&a)[8]) {
                               float* d = dst + (int64_t)v * 128 + col0;
                               if (in_place) {
#pragma unroll
                                 for (int e = 0; e < 8; ++e) atomicAdd(d + e, a[e]);
                               } else {
                                 *reinterpret_cast<float4*>(d) = make_float4(a[0], a[1], a[2], a[3]);
                                 *reinterpret_cast<float4*>(d + 4) = make_float4(a[4], a[5], a[6], a[7]);
                               }
                             });
        }
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E3: sink done (this thread)
      }
    }
    if (tid == 0 && stage_store_pending) bulk_wait0();
    mbar_wait(done_bar, 0);
    tc_fence_after();
    named_bar_sync(1, kEpi);   // the staging tile is free: its last bulk store has completed, its last reader is past its sink
    for (int b = 0; b < nblk; ++b) {
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        float v[32];
        tmem_ld32(t_lane + 128u * (1 + b) + c * 32, v);
        store_acc_chunk(s_stage + (uint32_t)warp * 4096u, v,
                        my_partial + (size_t)b * 16384 + (size_t)((warp & 3) * 32) * 128 + c * 32, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpM) tmem_dealloc(tmem, 512);
}
}  // namespace input

// ======================================================================================================
// CUDA-core helpers
// ======================================================================================================
// Block = 32 outputs x 8 part groups: thread (x, y) sums parts y, y+8, ... of output x (coalesced across x),
// then the 8 group sums are added in a fixed order -> deterministic, 8x the memory parallelism of a plain loop.
__global__ void __launch_bounds__(256) reduce_pieces_kernel(const Pieces pieces, int64_t total) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + x;
  int64_t e = i < total ? i : total - 1;
  int k = 0;
  while (k < pieces.n - 1 && e >= pieces.p[k].count) {
    e -= pieces.p[k].count;
    ++k;
  }
  const float* src = pieces.p[k].src + e;
  const int64_t stride = pieces.p[k].stride;
  const int n_parts = pieces.p[k].n_parts;
  float s = 0.f;
#pragma unroll 4
  for (int q = y; q < n_parts; q += 8) s += src[(int64_t)q * stride];
  red[y][x] = s;
  __syncthreads();
  if (y == 0 && i < total) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][x];
    pieces.p[k].dst[e] = t;
  }
}

// The same sum, four outputs per thread (128-bit loads): used when every piece is 16-byte aligned with counts and
// strides that are multiples of 4 - all pieces but the decoder's last layer.  Per output the parts are added in exactly
// the order of the scalar kernel (parts y, y+8, ... then the 8 groups), so both kernels give the same bits.
__global__ void __launch_bounds__(256) reduce_pieces4_kernel(const Pieces pieces, int64_t total4) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 red[8][33];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + x;
  int64_t e = i < total4 ? i : total4 - 1;
  int k = 0;
  while (k < pieces.n - 1 && e >= (pieces.p[k].count >> 2)) {
    e -= pieces.p[k].count >> 2;
    ++k;
  }
  const float4* src = reinterpret_cast<const float4*>(pieces.p[k].src) + e;
  const int64_t stride4 = pieces.p[k].stride >> 2;
  const int n_parts = pieces.p[k].n_parts;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int q = y; q < n_parts; q += 8) {
    const float4 v = src[(int64_t)q * stride4];
    s.x += v.x;
    s.y += v.y;
    s.z += v.z;
    s.w += v.w;
  }
  red[y][x] = s;
  __syncthreads();
  if (y == 0 && i < total4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 v = red[g][x];
      t.x += v.x;
      t.y += v.y;
      t.z += v.z;
      t.w += v.w;
    }
    reinterpret_cast<float4*>(pieces.p[k].dst)[e] = t;
  }
}

// Address of element (row, col) of a [tile][2][16 KB] image, in bytes from the tile base.
__device__ __forceinline__ uint32_t img_off(int row, int col) {
  return (uint32_t)(col >> 6) * (uint32_t)kTileB + t128_off(row, (col & 63) >> 3) + (uint32_t)(col & 7) * 2u;
}

// One CTA per 128-row tile, thread == hidden column c.
__global__ void __launch_bounds__(128) decoder_head_kernel(const float* __restrict__ dout, int od,
                                                           const float* __restrict__ w_last,
                                                           const __nv_bfloat16* __restrict__ h_img, int64_t M,
                                                           __nv_bfloat16* __restrict__ z_img,
                                                           float* __restrict__ partial, const FeatRecipe out_feat,
                                                           const float* __restrict__ val_mask) {
  __shared__ float dout_s[128 * 16];
  __shared__ FeatCol otab[16];
  pdl_trigger();
  pdl_wait();
  const int tile = blockIdx.x, c = threadIdx.x;
  const int64_t row0 = (int64_t)tile * kTile;
  const int cnt = (int)min((int64_t)kTile, M - row0);
  const bool fused_out = out_feat.n > 0;
  if (fused_out) {
    feat_table(out_feat, otab, c, 128);
    __syncthreads();
  }
  for (int i = c; i < 128 * od; i += 128) {
    const int r = i / od;
    float d = r < cnt ? dout[row0 * od + i] : 0.f;
    if (r < cnt) {  // pullback of `inverse_data(o_norm, out) .* val_mask` (src/solve.jl:205-218)
      if (val_mask) d = d * val_mask[row0 * od + i];
      if (fused_out) d = out_vjp(otab[i - r * od], d);
    }
    dout_s[i] = d;
  }
  float wc[16], dw[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    wc[j] = j < od ? w_last[c * od + j] : 0.f;
    dw[j] = 0.f;
  }
  __syncthreads();
  const uint8_t* hb = reinterpret_cast<const uint8_t*>(h_img) + (size_t)tile * kImg;
  uint8_t* zb = reinterpret_cast<uint8_t*>(z_img) + (size_t)tile * kImg;
  float dbz = 0.f;
  for (int r = 0; r < kTile; ++r) {
    const uint32_t off = img_off(r, c);
    __nv_bfloat16 z = __float2bfloat16_rn(0.f);
    if (r < cnt) {
      const float h = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(hb + off));
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < od) {
          const float d = dout_s[r * od + j];
          dot = fmaf(d, wc[j], dot);
          dw[j] = fmaf(h, d, dw[j]);
        }
      z = __float2bfloat16_rn(h > 0.f ? dot : 0.f);
      dbz += __bfloat162float(z);
    }
    *reinterpret_cast<__nv_bfloat16*>(zb + off) = z;
  }
  float* my = partial + (size_t)tile * (size_t)(128 * od + od + 128);
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < od) my[c * od + j] = dw[j];
  if (c < od) {
    float s = 0.f;
    for (int r = 0; r < cnt; ++r) s += dout_s[r * od + c];
    my[128 * od + c] = s;
  }
  my[128 * od + od + c] = dbz;
}

// One CTA per tile.  Phase 1 (thread == output column): dW_0[f][c] partials.  Phase 2 (thread == row):
// d_raw[row][f] = sum_c dZ_0[row][c] W_0[f][c].
__global__ void __launch_bounds__(128) encoder_input_kernel(const __nv_bfloat16* __restrict__ dz0,
                                                            const FeatRecipe feat,
                                                            const int32_t* __restrict__ raw_idx, int F,
                                                            const float* __restrict__ w0, int64_t M,
                                                            const int32_t* __restrict__ trs,
                                                            float* __restrict__ partial, float* __restrict__ d_raw) {
  extern __shared__ float es[];
  float* z_s = es;               // [128][129]
  float* x_s = es + 128 * 129;   // [128][F]
  __shared__ FeatCol ftab[kMaxFeat];
  const int tile = blockIdx.x, t = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  feat_table(feat, ftab, t, 128);
  int64_t row0;
  int cnt;
  tile_rows(trs, M, tile, row0, cnt);
  __syncthreads();  // ftab complete
  const uint8_t* zb = reinterpret_cast<const uint8_t*>(dz0) + (size_t)tile * kImg;
  for (int i = t; i < 128 * 16; i += 128) {  // 16-byte chunks of the image
    const int r = i >> 4, ch = i & 15;
    const uint4 q = *reinterpret_cast<const uint4*>(zb + (ch >> 3) * kTileB + t128_off(r, ch & 7));
    const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      z_s[r * 129 + ch * 8 + 2 * e] = r < cnt ? bf16_bits_to_float(qw[e] & 0xffffu) : 0.f;
      z_s[r * 129 + ch * 8 + 2 * e + 1] = r < cnt ? __uint_as_float(qw[e] & 0xffff0000u) : 0.f;
    }
  }
  for (int i = t; i < 128 * F; i += 128) {
    const int r = i / F, f = i - r * F;
    float v = 0.f;
    if (r < cnt) {
      const int64_t src = raw_idx ? (int64_t)raw_idx[row0 + r] : row0 + r;
      v = feat_eval(ftab[f], src);
    }
    x_s[i] = v;
  }
  __syncthreads();
  float* my = partial + (size_t)tile * (size_t)F * 128;
  for (int f0 = 0; f0 < F; f0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < cnt; ++r) {
      const float z = z_s[r * 129 + t];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (f0 + e < F) acc[e] = fmaf(x_s[r * F + f0 + e], z, acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (f0 + e < F) my[(size_t)(f0 + e) * 128 + t] = acc[e];
  }
  if (d_raw && t < cnt) {
    for (int f = 0; f < F; ++f) {
      float s = 0.f;
      for (int c = 0; c < 128; ++c) s = fmaf(z_s[t * 129 + c], w0[f * 128 + c], s);
      d_raw[(row0 + t) * F + f] = feat_vjp(ftab[f], s);   // pullback of the normaliser (identity recipe: s * 1)
    }
  }
}

__global__ void __launch_bounds__(256) sender_gather_add_kernel(float* __restrict__ d_nf,
                                                                const float* __restrict__ recv_sum,
                                                                const __nv_bfloat16* __restrict__ dxs_img,
                                                                const int32_t* __restrict__ col_ptr,
                                                                const int32_t* __restrict__ csc_pos, int64_t N) {
  pdl_trigger();
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= N) return;
  const int lane = threadIdx.x & 31;
  const int cb = col_ptr[v], ce = col_ptr[v + 1];
  float4 s = *reinterpret_cast<const float4*>(d_nf + v * 128 + lane * 4);
  if (recv_sum) {
    const float4 r = *reinterpret_cast<const float4*>(recv_sum + v * 128 + lane * 4);
    s.x += r.x; s.y += r.y; s.z += r.z; s.w += r.w;
  }
  // lane owns columns 4*lane .. 4*lane+3: half (lane & 1) of the 16-byte chunk (lane >> 1) & 7 of tile lane >> 4
  const uint32_t tsel = (uint32_t)lane >> 4, chunk = ((uint32_t)lane >> 1) & 7u, half = (uint32_t)lane & 1u;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(dxs_img);
  for (int j = cb; j < ce; ++j) {
    const uint32_t pos = (uint32_t)csc_pos[j], t = pos >> 7, r = pos & 127u;
    const uint2 q = *reinterpret_cast<const uint2*>(img + (size_t)t * kImg + tsel * (uint32_t)kTileB + t128_off((int)r, (int)chunk) +
                                                    half * 8u);
    s.x += bf16_bits_to_float(q.x & 0xffffu);
    s.y += __uint_as_float(q.x & 0xffff0000u);
    s.z += bf16_bits_to_float(q.y & 0xffffu);
    s.w += __uint_as_float(q.y & 0xffff0000u);
  }
  *reinterpret_cast<float4*>(d_nf + v * 128 + lane * 4) = s;
}

}  // namespace

int backward_grid(int n_tiles) { return n_tiles < device_sm_count() ? n_tiles : device_sm_count(); }

// ---- debug trace registry (see tc.cuh): compiled in only by `build.py --trace` (-DMGN_ENABLE_TRACE); the product
//      library has no such global and exports no mgn_debug_trace
#ifdef MGN_ENABLE_TRACE
static unsigned long long* g_trace_buf[3] = {nullptr, nullptr, nullptr};
static int g_trace_skip[3] = {0, 0, 0};
void set_trace(unsigned long long* d_buf, int kernel, int skip) {
  if (kernel < 0 || kernel > 2) return;
  g_trace_buf[kernel] = d_buf;
  g_trace_skip[kernel] = skip;
}
unsigned long long* take_trace(int kernel) {
  if (!g_trace_buf[kernel]) return nullptr;
  if (g_trace_skip[kernel]-- > 0) return nullptr;
  unsigned long long* b = g_trace_buf[kernel];
  g_trace_buf[kernel] = nullptr;
  return b;
}
#else
void set_trace(unsigned long long*, int, int) {}
unsigned long long* take_trace(int) { return nullptr; }
#endif

cudaError_t mlp_backward_chain_tc(const ChainParams& p, int* grid_out, cudaStream_t st) {
  static PerDeviceOnce configured;
  cudaError_t ce = configured.run([](int) {
    cudaError_t e = cudaFuncSetAttribute(chain::mlp_bwd_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)chain::kSmemLaunch);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(chain::mlp_bwd_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)chain::kSmemLaunch);
    return e;
  });
  if (ce != cudaSuccess) return ce;
  const int grid = backward_grid(p.n_tiles);
  if (grid_out) *grid_out = grid;
  if (grid == 0) return cudaSuccess;
  ProfScope ps(TAG_TC_MLP_BWD, st);
  ChainParams q = p;
  q.trace = take_trace(1);
  if (q.dy_a != nullptr && q.dy_b16 != nullptr) return cudaErrorInvalidValue;   // the gather exists in the image form only
  if (q.dy_a == nullptr && q.head_mode == HEAD_LN)   // image form (or no fp32 part at all: the last MP step's edge MLP)
    return launch_kernel(p.pdl != 0, chain::mlp_bwd_chain_kernel<true>, dim3(grid), dim3(chain::kThreads), chain::kSmemLaunch, st, q);
  return launch_kernel(p.pdl != 0, chain::mlp_bwd_chain_kernel<false>, dim3(grid), dim3(chain::kThreads), chain::kSmemLaunch, st, q);
}

cudaError_t mlp_backward_input_tc(const InputParams& p, int* grid_out, cudaStream_t st) {
  static PerDeviceOnce configured;
  cudaError_t ce = configured.run([](int) {
    return cudaFuncSetAttribute(input::mlp_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)input::kSmemLaunch);
  });
  if (ce != cudaSuccess) return ce;
  const int grid = backward_grid(p.n_tiles);
  if (grid_out) *grid_out = grid;
  if (grid == 0) return cudaSuccess;
  ProfScope ps(TAG_TC_DW, st);
  InputParams q = p;
  q.trace = take_trace(2);
  return launch_kernel(p.pdl != 0, input::mlp_bwd_input_kernel, dim3(grid), dim3(input::kThreads), input::kSmemLaunch, st, q);
}

cudaError_t reduce_pieces(const Pieces& pieces, cudaStream_t st, bool pdl) {
  int64_t total = 0;
  for (int i = 0; i < pieces.n; ++i) total += pieces.p[i].count;
  if (total == 0) return cudaSuccess;
  bool vec = true;
  for (int i = 0; i < pieces.n; ++i) {
    const Piece& q = pieces.p[i];
    vec = vec && (q.count & 3) == 0 && (q.stride & 3) == 0 && (reinterpret_cast<uintptr_t>(q.src) & 15) == 0 &&
          (reinterpret_cast<uintptr_t>(q.dst) & 15) == 0;
  }
  ProfScope ps(TAG_REDUCE_PARTIALS, st);
  if (vec) return launch_kernel(pdl, reduce_pieces4_kernel, dim3((unsigned)((total / 4 + 31) / 32)), dim3(256), 0, st, pieces, total / 4);
  return launch_kernel(pdl, reduce_pieces_kernel, dim3((unsigned)((total + 31) / 32)), dim3(256), 0, st, pieces, total);
}

cudaError_t decoder_head_bwd(const float* dout, int out_dim, const float* w_last, const __nv_bfloat16* h_img,
                             int n_tiles, int64_t M, __nv_bfloat16* z_img, float* partial, const FeatRecipe& out_feat,
                             const float* val_mask, cudaStream_t st, bool pdl) {
  if (n_tiles == 0) return cudaSuccess;
  ProfScope ps(TAG_TC_MISC, st);
  return launch_kernel(pdl, decoder_head_kernel, dim3(n_tiles), dim3(128), 0, st, dout, out_dim, w_last, h_img, M, z_img, partial,
                       out_feat, val_mask);
}

cudaError_t encoder_input_bwd(const __nv_bfloat16* dz0, const FeatRecipe& feat, const int32_t* raw_idx, int F,
                              const float* w0, int n_tiles, int64_t M, const int32_t* tile_row_start,
                              float* partial, float* d_raw, cudaStream_t st, bool pdl) {
  if (n_tiles == 0) return cudaSuccess;
  const size_t smem = (size_t)(128 * 129 + 128 * F) * sizeof(float);
  static PerDeviceOnce configured;
  cudaError_t ce = configured.run([](int) {
    return cudaFuncSetAttribute(encoder_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)((128 * 129 + 128 * 64) * sizeof(float)));
  });
  if (ce != cudaSuccess) return ce;
  ProfScope ps(TAG_TC_MISC, st);
  return launch_kernel(pdl, encoder_input_kernel, dim3(n_tiles), dim3(128), smem, st, dz0, feat, raw_idx, F, w0, M, tile_row_start,
                       partial, d_raw);
}

cudaError_t sender_gather_add(float* d_nf, const float* recv_sum, const __nv_bfloat16* dxs_img, const int32_t* col_ptr,
                              const int32_t* csc_pos, int64_t N, cudaStream_t st, bool pdl) {
  if (N == 0) return cudaSuccess;
  ProfScope ps(TAG_NODE_GRAD_GATHER, st);
  return launch_kernel(pdl, sender_gather_add_kernel, dim3((unsigned)((N + 7) / 8)), dim3(256), 0, st, d_nf, recv_sum, dxs_img, col_ptr,
                       csc_pos, N);
}

}  // namespace tc
}  // namespace mgn

#ifdef MGN_ENABLE_TRACE
// Debug build only: arm the per-role timestamp trace of the `skip`-th next launch of a tensor-core kernel family
// (0 forward, 1 backward chain, 2 backward input).  d_buf: 4 * 512 u64, zeroed by the caller.
extern "C" int32_t mgn_debug_trace(void* d_buf, int32_t kernel, int32_t skip) {
  mgn::tc::set_trace(static_cast<unsigned long long*>(d_buf), kernel, skip);
  return MGN_OK;
}
#endif
