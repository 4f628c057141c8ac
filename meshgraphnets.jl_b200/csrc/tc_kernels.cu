// tcgen05 kernels of the Encode-Process-Decode path (MGN_COMPUTE_BF16), sm_100a only.
//
// mlp_fwd_kernel: one launch runs a whole GraphNetCore MLP (Dense x L -> LayerNorm -> residual ->
// segmented aggregation) for every 128-row tile it owns.  Per tile:
//   * the producer warp stages the layer-0 operand (gathered sender / receiver node latents and the
//     edge latent) as 128B-swizzled K-major tiles with 16-byte cp.async, and streams
//     the pre-swizzled bf16 weight images with 1-D bulk (TMA) copies through a 4-slot ring;
//   * one thread of the MMA warp issues tcgen05.mma (M=128, N=128, K=16) into a 128-column fp32
//     TMEM accumulator;
//   * eight epilogue warps (two threads per tile row) read TMEM, add bias, apply ReLU and write the next
//     layer's A operand straight back into shared memory as bf16 - hidden activations never leave
//     the SM; the last layer's epilogue does LayerNorm, the residual add and the deterministic
//     CSR segmented sum (a11) from shared memory, so the per-edge message never touches HBM.
// mlp_fwd_persist_kernel: the same body for every MLP of an inference pass of a small graph, as stages of one
// cooperative launch with a grid barrier between them.
// Two CTAs are resident per SM (<= 113 KB smem, 128 TMEM columns each) so one CTA's epilogue
// overlaps the other's MMAs.
#include "tc.cuh"
#include "tc_ptx.cuh"

namespace mgn {
namespace tc {
namespace {

// Thread layout, EW = number of epilogue warps (4 or 8): warps 0..EW-1 epilogue, warp EW producer, warp EW+1 MMA
// issue.  With EW = 8 two threads share a tile row: thread (row, half) owns the 64 accumulator columns of its half
// (warps w and w+4 both address TMEM lanes 32*(w%4)..+31), which halves the serial length of every Dense / LayerNorm
// epilogue and doubles the loads in flight of the copy-out.
// Operand ring, RING = number of 16 KB slots: 4 when two CTAs share an SM (large graphs: the other CTA hides the
// staging latency), 10 when the graph has no more tiles than the chip has SMs (one CTA per SM anyway): then ten of the
// twelve layer-0 tiles of an edge MLP (6 gathered A tiles + 6 weight tiles) are in flight at once and the first
// accumulator is ready after one gather latency instead of six.
template <int RING>
struct Lay {
  static constexpr uint32_t kRing = 0;
  static constexpr uint32_t kH = RING * kTileB;                // 2 tiles: hidden activation / xhat
  static constexpr uint32_t kBias = kH + 2 * kTileB;           // [kMaxLayers][128] fp32
  static constexpr uint32_t kLn = kBias + kMaxLayers * 512;    // scale[128], bias[128]
  static constexpr uint32_t kRp = kLn + 1024;                  // tile-local CSR row pointer, 132 ints
  static constexpr uint32_t kStat = kRp + 132 * 4;             // EW = 8: LayerNorm partial sums [2 halves][128 rows] float2
  static constexpr uint32_t kFeat = kStat + 2 * 128 * 8;       // FeatCol[kMaxFeat]: raw-feature recipe per column (IN_RAW)
  static constexpr uint32_t kOutF = kFeat + kMaxFeat * 32;     // FeatCol[16]: inverse_data per output column (FIN_LINEAR)
  static constexpr uint32_t kBar = kOutF + 16 * 32;            // full[RING], empty[RING], acc_full, epi_done
  static constexpr uint32_t kTmem = kBar + (2 * RING + 2) * 8;
  static constexpr uint32_t kTotal = kTmem + 16;
  static constexpr uint32_t kLaunch = kTotal + 1024;           // slack for the 1024 B alignment
};
constexpr int kRingShared = 4, kRingDeep = 10;

// ---------------------------------------------------------------------------------------------------
// Weight packing: fp32 Julia-layout weights -> bf16 128B-swizzled K-major image tiles.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_kernel(const PackTile* __restrict__ tiles, const float* __restrict__ params,
            __nv_bfloat16* __restrict__ images) {
  pdl_trigger();
  pdl_wait();
  const PackTile t = tiles[blockIdx.x];
  const float* W = params + t.w_off;  // [in][out] row-major
  uint8_t* dst = reinterpret_cast<uint8_t*>(images) + (size_t)blockIdx.x * kTileB;
  for (int i = threadIdx.x; i < 1024; i += 256) {
    int n, c;
    if (t.kind == 0) {  // n fastest: W[k][n] is contiguous in n
      n = i & 127;
      c = i >> 7;
    } else {            // chunk fastest: W[n][k] is contiguous in k
      c = i & 7;
      n = i >> 3;
    }
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = t.kb * 64 + c * 8 + j;
      if (t.kind == 0) {
        v[j] = (k < t.in_dim && n < t.out_dim) ? W[(int64_t)k * t.out_dim + n] : 0.f;
      } else {
        const int row = t.nb * 128 + n;
        v[j] = (row < t.in_dim && k < t.out_dim) ? W[(int64_t)row * t.out_dim + k] : 0.f;
      }
    }
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]);
    q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + t128_off(n, c)) = q;
  }
}

// ---------------------------------------------------------------------------------------------------
// Fused MLP forward
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_rows(const FwdCore& p, int tile, int64_t& row0, int& cnt) {
  if (p.tile_row_start) {
    row0 = p.tile_row_start[tile];
    cnt = p.tile_row_start[tile + 1] - (int)row0;
  } else {
    row0 = (int64_t)tile * kTile;
    cnt = (int)min((int64_t)kTile, p.M - row0);
  }
}

// NP = producer warps: 1 when two CTAs share an SM, 4 in the deep-ring variant (each warp stages 32 of the 128 rows, so
// the gather instructions of a tile are issued four times faster - what bounds the latency of a single tile).
// kResImg: the residual input is the bf16 tile image of the latent itself (edge MLPs) - all its rows are requested at
// once as raw 16-byte chunks (same register budget as the double-buffered fp32 rows of the node MLPs).
// Rows of the gather sources for this lane's tile rows (lane + 32 * (pw * RPW + rr)).
template <int RPW>
__device__ __forceinline__ void gather_rows(const FwdCore& p, int64_t row0, int cnt, int lane, int pw, int32_t (&src0)[RPW],
                                            int32_t (&src1)[RPW]) {
#pragma unroll
  for (int rr = 0; rr < RPW; ++rr) {
    const int r = lane + 32 * (pw * RPW + rr);
    const bool ok = r < cnt;
    const int32_t* i0 = p.in_mode == IN_RAW ? p.raw_idx : (p.in_mode == IN_GATHER3 ? p.idx0 : nullptr);
    const int32_t* i1 = p.in_mode == IN_GATHER3 ? p.idx1 : nullptr;
    src0[rr] = ok ? (i0 ? i0[row0 + r] : (int32_t)(row0 + r)) : 0;
    src1[rr] = ok ? (i1 ? i1[row0 + r] : (int32_t)(row0 + r)) : 0;
  }
}

// Per-thread state that survives from one stage of the persistent kernel to the next (registers): the mbarrier phases run
// on across stages (no re-initialisation), and the producer warps fetch the row range and the gather rows of their next
// tile - static graph data - BEFORE the grid barrier, while the other roles still work on the current stage.
template <int RPW>
struct StageCarry {
  uint32_t it = 0;        // producer / MMA thread: ring item counter
  uint32_t par = 0;       // MMA thread: epi_done parity; epilogue threads: acc_full parity
  bool started = false;   // MMA thread: an accumulator round has been issued (the next one waits for epi_done)
  bool have = false;      // producer: row0 / cnt / src* below describe the CTA's first tile of the coming stage
  int32_t row0 = 0, cnt = 0;
  int32_t src0[RPW] = {}, src1[RPW] = {};
};

// The body of one launch / one stage of the persistent kernel.  `smem` is the 1024-byte aligned shared-memory base, `tmem`
// the CTA's 128 accumulator columns; `later_stage`: the mbarriers are live from a previous stage (see StageCarry); `next`:
// the parameters of the coming stage (nullptr: none).
// Ends with a CTA-wide barrier: every role is done and every store of the CTA has been issued (bulk stores completed).
template <int EW, int RING, int NP, bool kResImg = false>
__device__ __forceinline__ void fwd_body(const FwdCore& p, const FeatRecipe& feat, const FeatRecipe& out_feat, uint8_t* smem,
                                         const uint32_t tmem, const bool later_stage, StageCarry<4 / NP>& cy,
                                         const FwdCore* next) {
  using L_ = Lay<RING>;
  constexpr int kRing = RING;
  constexpr uint32_t kSmemRing = L_::kRing, kSmemH = L_::kH, kSmemBias = L_::kBias, kSmemLn = L_::kLn, kSmemRp = L_::kRp,
                     kSmemStat = L_::kStat, kSmemBar = L_::kBar, kSmemTmem = L_::kTmem;
  static_assert(sizeof(FeatCol) == 32, "FeatCol is 32 bytes");
  constexpr int kThreads = 32 * (EW + NP + 1);
  constexpr int kEpi = 32 * EW;        // epilogue threads
  constexpr int kHalves = EW / 4;      // threads per tile row
  constexpr int kWarpP = EW, kWarpM = EW + NP;
  constexpr int RPW = 4 / NP;          // 32-row groups staged by one producer warp
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_ring = s_base + kSmemRing, s_h = s_base + kSmemH;
  float* bias_s = reinterpret_cast<float*>(smem + kSmemBias);
  float* ln_s = reinterpret_cast<float*>(smem + kSmemLn);
  int* rp_s = reinterpret_cast<int*>(smem + kSmemRp);
  FeatCol* ftab = reinterpret_cast<FeatCol*>(smem + L_::kFeat);
  FeatCol* otab = reinterpret_cast<FeatCol*>(smem + L_::kOutF);
  const uint32_t bar0 = s_base + kSmemBar;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kRing + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * kRing), epi_done = bar0 + 8u * (2 * kRing + 1);
  (void)kSmemTmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;

  // ---- setup of a launch / stage: barriers, per-MLP tables
  if (tid == 0 && !later_stage) {
    for (int s = 0; s < kRing; ++s) {
      mbar_init(full_bar(s), 32 * NP);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(epi_done, kEpi);
    fence_mbar_init();
  }
  if (p.in_mode == IN_RAW) feat_table(feat, ftab, tid, kThreads);
  if (p.fin_mode == FIN_LINEAR && out_feat.n > 0) feat_table(out_feat, otab, tid, kThreads);
  for (int i = tid; i < L * 128; i += kThreads) {
    const int l = i >> 7, c = i & 127;
    const int nout = (l == L - 1) ? p.n_out_last : 128;
    bias_s[i] = c < nout ? p.bias[l][c] : 0.f;
  }
  if (p.fin_mode != FIN_LINEAR)
    for (int i = tid; i < 256; i += kThreads) ln_s[i] = i < 128 ? p.ln_scale[i] : p.ln_bias[i - 128];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= kWarpP && warp < kWarpM) {
    // ================================ producer ================================
    uint32_t it = cy.it;
    int tn = 0;
    const int pw = warp - kWarpP;
    const bool lead = lane == 0 && pw == 0;  // issues the bulk copies
    // De-phase the two CTAs of an SM (and neighbouring SMs): all CTAs start together, so without a stagger the
    // whole chip gathers at once and then writes at once, saturating HBM in bursts and idling it in between.
    if (p.stagger_ns > 0) {
      const uint32_t slot = (blockIdx.x >= (gridDim.x + 1) / 2 ? 2u : 0u) + (blockIdx.x & 1u);
      for (uint32_t i = 0; i < slot; ++i) __nanosleep(p.stagger_ns);
    }
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      int64_t row0;
      int cnt;
      // source rows of this lane's 4 tile rows, for both gather index vectors: all index loads of the tile
      // are issued together, so no K-block waits on a dependent index load
      int32_t src0[RPW], src1[RPW];
      if (cy.have && tile == (int)blockIdx.x) {   // fetched before the grid barrier that opened this stage
        row0 = cy.row0;
        cnt = cy.cnt;
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
          src0[rr] = cy.src0[rr];
          src1[rr] = cy.src1[rr];
        }
      } else {
        tile_rows(p, tile, row0, cnt);
        gather_rows<RPW>(p, row0, cnt, lane, pw, src0, src1);
      }
      for (int l = 0; l < L; ++l) {
        for (int kb = 0; kb < p.nkb[l]; ++kb) {
          if (l == 0) {
            // ---- A tile kb of the layer-0 operand
            const int s = it % kRing;
            if (lead) trace_ev(p.trace, 3, tn);  // P0: before slot wait (A tile)
            mbar_wait(empty_bar(s), ((it / kRing) & 1) ^ 1);
            if (lead) trace_ev(p.trace, 3, tn);  // P1: slot free
            const uint32_t dst = s_ring + s * kTileB;
            if (p.in_mode == IN_RAW) {
#pragma unroll
              for (int rr = 0; rr < RPW; ++rr) {
                const int r = lane + 32 * (pw * RPW + rr);
                const bool ok = r < cnt;
                const int64_t src_row = src0[rr];
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                  float v[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const int f = kb * 64 + c * 8 + j;
                    v[j] = (ok && f < p.raw_F) ? feat_eval(ftab[f], src_row) : 0.f;
                  }
                  st_shared_v4(dst + t128_off(r, c), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                               pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                }
              }
              fence_proxy_async();
              mbar_arrive(full_bar(s));
            } else if (p.in_mode == IN_GATHER3 && (kb >> 1) == 2 && p.x2_img != nullptr) {
              // the tile's own rows of the third segment are stored as a tile image: one bulk copy per K-block
              if (lead) {
                mbar_arrive_expect_tx(full_bar(s), (uint32_t)kTileB);
                bulk_g2s(dst, reinterpret_cast<const uint8_t*>(p.x2_img) + (size_t)tile * 2 * kTileB + (size_t)(kb & 1) * kTileB,
                         (uint32_t)kTileB, full_bar(s));
              } else {
                mbar_arrive(full_bar(s));
              }
            } else {
              const __nv_bfloat16* src_base;
              const int seg = kb >> 1;
              int which = 2;  // 0: rows src0, 1: rows src1, 2: identity
              if (p.in_mode == IN_GATHER3) {
                src_base = seg == 2 ? p.x2 : p.x0;
                which = seg;
              } else {
                src_base = seg == 0 ? p.x0 : p.x1;
              }
              // One cp.async instruction covers FOUR source rows with eight lanes each: eight lanes read the 128
              // contiguous bytes (one line) of a row's K-block half, so an instruction touches 4 lines instead of the 32 of
              // a lane-per-row mapping - the load/store unit processes one line per cycle, and it is shared with everything
              // the epilogue warps (of both resident CTAs) do.  The row index lives in the lane that loaded it (lane == row
              // of the 32-row group) and is fetched with a shuffle.
              const int sub = lane >> 3, ch = lane & 7;
              const uint8_t* src_kb = reinterpret_cast<const uint8_t*>(src_base + (kb & 1) * 64) + ch * 16;
#pragma unroll
              for (int rr = 0; rr < RPW; ++rr) {
                const int g32 = 32 * (pw * RPW + rr);
                const int32_t mine = which == 0 ? src0[rr] : (which == 1 ? src1[rr] : (int32_t)min(row0 + g32 + lane, p.M - 1));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int rl = 4 * i + sub;
                  const int r = g32 + rl;
                  const int64_t src_row = __shfl_sync(0xffffffffu, mine, rl);
                  cp_async16(dst + t128_off(r, ch), src_kb + src_row * 256, r < cnt ? 16u : 0u);
                }
              }
              cp_async_arrive_noinc(full_bar(s));
            }
            ++it;
          }
          // ---- weight tile (l, kb)
          const int s = it % kRing;
          mbar_wait(empty_bar(s), ((it / kRing) & 1) ^ 1);
          if (lead) {
            const uint32_t bytes = (l == L - 1 && p.fin_mode == FIN_LINEAR) ? 16u * 128u : (uint32_t)kTileB;
            mbar_arrive_expect_tx(full_bar(s), bytes);
            bulk_g2s(s_ring + s * kTileB, reinterpret_cast<const uint8_t*>(p.wimg[l]) + (size_t)kb * kTileB, bytes,
                     full_bar(s));
          } else {
            mbar_arrive(full_bar(s));
          }
          ++it;
        }
      }
    }
    cy.it = it;
    cy.have = false;
    if (next != nullptr && (int)blockIdx.x < next->n_tiles) {   // static data of the coming stage, ahead of the grid barrier
      int64_t r0;
      int n;
      tile_rows(*next, blockIdx.x, r0, n);
      cy.row0 = (int32_t)r0;
      cy.cnt = n;
      gather_rows<RPW>(*next, r0, n, lane, pw, cy.src0, cy.src1);
      cy.have = true;
    }
  } else if (warp == kWarpM) {
    // ================================ MMA issue ================================
    if (lane == 0) {
      uint32_t it = cy.it, epi_par = cy.par;
      bool first = !cy.started;
      int tn = 0;
      const uint32_t idesc_full = umma_idesc(128, 128, false, false);
      const uint32_t idesc_last = p.fin_mode == FIN_LINEAR ? umma_idesc(128, 16, false, false) : idesc_full;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int l = 0; l < L; ++l) {
          trace_ev(p.trace, 1, tn);  // M0: layer start
          if (!first) {  // TMEM drained and (l > 0) the next A operand written by the epilogue
            mbar_wait(epi_done, epi_par);
            epi_par ^= 1;
          }
          first = false;
          trace_ev(p.trace, 1, tn);  // M1: epilogue of the previous layer done
          tc_fence_after();
          const uint32_t idesc = l == L - 1 ? idesc_last : idesc_full;
          const int ksteps = l == 0 ? p.ksteps0 : 4;
          for (int kb = 0; kb < p.nkb[l]; ++kb) {
            int sa = -1;
            if (l == 0) {
              sa = it % kRing;
              mbar_wait(full_bar(sa), (it / kRing) & 1);
              ++it;
            }
            const int sw = it % kRing;
            mbar_wait(full_bar(sw), (it / kRing) & 1);
            ++it;
            fence_proxy_async();
            tc_fence_after();
            const uint32_t a_lo = kdesc_lo(l == 0 ? s_ring + sa * kTileB : s_h + kb * kTileB);
            const uint32_t b_lo = kdesc_lo(s_ring + sw * kTileB);
            umma_lo(tmem, a_lo, b_lo, idesc, kb != 0);
            for (int k = 1; k < ksteps; ++k) umma_lo(tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, true);
            umma_commit(empty_bar(sw));
            if (l == 0) umma_commit(empty_bar(sa));
          }
          umma_commit(acc_full);
          trace_ev(p.trace, 1, tn);  // M2: layer issued (operands were there)
        }
      }
      cy.it = it;
      cy.par = epi_par;
      cy.started = !first;
    }
  } else {
    // ================================ epilogue (thread == (row, column half)) ================================
    uint32_t acc_par = cy.par;
    bool store_pending = false;
    int tn = 0;
    const int row = tid & 127, half = tid >> 7;  // half == 0 when EW == 4
    const int c_lo = half * (4 / kHalves), c_hi = c_lo + 4 / kHalves;  // 32-column accumulator chunks of this thread
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float2* stat_s = reinterpret_cast<float2*>(smem + kSmemStat);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      int64_t row0;
      int cnt;
      tile_rows(p, tile, row0, cnt);
      // tile-local CSR row pointer of the aggregation: requested now (two dependent loads), stored to shared memory in
      // layer 0's epilogue - the wait for the first accumulator hides them
      int rp_n0 = 0, rp_nn = 0, rp_raw = 0, rp_last = 0;
      if (p.fin_mode == FIN_LN_RESID_AGG) {
        rp_n0 = p.tile_node_start[tile];
        rp_nn = p.tile_node_start[tile + 1] - rp_n0;
        if (tid <= rp_nn) rp_raw = p.row_ptr[rp_n0 + tid];
        if (tid == 0 && rp_nn == 128) rp_last = p.row_ptr[rp_n0 + 128];
      }
      for (int l = 0; l < L; ++l) {
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E0: layer start
        mbar_wait(acc_full, acc_par);
        acc_par ^= 1;
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E1: accumulator full
        tc_fence_after();
        const bool last = l == L - 1;
        if (last && p.fin_mode == FIN_LINEAR) {
          float v[16];
          if (half == 0) tmem_ld16(t_lane, v);  // warp-uniform: a warp belongs to one half
          tc_fence_before();
          mbar_arrive(epi_done);
          if (half == 0 && row < cnt) {
            const bool fused_out = out_feat.n > 0;
            for (int j = 0; j < p.out_dim; ++j) {
              float y = v[j] + bias_s[l * 128 + j];
              if (fused_out) y = out_eval(otab[j], y);                                   // inverse_data (src/solve.jl:205-210)
              if (p.val_mask) y = y * p.val_mask[(row0 + row) * p.out_dim + j];          // .* val_mask (src/solve.jl:218)
              p.out[(row0 + row) * p.out_dim + j] = y;
            }
          }
          continue;
        }
        // the shared activation tile is about to be overwritten: every thread must be done reading the
        // previous tile's copy-out / aggregation, and a pending bulk store must have read it
        // (layers > 0 of an inference pass: the tile's last reader was this layer's MMA, which has completed - no barrier)
        // Two threads per row: the barrier stands in front of the first STORE into the tile (the accumulator is loaded and
        // converted before it, off the wait for the bulk store's read of the tile); `pre_bar`: thread 0's part of it.
        auto pre_bar = [&]() {
          if (tid == 0 && store_pending) {
            bulk_wait_read0();
            store_pending = false;
          }
        };
        auto post_bar = [&]() {
          if (l == 0 && p.fin_mode == FIN_LN_RESID_AGG) {
            // tile-local CSR row pointer -> shared memory (read by the aggregation after two more barriers); the values
            // were requested at the top of the tile
            if (tid <= rp_nn) rp_s[tid] = rp_raw - (int)row0;
            if (tid == 0) {
              if (rp_nn == 128) rp_s[128] = rp_last - (int)row0;
              rp_s[130] = rp_n0;
              rp_s[131] = rp_nn;
            }
          }
        };
        const bool need_bar = l == 0 || p.save_h[0] != nullptr;
        if (kHalves == 1 || last) {   // (the LayerNorm layer synchronises on its statistics exchange anyway)
          if (need_bar) {
            pre_bar();
            if (kHalves == 1 || !last) named_bar_sync(1, kEpi);
          }
          if (kHalves == 1) post_bar();
        }
        // residual rows of the first copy-out batch: issued now so that their latency hides behind the LayerNorm
        // residual rows are fetched 4 per thread at a time into one half of r0/r1 while the other half is consumed
        constexpr int RG = kEpi / 16;  // row groups of the copy-out (8 or 16)
        constexpr int U = 32 / RG;     // rows per thread and 32-row batch (4 or 2)
        float4 r0[kResImg ? 1 : 2 * U], r1[kResImg ? 1 : 2 * U];
        uint4 rq[kResImg ? 4 * U : 1];   // image form: one raw 16-byte chunk per row, every row of the tile (slot U * k + u)
        const int cc = tid & 15, rg = tid >> 4;
        const bool resid = p.fin_mode != FIN_LN && (kResImg ? p.lat_img_in != nullptr : p.lat_in != nullptr);
        // false: only the aggregation is wanted (last MP step's edge latent)
        const bool write_lat = p.lat_out != nullptr || p.lat_img_out != nullptr || p.lat_bf16_out != nullptr;
        auto issue_residual = [&](int k, int h) {  // batch k covers rows 32k + rg + RG*u
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = 32 * k + rg + RG * u;
            if constexpr (kResImg) {
              rq[U * k + u] = make_uint4(0u, 0u, 0u, 0u);
              if (resid && i < cnt)  // an L2 hit: the producer staged this very tile as an operand a few microseconds ago
                rq[U * k + u] = ld_global_cg_v4(reinterpret_cast<const uint8_t*>(p.lat_img_in) + (size_t)tile * 2 * kTileB +
                                                (cc >> 3) * kTileB + t128_off(i, cc & 7));   // L2: the image is written by
                                                                                             // bulk stores, which L1 never sees
            } else {
              r0[U * h + u] = make_float4(0.f, 0.f, 0.f, 0.f);
              r1[U * h + u] = r0[U * h + u];
              if (resid && i < cnt) {
                const int64_t o = (row0 + i) * 128 + cc * 8;
                r0[U * h + u] = *reinterpret_cast<const float4*>(p.lat_in + o);
                r1[U * h + u] = *reinterpret_cast<const float4*>(p.lat_in + o + 4);
              }
            }
          }
        };
        float mean = 0.f, rstd = 1.f;
        bool arrived = false;   // epi_done already signalled (the accumulator was drained into registers)
        if constexpr (kHalves == 2) {
          // Two threads per row: the 64 accumulator columns of this thread are loaded with TWO tcgen05.ld in flight and ONE
          // wait, and stay in registers - the LayerNorm layer reads TMEM once (statistics and normalisation from the same
          // registers) and releases the accumulator to the MMA warp before the statistics are even exchanged.
          const uint32_t bs = s_base + kSmemBias + (uint32_t)(l * 128 + c_lo * 32) * 4u;   // this thread's 64 biases
          auto store_chunk = [&](int c, const uint32_t (&w)[16]) {
            const uint32_t tb = s_h + (c >> 1) * kTileB;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              st_shared_v4(tb + t128_off(row, (c & 1) * 4 + q4), w[4 * q4], w[4 * q4 + 1], w[4 * q4 + 2], w[4 * q4 + 3]);
          };
          if (!last) {
#pragma unroll 1
            for (int cc2 = 0; cc2 < 2; ++cc2) {
              uint32_t r[32], w[16];
              tmem_ld32_issue(t_lane + (c_lo + cc2) * 32, r);
              tmem_ld_wait();
              tmem_regs_fence(r);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 b4 = ld_shared_f4(bs + (uint32_t)(cc2 * 128 + q * 16));
                w[2 * q] = pack_bf16x2_relu(__uint_as_float(r[4 * q]) + b4.x, __uint_as_float(r[4 * q + 1]) + b4.y);
                w[2 * q + 1] = pack_bf16x2_relu(__uint_as_float(r[4 * q + 2]) + b4.z, __uint_as_float(r[4 * q + 3]) + b4.w);
              }
              if (cc2 == 0) {
                if (need_bar) {
                  pre_bar();
                  named_bar_sync(1, kEpi);
                }
                post_bar();
              }
              store_chunk(c_lo + cc2, w);
            }
          } else {
            // LayerNorm statistics: sums of the data shifted by the row's first element (the shifted-data formula keeps the
            // fp32 variance accurate), biased variance; sums over the left and the right 64 columns separately, in column
            // order, then added: the same arithmetic for both thread layouts (mirrored by oracle/mgn_oracle_bf16.py)
            if constexpr (RING == kRingDeep) {
              // One CTA per SM (small graphs, persistent kernel): registers abound, so the thread's 64 columns are loaded
              // ONCE, stay in registers between the statistics and the normalisation, and the accumulator is released to
              // the MMA warp before the statistics are even exchanged.  Same arithmetic, same bits as the two-pass form.
              uint32_t ra[32], rb[32], r_first;
              tmem_ld1_issue(t_lane, r_first);
              tmem_ld32_issue(t_lane + c_lo * 32, ra);
              tmem_ld32_issue(t_lane + (c_lo + 1) * 32, rb);
              tmem_ld_wait();
              tmem_regs_fence(ra);
              tmem_regs_fence(rb);
              const float shift = __uint_as_float(r_first) + bias_s[l * 128];
              float s = 0.f, q = 0.f;
#pragma unroll
              for (int cc2 = 0; cc2 < 2; ++cc2) {
                uint32_t* r = cc2 == 0 ? ra : rb;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  const float4 b4 = ld_shared_f4(bs + (uint32_t)(cc2 * 128 + g * 16));
                  const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float vb = __uint_as_float(r[4 * g + e]) + bb[e];
                    r[4 * g + e] = __float_as_uint(vb);
                    const float d = vb - shift;
                    s += d;
                    q = fmaf(d, d, q);
                  }
                }
              }
              tc_fence_before();
              mbar_arrive(epi_done);   // TMEM is drained: the MMA warp may start the next tile / stage
              arrived = true;
              stat_s[half * 128 + row] = make_float2(s, q);
              named_bar_sync(1, kEpi);   // also the barrier in front of the stores into the tile (thread 0 waited in pre_bar)
              post_bar();
              const float2 a = stat_s[row], b = stat_s[128 + row];
              s = a.x + b.x;
              q = a.y + b.y;
              const float ms = s * (1.f / 128.f);
              mean = shift + ms;
              q = fmaxf(q * (1.f / 128.f) - ms * ms, 0.f) * 128.f;
              rstd = 1.f / sqrtf(q * (1.f / 128.f) + p.eps);
              if (p.save_rstd && row < cnt && half == 0) p.save_rstd[row0 + row] = rstd;
#pragma unroll
              for (int cc2 = 0; cc2 < 2; ++cc2) {
                const uint32_t* r = cc2 == 0 ? ra : rb;
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  w[j] = pack_bf16x2((__uint_as_float(r[2 * j]) - mean) * rstd, (__uint_as_float(r[2 * j + 1]) - mean) * rstd);
                store_chunk(c_lo + cc2, w);
              }
            } else {
            // (the 64 columns do not fit in registers beside the loop state at 96 registers per thread: TMEM is read twice)
            uint32_t r[32], r_first;
            float shift = 0.f, s = 0.f, q = 0.f;
#pragma unroll 1
            for (int cc2 = 0; cc2 < 2; ++cc2) {
              if (cc2 == 0) tmem_ld1_issue(t_lane, r_first);
              tmem_ld32_issue(t_lane + (c_lo + cc2) * 32, r);
              tmem_ld_wait();
              tmem_regs_fence(r);
              if (cc2 == 0) shift = __uint_as_float(r_first) + bias_s[l * 128];
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b4 = ld_shared_f4(bs + (uint32_t)(cc2 * 128 + g * 16));
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float d = __uint_as_float(r[4 * g + e]) + bb[e] - shift;
                  s += d;
                  q = fmaf(d, d, q);
                }
              }
            }
            stat_s[half * 128 + row] = make_float2(s, q);
            named_bar_sync(1, kEpi);   // also the barrier in front of the stores into the tile (thread 0 waited in pre_bar)
            post_bar();
            const float2 a = stat_s[row], b = stat_s[128 + row];
            s = a.x + b.x;
            q = a.y + b.y;
            const float ms = s * (1.f / 128.f);
            mean = shift + ms;
            q = fmaxf(q * (1.f / 128.f) - ms * ms, 0.f) * 128.f;
            rstd = 1.f / sqrtf(q * (1.f / 128.f) + p.eps);
            if (p.save_rstd && row < cnt && half == 0) p.save_rstd[row0 + row] = rstd;
#pragma unroll 1
            for (int cc2 = 0; cc2 < 2; ++cc2) {
              uint32_t w[16];
              tmem_ld32_issue(t_lane + (c_lo + cc2) * 32, r);
              tmem_ld_wait();
              tmem_regs_fence(r);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b4 = ld_shared_f4(bs + (uint32_t)(cc2 * 128 + g * 16));
                w[2 * g] = pack_bf16x2((__uint_as_float(r[4 * g]) + b4.x - mean) * rstd, (__uint_as_float(r[4 * g + 1]) + b4.y - mean) * rstd);
                w[2 * g + 1] = pack_bf16x2((__uint_as_float(r[4 * g + 2]) + b4.z - mean) * rstd, (__uint_as_float(r[4 * g + 3]) + b4.w - mean) * rstd);
              }
              store_chunk(c_lo + cc2, w);
            }
            }
          }
        } else {
        if (last) {  // LayerNorm statistics in ONE extra pass over TMEM (one thread per row: 128 columns do not fit in registers)
          float s = 0.f, q = 0.f, s_lo = 0.f, q_lo = 0.f, shift = 0.f;
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            float v[32];
            tmem_ld32(t_lane + c * 32, v);
            if (c == 0) shift = v[0] + bias_s[l * 128];
            if (c == 2) {
              s_lo = s;
              q_lo = q;
              s = 0.f;
              q = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = v[j] + bias_s[l * 128 + c * 32 + j] - shift;
              s += d;
              q = fmaf(d, d, q);
            }
          }
          s = s_lo + s;
          q = q_lo + q;
          const float ms = s * (1.f / 128.f);
          mean = shift + ms;
          q = fmaxf(q * (1.f / 128.f) - ms * ms, 0.f) * 128.f;
          rstd = 1.f / sqrtf(q * (1.f / 128.f) + p.eps);
          if (p.save_rstd && row < cnt && half == 0) p.save_rstd[row0 + row] = rstd;
        }
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          float v[32];
          tmem_ld32(t_lane + c * 32, v);
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = v[2 * j] + bias_s[l * 128 + c * 32 + 2 * j];
            float b = v[2 * j + 1] + bias_s[l * 128 + c * 32 + 2 * j + 1];
            if (last) {
              a = (a - mean) * rstd;
              b = (b - mean) * rstd;
            } else {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            w[j] = pack_bf16x2(a, b);
          }
          const uint32_t tb = s_h + (c >> 1) * kTileB;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            st_shared_v4(tb + t128_off(row, (c & 1) * 4 + q4), w[4 * q4], w[4 * q4 + 1], w[4 * q4 + 2], w[4 * q4 + 3]);
        }
        }
        fence_proxy_async();
        tc_fence_before();
        __nv_bfloat16* save = last ? p.save_xhat : p.save_h[l];
        if (!last) {
          if (save) {
            named_bar_sync(1, kEpi);
            if (tid == 0) {
              bulk_s2g(reinterpret_cast<uint8_t*>(save) + (size_t)tile * 2 * kTileB, s_h, 2 * kTileB);
              bulk_commit();
              store_pending = true;
            }
          }
          mbar_arrive(epi_done);
          continue;
        }
        // ---- last layer: TMEM is drained -> the MMA warp may start the next tile
        if (!arrived) mbar_arrive(epi_done);
        named_bar_sync(1, kEpi);
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E2: LayerNorm done, xhat staged
        if (save && tid == 0) {
          bulk_s2g(reinterpret_cast<uint8_t*>(save) + (size_t)tile * 2 * kTileB, s_h, 2 * kTileB);
          bulk_commit();
          store_pending = true;
        }
        // ---- first residual batch in flight, then the deterministic segmented sum of the (pre-residual) messages
        //      (rows in ascending CSR slot = ascending original edge id, the CPU scatter order, a11).  The
        //      aggregation is shared-memory bound, so it runs BEFORE the copy-out's burst of global stores fills
        //      the SM's memory pipeline.
        if (tid == 0) trace_ev(p.trace, 0, tn);  // C0: xhat bulk store issued
        // aggregate_post_residual (mgn_model_config): the aggregation sums the UPDATED edge latent ef + m, so the copy-out
        // (in place, in the staged tile) comes first and the segmented sum reads its result - also in the last MP step,
        // whose latent is not written anywhere
        const bool post = p.fin_mode == FIN_LN_RESID_AGG && p.agg_post_residual != 0;
        if (write_lat || post) {
          issue_residual(0, 0);
          if constexpr (kResImg) issue_residual(1, 0);   // half of the tile's rows hide behind the aggregation (as many
                                                         // registers as one fp32 batch); the other half is requested at the
                                                         // start of the copy-out, two batches ahead of its use
        }
        if (tid == 0) trace_ev(p.trace, 0, tn);  // C1: first residual batch issued
        __nv_bfloat16* const agg = p.fin_mode == FIN_LN_RESID_AGG ? p.agg_bf16 + (int64_t)rp_s[130] * 128 : nullptr;
        auto agg_flush = [&](int v, int col0, const float (&a)[8]) {
          uint4 o;
          o.x = pack_bf16x2(a[0], a[1]);
          o.y = pack_bf16x2(a[2], a[3]);
          o.z = pack_bf16x2(a[4], a[5]);
          o.w = pack_bf16x2(a[6], a[7]);
          *reinterpret_cast<uint4*>(agg + (int64_t)v * 128 + col0) = o;
        };
        if (p.fin_mode == FIN_LN_RESID_AGG && !post)
          segsum_tile<true, (EW == 8 ? 4 : 3)>(s_h, s_base + kSmemRp, rp_s[131], tid, ln_s, ln_s + 128, agg_flush);
        if (tid == 0) trace_ev(p.trace, 0, tn);  // C: aggregation done
        const bool img_out = (write_lat && p.lat_img_out != nullptr) || post;   // the new latent is formed in the staged tile
        if (img_out) {
          // the bf16 shadow of the new latent overwrites the xhat tile in place (same thread, same 16 bytes) and leaves as
          // one bulk store: every reader of xhat (aggregation, the xhat save) must be done first
          if (tid == 0 && store_pending) {
            bulk_wait_read0();
            store_pending = false;
          }
          named_bar_sync(1, kEpi);
        }
        // ---- copy-out: m = xhat * scale + bias ; residual ; fp32 master + bf16 shadow (coalesced), four batches of
        //      32 rows, the residual rows of batch k+1 in flight while batch k is written
        if (write_lat || post) {
          if constexpr (kResImg) {
            issue_residual(2, 0);
            issue_residual(3, 0);
          }
          const uint32_t ln_a = s_base + kSmemLn + (uint32_t)cc * 32u;
          const float4 s0 = ld_shared_f4(ln_a), s1 = ld_shared_f4(ln_a + 16u), b0 = ld_shared_f4(ln_a + 512u), b1 = ld_shared_f4(ln_a + 528u);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float bi[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int h = k & 1;
            if (!kResImg && k + 1 < 4) issue_residual(k + 1, h ^ 1);
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int i = 32 * k + rg + RG * u;
              if (i >= cnt) {
                if (img_out) st_shared_v4(s_h + (cc >> 3) * kTileB + t128_off(i, cc & 7), 0u, 0u, 0u, 0u);
                continue;
              }
              const uint4 xq = ld_shared_v4(s_h + (cc >> 3) * kTileB + t128_off(i, cc & 7));
              const uint32_t xw[4] = {xq.x, xq.y, xq.z, xq.w};
              float m[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&xw[j]);
                m[2 * j] = fmaf(__low2float(hh), sc[2 * j], bi[2 * j]);
                m[2 * j + 1] = fmaf(__high2float(hh), sc[2 * j + 1], bi[2 * j + 1]);
              }
              if constexpr (kResImg) {
                const uint4 q = rq[U * k + u];
                m[0] += __uint_as_float(q.x << 16); m[1] += __uint_as_float(q.x & 0xffff0000u);
                m[2] += __uint_as_float(q.y << 16); m[3] += __uint_as_float(q.y & 0xffff0000u);
                m[4] += __uint_as_float(q.z << 16); m[5] += __uint_as_float(q.z & 0xffff0000u);
                m[6] += __uint_as_float(q.w << 16); m[7] += __uint_as_float(q.w & 0xffff0000u);
              } else {
                const float4 a0 = r0[U * h + u], a1 = r1[U * h + u];
                m[0] += a0.x; m[1] += a0.y; m[2] += a0.z; m[3] += a0.w;
                m[4] += a1.x; m[5] += a1.y; m[6] += a1.z; m[7] += a1.w;
              }
              const int64_t o = (row0 + i) * 128 + cc * 8;
              if (p.lat_out) {
                *reinterpret_cast<float4*>(p.lat_out + o) = make_float4(m[0], m[1], m[2], m[3]);
                *reinterpret_cast<float4*>(p.lat_out + o + 4) = make_float4(m[4], m[5], m[6], m[7]);
              }
              uint4 bq;
              bq.x = pack_bf16x2(m[0], m[1]);
              bq.y = pack_bf16x2(m[2], m[3]);
              bq.z = pack_bf16x2(m[4], m[5]);
              bq.w = pack_bf16x2(m[6], m[7]);
              if (img_out) st_shared_v4(s_h + (cc >> 3) * kTileB + t128_off(i, cc & 7), bq.x, bq.y, bq.z, bq.w);
              else if (p.lat_bf16_out) *reinterpret_cast<uint4*>(p.lat_bf16_out + o) = bq;
            }
          }
        }
        if (img_out) {
          fence_proxy_async();
          named_bar_sync(1, kEpi);
          if (tid == 0 && p.lat_img_out != nullptr) {
            bulk_s2g(reinterpret_cast<uint8_t*>(p.lat_img_out) + (size_t)tile * 2 * kTileB, s_h, 2 * kTileB);
            bulk_commit();
            store_pending = true;
          }
          if (post)   // rows >= cnt of the tile are zero and belong to no node
            segsum_tile<false, (EW == 8 ? 4 : 3)>(s_h, s_base + kSmemRp, rp_s[131], tid, nullptr, nullptr, agg_flush);
        }
        if (tid == 0) trace_ev(p.trace, 2, tn);  // E3: copy-out + aggregation done (this thread)
      }
    }
    if (tid == 0 && store_pending) bulk_wait0();
    cy.par = acc_par;
  }
  tc_fence_before();
  __syncthreads();
}

// One launch = one MLP.
template <int EW, int RING, int NP, bool kResImg = false>
__global__ void __launch_bounds__(32 * (EW + NP + 1), RING == kRingShared ? 2 : 1) mlp_fwd_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Lay<RING>::kTmem);
  const int warp = threadIdx.x >> 5;
  // no global memory access before pdl_wait(): with programmatic dependent launch the TMEM allocation overlaps the tail
  // of the previous kernel
  pdl_trigger();
  if (warp == EW + NP) tmem_alloc(smem_u32(tmem_slot), 128);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  StageCarry<4 / NP> cy;
  fwd_body<EW, RING, NP, kResImg>(p, p.feat, p.out_feat, smem, tmem, false, cy, nullptr);
  if (warp == EW + NP) tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------------
// Persistent forward pass for graphs with no more tiles than SMs: ONE cooperative launch runs every MLP of the pass
// (encoders, mps x (edge, node), decoder) as a stage, with a grid-wide barrier where a launch boundary used to be - a
// stage of such a graph is one tile's dependency chain per CTA (10 - 12 us), and a third of an inference pass was the
// launch gaps and kernel prologues between the stages.  The stages' parameters live in kernel-parameter space (a FwdCore
// each, indexed by the stage: constant-bank loads); ring and accumulator barriers keep their phases across stages.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();   // every global write of the CTA happens-before thread 0's release below
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int seen;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (clock64() - t0 > 4000000000LL) __trap();   // a CTA that never arrived: fail the launch instead of hanging
    } while (seen < target);
  }
  __syncthreads();
}

template <int EW, int RING, int NP>
__global__ void __launch_bounds__(32 * (EW + NP + 1), 1) mlp_fwd_persist_kernel(const __grid_constant__ PersistParams pp) {
  using L_ = Lay<RING>;
  constexpr int kThreads = 32 * (EW + NP + 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L_::kTmem);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == EW + NP) tmem_alloc(smem_u32(tmem_slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  StageCarry<4 / NP> cy;
  for (int si = 0; si < pp.n_stages; ++si) {
    const FwdCore& p = pp.stage[si];   // kernel-parameter space, indexed: the fields stay constant-bank loads
    const FeatRecipe& feat = pp.feat[si == 1 ? 1 : 0];
    const FwdCore* next = si + 1 < pp.n_stages ? &pp.stage[si + 1] : nullptr;
#ifdef MGN_ENABLE_TRACE
    const bool stamp = pp.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    if (stamp) pp.dbg[3 * si] = globaltimer_ns();
#endif
    if (p.lat_img_in != nullptr) fwd_body<EW, RING, NP, true>(p, feat, pp.feat[2], smem, tmem, si > 0, cy, next);
    else fwd_body<EW, RING, NP, false>(p, feat, pp.feat[2], smem, tmem, si > 0, cy, next);
#ifdef MGN_ENABLE_TRACE
    if (stamp) pp.dbg[3 * si + 1] = globaltimer_ns();
#endif
    if (si + 1 < pp.n_stages) grid_barrier(pp.sync, (unsigned int)(si + 1) * gridDim.x);
#ifdef MGN_ENABLE_TRACE
    if (stamp) pp.dbg[3 * si + 2] = globaltimer_ns();
#endif
  }
  if (warp == EW + NP) tmem_dealloc(tmem, 128);
}

}  // namespace

cudaError_t pack_weights(const ModelImages& im, const float* params, __nv_bfloat16* images, cudaStream_t st, bool pdl) {
  if (im.n_tiles == 0) return cudaSuccess;
  ProfScope ps(TAG_TC_PACK, st);
  return launch_kernel(pdl, pack_kernel, dim3(im.n_tiles), dim3(256), 0, st, im.d_tiles, params, images);
}

// Variant selection comes from the model handle (FwdParams::epi_warps / deep_ring / stagger_ns, frozen at
// mgn_model_create); the opt-in shared-memory limits are set once per device.
// The opt-in shared-memory limits of every forward kernel, once per device.
static cudaError_t configure_forward() {
  static PerDeviceOnce configured;
  return configured.run([](int) {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f, uint32_t bytes) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    };
    set((const void*)mlp_fwd_kernel<4, kRingShared, 1, false>, Lay<kRingShared>::kLaunch);
    set((const void*)mlp_fwd_kernel<8, kRingShared, 1, false>, Lay<kRingShared>::kLaunch);
    set((const void*)mlp_fwd_kernel<8, kRingDeep, 4, false>, Lay<kRingDeep>::kLaunch);
    set((const void*)mlp_fwd_kernel<4, kRingShared, 1, true>, Lay<kRingShared>::kLaunch);
    set((const void*)mlp_fwd_kernel<8, kRingShared, 1, true>, Lay<kRingShared>::kLaunch);
    set((const void*)mlp_fwd_kernel<8, kRingDeep, 4, true>, Lay<kRingDeep>::kLaunch);
    set((const void*)mlp_fwd_persist_kernel<8, kRingDeep, 4>, Lay<kRingDeep>::kLaunch);
    return e;
  });
}

cudaError_t mlp_forward_tc(const FwdParams& p, cudaStream_t st) {
  if (p.n_tiles == 0) return cudaSuccess;
  cudaError_t ce = configure_forward();
  if (ce != cudaSuccess) return ce;
  const int n_sm = device_sm_count();
  const int grid = p.n_tiles < 2 * n_sm ? p.n_tiles : 2 * n_sm;
  ProfScope ps(TAG_TC_MLP_FWD, st);
  FwdParams q = p;
  q.trace = take_trace(0);
  if (grid <= n_sm) q.stagger_ns = 0u;
  const bool pdl = p.pdl != 0;
  const bool img = p.lat_img_in != nullptr;   // residual from the bf16 latent image (edge MLPs)
  if (img && p.lat_in != nullptr) return cudaErrorInvalidValue;
  if (p.epi_warps == 4)
    return img ? launch_kernel(pdl, mlp_fwd_kernel<4, kRingShared, 1, true>, dim3(grid), dim3(32 * 6), Lay<kRingShared>::kLaunch, st, q)
               : launch_kernel(pdl, mlp_fwd_kernel<4, kRingShared, 1, false>, dim3(grid), dim3(32 * 6), Lay<kRingShared>::kLaunch, st, q);
  if (p.deep_ring && p.n_tiles <= n_sm)
    return img ? launch_kernel(pdl, mlp_fwd_kernel<8, kRingDeep, 4, true>, dim3(grid), dim3(32 * 13), Lay<kRingDeep>::kLaunch, st, q)
               : launch_kernel(pdl, mlp_fwd_kernel<8, kRingDeep, 4, false>, dim3(grid), dim3(32 * 13), Lay<kRingDeep>::kLaunch, st, q);
  return img ? launch_kernel(pdl, mlp_fwd_kernel<8, kRingShared, 1, true>, dim3(grid), dim3(32 * 10), Lay<kRingShared>::kLaunch, st, q)
             : launch_kernel(pdl, mlp_fwd_kernel<8, kRingShared, 1, false>, dim3(grid), dim3(32 * 10), Lay<kRingShared>::kLaunch, st, q);
}

bool forward_persist_ok(int max_tiles, int n_stages) {
  return max_tiles > 0 && max_tiles <= device_sm_count() && n_stages <= kMaxStages;
}

cudaError_t mlp_forward_persist_tc(const PersistParams& pp, int max_tiles, cudaStream_t st) {
  cudaError_t ce = configure_forward();
  if (ce != cudaSuccess) return ce;
  ce = cudaMemsetAsync(pp.sync, 0, sizeof(unsigned int), st);
  if (ce != cudaSuccess) return ce;
  ProfScope ps(TAG_TC_MLP_FWD, st);
  PersistParams q = pp;
  q.dbg = take_trace(0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(max_tiles);
  cfg.blockDim = dim3(32 * 13);
  cfg.dynamicSmemBytes = Lay<kRingDeep>::kLaunch;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // every CTA resident before any runs: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, mlp_fwd_persist_kernel<8, kRingDeep, 4>, q);
}

}  // namespace tc
}  // namespace mgn
