// Inline-PTX building blocks for the sm_100a tensor-core path: mbarrier, bulk (TMA) copies,
// cp.async, tcgen05 (TMEM alloc / mma / commit / ld) and UMMA descriptors.
//
// Shared-memory operand tiles use ONE physical format ("T128"): 128 rows x 128 bytes, 16-byte
// chunk c of row r stored at chunk position c ^ (r & 7) (the 128B swizzle), 1024-byte aligned.
// Read as a K-major operand the rows are the M/N index and the 64 bf16 of a row are K; read as
// an MN-major operand the rows are K and the 64 bf16 are M/N.  Same bytes, different descriptor.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace mgn {
namespace tc {

constexpr uint32_t kTileBytes = 16384;  // one T128 tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---- proxy / tcgen05 fences ---------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Register reallocation between warpgroups (setmaxnreg): every warp of a warpgroup (4 consecutive warps) must execute
// the same call.  N: multiple of 8 in [24, 256].
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- copies -------------------------------------------------------------------------------------
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 1-D bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 16-byte cp.async (LDGSTS); src_bytes == 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// The mbarrier receives one arrival when all cp.async issued so far by this thread have landed.
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- UMMA ---------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in
// [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in
// [46,48), layout type SWIZZLE_128B (= 2) in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major T128 tile: 8-row groups are 1024 B apart; a UMMA_K = 16 step advances 32 B.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int kstep) {
  return umma_desc(tile_addr + kstep * 32, 16, 1024);
}
// MN-major operand made of T128 tiles (each 64 M/N wide, `slab_bytes` apart): 8 K-rows are
// 1024 B apart; a UMMA_K = 16 step advances 16 rows = 2048 B.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, uint32_t slab_bytes, int kstep) {
  return umma_desc(tile_addr + kstep * 2048, slab_bytes, 1024);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 A/B.
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// The same with the descriptors split into 32-bit halves.  Only the start-address field (low 14 bits of the low word,
// in 16-byte units) changes between the MMAs of a K loop, so the issuing thread builds a descriptor ONCE per operand tile
// and adds a constant per step - the single thread that issues the MMAs is latency bound on exactly this arithmetic
// (a descriptor rebuilt from the address costs ~8 dependent integer operations per operand and MMA, more than the
// 64 cycles the tensor pipe needs for the MMA itself).  Shared-memory addresses are < 256 KB: the field cannot overflow.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t kdesc_lo(uint32_t tile_addr) {      // K-major T128 tile; K step: + 2 (32 bytes)
  return ((tile_addr >> 4) & 0x3FFFu) | ((16u >> 4) << 16);
}
__device__ __forceinline__ uint32_t mndesc_lo(uint32_t tile_addr, uint32_t slab_bytes) {   // MN-major; K step: + 128 (2048 bytes)
  return ((tile_addr >> 4) & 0x3FFFu) | (((slab_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ void umma_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"((uint32_t)accumulate), "r"(kDescHi)
      : "memory");
}
// All MMAs issued so far by this thread arrive on the mbarrier when they complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (thread i of the warp gets lane i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load WITHOUT the wait: several loads can be in flight before one tmem_ld_wait().  The wait names the
// destination registers as read-write operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld1_issue(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_regs_fence(uint32_t (&r)[32]) {  // compiler-only: pins the registers below the wait
  asm volatile(""
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One fp32 column (lane i of the warp's 32 lanes -> thread i).
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __uint_as_float(r);
}

// ---- debug trace ----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// One thread per role calls this; n is the role's private event counter.
__device__ __forceinline__ void trace_ev(unsigned long long* buf, int role, int& n) {
#ifdef MGN_ENABLE_TRACE   // debug build only (build.py --trace): the product kernels carry no trace instructions
  if (buf != nullptr && blockIdx.x == 0 && n < 512) buf[role * 512 + n++] = globaltimer_ns();
#else
  (void)buf; (void)role; (void)n;
#endif
}

// ---- T128 addressing ----------------------------------------------------------------------------
// Byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a T128 tile.
__device__ __forceinline__ uint32_t t128_off(int row, int chunk) {
  return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// bf16x2(max(lo, 0), max(hi, 0)) in one conversion (F2FP with the ReLU modifier)
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// 16 bytes of fp32 from a SHARED address (LDS.128; a generic load of the same broadcast costs twice the wavefronts)
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_global_cg_v4(const void* p) {   // cached in L2 only
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// ---- deterministic segmented sum over a staged tile ---------------------------------------------------
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// Sums the rows of a 128 x 128 bf16 tile image (two T128 tiles at shared address `tile`) over the CSR segments
// of the tile's nodes: node v (local index) owns rows [rp[v], rp[v+1]) (rp at shared address `rp`).  Rows are
// added in ascending order (= ascending original edge id, the CPU scatter order); no atomics.
// 128 threads: thread t owns one 16-byte chunk of columns (8 columns: chunk t & 7 of tile (t >> 3) & 1) and one
// eighth of the node list, so a row costs ONE 128-bit shared load per thread (quarter-warps read 128 contiguous
// bytes: conflict-free) - the shared-memory pipe is shared with the co-resident CTA's global traffic, so the
// number of memory instructions, not the arithmetic, is what this loop pays for.  Node-major: up to 4 rows of
// a node are in flight together, then added in order.  kAffine: every element is first mapped to
// fma(x, scale, bias) (the LayerNorm affine of the message; sc / bi at shared float pointers).
// flush(v_local, col0, sums[8]) once per node.
// kGroupsLog2: the node list is split over 2^kGroupsLog2 thread groups of 16 (3: 128 threads, 4: 256 threads).
template <bool kAffine, int kGroupsLog2 = 3, class Flush>
__device__ __forceinline__ void segsum_tile(uint32_t tile, uint32_t rp, int n_nodes, int t, const float* sc_s,
                                            const float* bi_s, Flush flush) {
  const uint32_t chunk = (uint32_t)t & 7u, tsel = ((uint32_t)t >> 3) & 1u;
  const int grp = t >> 4;
  const int vb = (n_nodes * grp) >> kGroupsLog2, ve = (n_nodes * (grp + 1)) >> kGroupsLog2;
  const int col0 = (int)(tsel * 64u + chunk * 8u);
  const uint32_t base = tile + tsel * 16384u;
  float sc[8], bi[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = 1.f;
    bi[e] = 0.f;
  }
  if (kAffine) {
    const float4 s0 = ld_shared_f4(smem_u32(sc_s + col0)), s1 = ld_shared_f4(smem_u32(sc_s + col0 + 4));
    const float4 b0 = ld_shared_f4(smem_u32(bi_s + col0)), b1 = ld_shared_f4(smem_u32(bi_s + col0 + 4));
    sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
    bi[0] = b0.x; bi[1] = b0.y; bi[2] = b0.z; bi[3] = b0.w; bi[4] = b1.x; bi[5] = b1.y; bi[6] = b1.z; bi[7] = b1.w;
  }
  int jb = vb < ve ? (int)ld_shared_u32(rp + 4u * vb) : 0;
  for (int v = vb; v < ve; ++v) {
    const int je = (int)ld_shared_u32(rp + 4u * (v + 1));
    float a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = 0.f;
    for (int j0 = jb; j0 < je; j0 += 4) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t j = (uint32_t)(j0 + u);
        q[u] = make_uint4(0u, 0u, 0u, 0u);
        if ((int)j < je) q[u] = ld_shared_v4(base + j * 128u + ((chunk ^ (j & 7u)) << 4));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j0 + u < je) {
          const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x0 = __uint_as_float(w[e] << 16), x1 = __uint_as_float(w[e] & 0xffff0000u);
            if (kAffine) {
              x0 = fmaf(x0, sc[2 * e], bi[2 * e]);
              x1 = fmaf(x1, sc[2 * e + 1], bi[2 * e + 1]);
            }
            a[2 * e] += x0;
            a[2 * e + 1] += x1;
          }
        }
      }
    }
    flush(v, col0, a);
    jb = je;
  }
}

}  // namespace tc
}  // namespace mgn
