// gemm_tcgen05.cu — bf16 Linear on the 5th-gen tensor cores (sm_100a): out = epi(A[M,K] . W[N,K]^T).
//
// Persistent, warp-specialised, one CTA per SM:
//   (role -> warp id map: see the kernel; control warps sit above the epilogue warps)
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor 2-D loads of a 128x64 A tile and a BNx64 W tile (both
//                      K-major, 128-byte swizzle) into a STAGES-deep shared-memory ring, mbarrier complete_tx.
//   warp 1 (one lane)  MMA issuer: tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16, fp32 accumulators in
//                      TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of i+1;
//                      tcgen05.commit releases smem slots / publishes the accumulator.
//   warp 2             TMEM allocator (tcgen05.alloc / dealloc).
//   warps 4..19        epilogue: tcgen05.ld 32x32b (one accumulator row per thread, warp w owns TMEM lanes
//                      32*(w%4).., column group (w-4)/4 of four), bias / gather-add / GELU / residual in fp32, then the
//                      32-row x 128-byte sub-tile goes through a per-warp 128B-swizzled staging buffer and leaves as ONE
//                      TMA store (cp.async.bulk.tensor, hardware bounds clipping, no per-element address math); the
//                      residual sub-tile arrives the same way (TMA load into the staging buffer).  A slow element-wise
//                      epilogue covers unaligned outputs / mixed residual dtypes.
// K tails and M/N tails rely on TMA out-of-bounds zero fill + masked stores.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm.h"

namespace anemoi {

namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && it > (1u << 27)) __trap();  // a protocol bug must fail loudly, never hang the device
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// ---- 2-CTA (cta_group::2) helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {  // same smem offset in CTA `cta` of the cluster (shared::cluster address)
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// relaxed: the arrive orders nothing by itself (TMEM reads are ordered by tcgen05.wait::ld + tcgen05.fence, TMA data by complete_tx).
// A .release.cluster arrive compiles to MEMBAR.ALL.CTA + ERRBAR, which in the producer thread waits for the bulk copies it has just
// issued and serialises the pipeline (measured: 2080 cycles per k-block instead of ~512).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are accounted on the LEADER CTA's mbarrier (cluster address), data into this CTA's smem
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {  // arrives on the barrier at this smem offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// L2 prefetch of a tile (no shared-memory destination): lets the producer look further ahead than the shared-memory ring is deep
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace ptx

// -DANEMOI_GEMM_TIMELINE builds only: cycle stamps of CTA 0 at the milestones of a launch (profiles/gemm_timeline.py reads them back through
// anemoi_b200_debug_timeline): where do the ~20 us of fixed cost per launch go?
#ifdef ANEMOI_GEMM_TIMELINE
__device__ long long g_timeline[32];
#define TL(i)                                                   \
  do {                                                          \
    if (blockIdx.x == 0) g_timeline[i] = clock64();             \
  } while (0)
#else
#define TL(i) \
  do {        \
  } while (0)
#endif
constexpr int kBM = 128, kBK = 64;
// Epilogue warps come in groups of four (a warp can only read the TMEM lanes 32*(warp%4)..+31): G groups split the tile's BN columns
// G ways.  Measured (profiles/README.md, v9 A/B): G = 4 (16 epilogue warps, which costs one of the six operand stages) is NOT faster than
// G = 2 - mlp1+GELU 86.5 vs 86.0 us, qkv 87 vs 79 us - so the epilogue is not latency-bound per warp and the ring depth matters more.
#ifndef ANEMOI_GEMM_EPI_GROUPS
#define ANEMOI_GEMM_EPI_GROUPS 2
#endif
constexpr int kEpiGroups = ANEMOI_GEMM_EPI_GROUPS;
constexpr int kEpiWarps = 4 * kEpiGroups;
constexpr int kThreads = 128 + 32 * kEpiWarps;  // epilogue warps + 4 control warps (role map in the kernel)
#ifndef ANEMOI_GEMM_CTRL_HI
#define ANEMOI_GEMM_CTRL_HI 1
#endif
#ifndef ANEMOI_GEMM_L2_AHEAD
#define ANEMOI_GEMM_L2_AHEAD 0
#endif
constexpr int kL2Ahead = ANEMOI_GEMM_L2_AHEAD;
#ifndef ANEMOI_GEMM_RES_PREFETCH
#define ANEMOI_GEMM_RES_PREFETCH 1
#endif
// L2 prefetch of the residual sub-tiles of the NEXT tile (cp.async.bulk.prefetch.tensor, no shared-memory destination), issued by the
// epilogue warp at the start of the current tile.  Measured (profiles/r2/call13_ab_res_l2.txt, same call): projection 42.0 -> 44.0 us,
// MLP-2 80.8 -> 82.8 us - the extra TMA requests cost more than the L2 hits save (the one-round-ahead shared-memory prefetch already
// covers the latency).  Off; kept as a compile-time option.
#ifndef ANEMOI_GEMM_RES_L2_AHEAD
#define ANEMOI_GEMM_RES_L2_AHEAD 0
#endif
constexpr int kSmemLimit = 232448;              // 227 KB: the most one CTA may own

// CG = 1: one CTA computes a 128 x BN tile.  CG = 2 (cta_group::2): a CTA pair computes 256 x BN; each CTA holds 128 accumulator
// rows and loads its own 128 A rows plus HALF of the W rows (BN/2), so operand traffic per flop drops by a third
// (128 flop per L2 byte instead of 85 at BN = 256) and the W tile is read from shared memory once per pair.
// RESD ("deep residual staging", CTA-pair 256 x 256 kernels with a residual): FOUR staging buffers per epilogue warp instead of two, paid
// for with one operand stage (5 instead of 6; 3 KB + 5 x 32 KB + 64 KB = exactly the 227 KB a CTA may own).  The residual sub-tiles then
// arrive two rounds ahead of their use instead of one.  Why: a residual epilogue is a chain of (TMA load -> add -> TMA store) rounds, each
// waiting a DRAM round trip; with one load in flight per warp the projection GEMM spent ~16 k cycles per tile in its epilogue against
// 5.6 k cycles of MMAs (the epilogue, not the tensor pipe, set the pace), and every launch's LAST tile exposed the whole chain.
template <int BN, int CG = 1, bool RESD = false>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;  // 16 KB
  static constexpr int kWBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kWBytes;
  static constexpr int kTmemCols = 2 * BN;  // 128 / 256 / 512: powers of two
  // per epilogue warp: one, two or four 32-row x 64-byte (64B-swizzled) staging buffers for the TMA stores / residual loads
  static constexpr int kStgBufs = RESD ? 4 : ((kEpiWarps == 16 && kStageBytes == 49152) ? 1 : 2);
  static constexpr int kStagingBytes = kEpiWarps * kStgBufs * 2048;
  // layout: [barriers 256 B | bias tile 1 KB | LN column-sum tile 1 KB | residual barriers 512 B | pad to 3 KB][operand ring][epilogue staging]
  // = exactly the 227 KB a CTA may own (BN = 256); relies on the dynamic shared window starting 1024-byte aligned (checked in the kernel)
  static constexpr int kHeadBytes = 3072;
  static constexpr int kResBarOffset = 2304;  // kEpiWarps x kStgBufs mbarriers (8 B each, <= 16 x 4)
  static constexpr int kStagesFit = (kSmemLimit - kHeadBytes - kStagingBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static_assert(kStages >= 3 && 2 * kStages + 4 <= 31, "barrier block: 31 slots + the TMEM address slot");
  static_assert(kResBarOffset + kEpiWarps * kStgBufs * 8 <= kHeadBytes, "residual barriers must fit the head block");
  static constexpr int kSmemBytes = kHeadBytes + kStages * kStageBytes + kStagingBytes;
};

// UMMA shared-memory descriptor for a K-major, 128-byte-swizzled tile (rows of 64 bf16 = 128 B; 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                      // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset between 8-row core-matrix groups
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

// Epilogue mode bits (host-selected, warp-uniform): 0 = slow element-wise path.
constexpr int kEpiFast = 1, kEpiOutF32 = 2, kEpiGelu = 4, kEpiRes = 8, kEpiGather = 16, kEpiLnFold = 32, kEpiStats = 64;
// debug ablations (ANEMOI_B200_GEMM_ABLATE, results are then wrong by construction): skip TMA loads / skip the epilogue body / skip the MMAs
constexpr int kAblNoLoad = 256, kAblNoEpi = 512, kAblNoMma = 1024, kAblNoStore = 2048, kAblNoBias = 4096, kAblNoTmemLd = 8192;
constexpr int kAblMask = kAblNoLoad | kAblNoEpi | kAblNoMma | kAblNoStore | kAblNoBias | kAblNoTmemLd;
// the ablation branches only exist in -DANEMOI_GEMM_ABLATE builds: in the product build they cost ~40 instructions per epilogue round
#ifdef ANEMOI_GEMM_ABLATE
#define ABL(bit) (cx.abl & (bit))
#define ABLK(bit) (epi_mode & (bit))
#else
#define ABL(bit) false
#define ABLK(bit) false
#endif

// Tile order.  ANEMOI_EPI_REVERSE walks the row blocks from the bottom of the matrix up, so that a GEMM which consumes what the previous kernel
// wrote LAST starts where L2 still holds it (the serpentine schedule of the GraphTransformer block, layers/block.py).
__device__ __forceinline__ int tile_at(int t, int num_tiles, bool rev) { return rev ? num_tiles - 1 - t : t; }

struct EpiCtx {
  uint32_t bias_smem;  // 256 floats: the tile's bias slice, shared by the four warps of a column group
  uint32_t tmem_base, stg, res_bar, tfull0, tempty0;  // tempty0: cluster address of the (leader's) accumulator-empty barriers when CG = 2
  int lane, q, grp, num_tiles, tiles_n, first_tile, tile_stride;
  bool rev;  // walk the tiles from the last row block to the first (ANEMOI_EPI_REVERSE)
  int tile_m, row_off;  // rows per tile (128 * CG) and this CTA's row offset inside the tile
  bool remote_empty;
  int abl;  // debug ablation bits
};

// Fast epilogue: templated so that the inner loop has no runtime dtype / flag branches.
// Each warp owns 32 accumulator rows x kColsPerWarp columns and walks them in rounds of 64 bytes of output per row (32 bf16 / 16 fp32
// columns).  The warp's 4 KB staging area is split into TWO 32-row x 64-byte buffers (64-byte-swizzled, the layout of a TMA box with a
// 64-byte inner extent) used alternately: the TMA store of round r drains while round r+1 is computed (cp.async.bulk.wait_group.read 1).
template <int BN, int STG_BUFS, bool OUT_F32, bool GELU, bool RES, bool GATHER, bool LNF, bool STATS = false>
__device__ __forceinline__ void epilogue_fast(const EpiCtx& cx, const EpiParams& ep, const CUtensorMap* tmOut, const CUtensorMap* tmRes) {
  constexpr int kColsPerWarp = BN / kEpiGroups;
  constexpr int CW = OUT_F32 ? 16 : 32;  // columns per 64-byte staging row
  constexpr int ROUNDS = kColsPerWarp / CW;
  constexpr int NG = CW / 8;  // groups of 8 columns per round
  static_assert(kColsPerWarp % CW == 0, "tile width");
  const int lane = cx.lane;
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);  // 64B swizzle: 16-byte chunk index ^= (row >> 1) & 3
  uint32_t rcount = 0;
  int it = 0;
  // Residual pipeline (STG_BUFS >= 2): round r of this warp (rounds are numbered across tiles) works in staging buffer r % NB, which has
  // its own mbarrier; the residual sub-tile of round r + D is TMA-loaded while round r is computed.  NB = 2: D = 1 (the load is issued once
  // the store of round r - 1 has finished reading the other buffer).  NB = 4: D = 2 - two loads in flight per warp, and the buffer being
  // refilled was stored from two rounds ago, so nothing waits on the store just issued.  v8 profile: with the load issued in the round that
  // consumes it the RES GEMMs sat at 14 k cycles per tile against 5.6 k of MMAs; D = 1 brought the projection from 45 to 42 us.
  constexpr int NB = STG_BUFS;
  constexpr int D = NB > 2 ? NB - 2 : 1;
  constexpr bool PREFETCH = RES && STG_BUFS >= 2 && ANEMOI_GEMM_RES_PREFETCH;
  int pf_tile = cx.first_tile, pf_rd = 0;
  uint32_t pf_count = 0;
  auto pf_issue = [&]() {  // lane 0: load the residual sub-tile of the next not-yet-requested round of this warp, if there is one
    if (pf_tile < cx.num_tiles) {
      const int tl = tile_at(pf_tile, cx.num_tiles, cx.rev), pm = tl / cx.tiles_n, pn = tl - pm * cx.tiles_n;
      const uint32_t b = pf_count % NB;
      ptx::mbar_expect_tx(cx.res_bar + 8u * b, 2048);
      ptx::tma_load_2d(cx.stg + b * 2048u, tmRes, cx.res_bar + 8u * b, pn * BN + cx.grp * kColsPerWarp + pf_rd * CW,
                       pm * cx.tile_m + cx.row_off + cx.q * 32);  // OOB rows / columns arrive as zeros
      ++pf_count;
      if (++pf_rd == ROUNDS) pf_rd = 0, pf_tile += cx.tile_stride;
    }
  };
  if constexpr (PREFETCH) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) pf_issue();
    }
  }
  for (int tile = cx.first_tile; tile < cx.num_tiles; tile += cx.tile_stride, ++it) {
    const int tl = tile_at(tile, cx.num_tiles, cx.rev), m_blk = tl / cx.tiles_n, n_blk = tl - m_blk * cx.tiles_n;
    const int as = it & 1;
    const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
    const int row0 = m_blk * cx.tile_m + cx.row_off + cx.q * 32;
    if (ep.bias) {
      // the tile's bias slice goes to shared memory once per column group (one coalesced load) instead of two broadcast
      // global loads per 8 columns per thread (measured: the bias loads were ~25 % of the kernel time)
      asm volatile("bar.sync %0, 128;" ::"r"(1 + cx.grp) : "memory");  // previous tile's readers (the four warps of this column group) are done
      if (cx.q == 0) {
        const int c = n_blk * BN + cx.grp * kColsPerWarp + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c + 4 <= (int)ep.N) {
          b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + c));
        } else {
          if (c < (int)ep.N) b4.x = ep.bias[c];
          if (c + 1 < (int)ep.N) b4.y = ep.bias[c + 1];
          if (c + 2 < (int)ep.N) b4.z = ep.bias[c + 2];
        }
        if (lane * 4 < kColsPerWarp)
          ptx::sts128(cx.bias_smem + (uint32_t)(cx.grp * kColsPerWarp + lane * 4) * 4u, __float_as_uint(b4.x), __float_as_uint(b4.y),
                      __float_as_uint(b4.z), __float_as_uint(b4.w));
      }
      if constexpr (LNF) {
        if (cx.q == 1) {  // the second warp of the group stages the column sums of the gamma-scaled weight
          const int c = n_blk * BN + cx.grp * kColsPerWarp + lane * 4;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c + 4 <= (int)ep.N) {
            b4 = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + c));
          } else {
            if (c < (int)ep.N) b4.x = ep.ln_colsum[c];
            if (c + 1 < (int)ep.N) b4.y = ep.ln_colsum[c + 1];
            if (c + 2 < (int)ep.N) b4.z = ep.ln_colsum[c + 2];
          }
          if (lane * 4 < kColsPerWarp)
            ptx::sts128(cx.bias_smem + 1024u + (uint32_t)(cx.grp * kColsPerWarp + lane * 4) * 4u, __float_as_uint(b4.x), __float_as_uint(b4.y),
                        __float_as_uint(b4.z), __float_as_uint(b4.w));
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + cx.grp) : "memory");
    }
    float ln_mean = 0.f, ln_rstd = 1.f;
    if constexpr (LNF) {
      const float2 st = ln_row_mean_rstd(ep, min((int64_t)row0 + lane, ep.M - 1));  // per tile, not per round
      ln_mean = st.x, ln_rstd = st.y;
    }
    // gathered rows of this thread's edge: byte pointers (a table is fp32 or, ANEMOI_EPI_G*_BF16, bf16)
    const char* g1row = nullptr;
    const char* g2row = nullptr;
    const bool g1h = GATHER && (ep.flags & ANEMOI_EPI_G1_BF16), g2h = GATHER && (ep.flags & ANEMOI_EPI_G2_BF16);
    if constexpr (GATHER) {
      const int64_t r = min((int64_t)row0 + lane, ep.M - 1);
      if (ep.g1) g1row = reinterpret_cast<const char*>(ep.g1) + (int64_t)__ldg(ep.idx1 + r) * ep.ldg * (g1h ? 2 : 4);
      if (ep.g2) g2row = reinterpret_cast<const char*>(ep.g2) + (int64_t)__ldg(ep.idx2 + r) * ep.ldg * (g2h ? 2 : 4);
    }
    // STATS: (mean, M2) of this thread's output row over the current 64-column block, taken from the bf16-ROUNDED values (what the
    // consuming GEMM will read: a constant row must cancel exactly against its column sums).  Sums are accumulated SHIFTED by the block's
    // first element, so that mean^2 >> variance rows lose nothing to cancellation.
    float2 st_s = make_float2(0.f, 0.f), st_q = make_float2(0.f, 0.f);
    float st_x0 = 0.f;
#pragma unroll 1
    for (int rd = 0; rd < ROUNDS; ++rd, ++rcount) {
      const int col_in_tile = cx.grp * kColsPerWarp + rd * CW;
      const int col0 = n_blk * BN + col_in_tile;
      const uint32_t buf = cx.stg + (rcount % NB) * 2048u;
      const uint32_t my_row = buf + lane * 64;
      if constexpr (!PREFETCH) {
        // the TMA store issued from this buffer two rounds ago must have finished READING it
        if (lane == 0 && !ABL(kAblNoStore)) {
          ptx::bulk_wait_read<NB - 1>();
        }
        __syncwarp();
      }
      if constexpr (RES && !PREFETCH) {
        if (lane == 0) {
          ptx::mbar_expect_tx(cx.res_bar + 8u * (rcount % NB), 2048);
          ptx::tma_load_2d(buf, tmRes, cx.res_bar + 8u * (rcount % NB), col0, row0);  // OOB rows / columns arrive as zeros
        }
      }
      if (rd == 0) {
        ptx::mbar_wait(cx.tfull0 + 8u * as, aphase);
        ptx::tc_fence_after();
        if (cx.q == 0 && cx.grp == 0 && lane == 0 && tile == cx.first_tile) TL(6);
      }
      uint32_t r[32];
      if (!ABL(kAblNoTmemLd)) {
        const uint32_t taddr = cx.tmem_base + ((uint32_t)(cx.q * 32) << 16) + (uint32_t)(as * BN + col_in_tile);
        if constexpr (OUT_F32) ptx::tmem_ld_32x32b_x16(taddr, r); else ptx::tmem_ld_32x32b_x32(taddr, r);
        ptx::tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0x3f800000u + j;
      }
      if (rd == ROUNDS - 1) {  // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (cx.remote_empty)
            ptx::mbar_arrive_cluster(cx.tempty0 + 8u * as);
          else
            ptx::mbar_arrive(cx.tempty0 + 8u * as);
        }
      }
      if constexpr (RES) ptx::mbar_wait(cx.res_bar + 8u * (rcount % NB), (rcount / NB) & 1u);
      if (cx.q == 0 && cx.grp == 0 && lane == 0 && tile == cx.first_tile) TL(16 + rd);
      if constexpr (PREFETCH) {
        if (lane == 0) {
          // the buffer of round r + D was last stored from in round r + D - NB: with NB - 1 - D newer stores allowed in flight it is free
          ptx::bulk_wait_read<NB - 1 - D>();
          pf_issue();
        }
      }
      // Per-column additive term of the whole round (bias, or for the folded LayerNorm s * (-rstd * mean) + b), loaded from shared memory
      // BEFORE any of the round's shared-memory stores.  The loads and stores are volatile asm and keep their order, so with the loads
      // inside the 8-column groups every group waited for the previous group's store - its LDS, its FFMA2 / MUFU chain and its STS ran as
      // four serial latency chains per round (SASS: [LDS x4, MUFU x8, F2FP x4, STS] x 4).  Hoisted, the 16 column pairs of a round are
      // independent chains and the SASS interleaves them.  Measured (profiles/r2/call43_ab_epilogue_hoist.txt, same call): the replayed cfg2
      // step 8.33 / 8.36 ms against 8.44 / 8.41 ms, the isolated kernels unchanged to 2 % slower (142 registers instead of 104) - a small
      // gain, i.e. the serialisation was not what paces the GELU rounds either.
      const bool has_bias = ep.bias && !ABL(kAblNoBias);
      float cadd[CW];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
#pragma unroll
        for (int j = 0; j < 8; ++j) cadd[g * 8 + j] = 0.f;
        if (has_bias) {
          const uint32_t ba = cx.bias_smem + (uint32_t)(col_in_tile + g * 8) * 4u;
          const uint4 b0 = ptx::lds128(ba), b1 = ptx::lds128(ba + 16);
          cadd[g * 8 + 0] = __uint_as_float(b0.x), cadd[g * 8 + 1] = __uint_as_float(b0.y), cadd[g * 8 + 2] = __uint_as_float(b0.z);
          cadd[g * 8 + 3] = __uint_as_float(b0.w), cadd[g * 8 + 4] = __uint_as_float(b1.x), cadd[g * 8 + 5] = __uint_as_float(b1.y);
          cadd[g * 8 + 6] = __uint_as_float(b1.z), cadd[g * 8 + 7] = __uint_as_float(b1.w);
        }
        if constexpr (LNF) {
          const uint32_t sa = cx.bias_smem + 1024u + (uint32_t)(col_in_tile + g * 8) * 4u;
          const uint4 s0 = ptx::lds128(sa), s1 = ptx::lds128(sa + 16);
          // rstd * (acc - mean * s) + b = acc * rstd + (s * (-rstd * mean) + b): two packed FMAs per pair of columns
          const float2 m2 = make_float2(-ln_rstd * ln_mean, -ln_rstd * ln_mean);
          const float sv[8] = {__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s0.z), __uint_as_float(s0.w),
                               __uint_as_float(s1.x), __uint_as_float(s1.y), __uint_as_float(s1.z), __uint_as_float(s1.w)};
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float2 c2 = __ffma2_rn(make_float2(sv[j], sv[j + 1]), m2, make_float2(cadd[g * 8 + j], cadd[g * 8 + j + 1]));
            cadd[g * 8 + j] = c2.x, cadd[g * 8 + j + 1] = c2.y;
          }
        }
      }
      // bf16 gather tables: the round's 16-byte pieces (8 columns each) are all requested HERE, before any of the round's shared-memory
      // stores (memory clobbers: a load inside the group loop cannot be hoisted above the previous group's store), so they are in flight
      // together; measured with fp32 tables loaded inside the groups the gather-add epilogue doubled the edge GEMM (552 -> 1 099 us at
      // C = 1024: 8 KB of L2 reads per edge)
      uint4 gq1[GATHER ? NG : 1], gq2[GATHER ? NG : 1];
      if constexpr (GATHER) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int col = col0 + g * 8;
          gq1[g] = gq2[g] = make_uint4(0u, 0u, 0u, 0u);
          if (col + 8 <= (int)ep.N) {
            if (g1h && g1row) gq1[g] = __ldg(reinterpret_cast<const uint4*>(g1row + (int64_t)col * 2));
            if (g2h && g2row) gq2[g] = __ldg(reinterpret_cast<const uint4*>(g2row + (int64_t)col * 2));
          }
        }
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int col = col0 + g * 8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
        if constexpr (LNF) {
          const float2 r2 = make_float2(ln_rstd, ln_rstd);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float2 t2 = __ffma2_rn(make_float2(v[j], v[j + 1]), r2, make_float2(cadd[g * 8 + j], cadd[g * 8 + j + 1]));
            v[j] = t2.x, v[j + 1] = t2.y;
          }
        } else if (has_bias) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float2 t2 = __fadd2_rn(make_float2(v[j], v[j + 1]), make_float2(cadd[g * 8 + j], cadd[g * 8 + j + 1]));
            v[j] = t2.x, v[j + 1] = t2.y;
          }
        }
        if constexpr (GATHER) {
          auto add_bf16 = [&](const uint4& u) {
            v[0] += __uint_as_float(u.x << 16), v[1] += __uint_as_float(u.x & 0xffff0000u), v[2] += __uint_as_float(u.y << 16);
            v[3] += __uint_as_float(u.y & 0xffff0000u), v[4] += __uint_as_float(u.z << 16), v[5] += __uint_as_float(u.z & 0xffff0000u);
            v[6] += __uint_as_float(u.w << 16), v[7] += __uint_as_float(u.w & 0xffff0000u);
          };
          auto add_f32 = [&](const char* row) {
            const float4* p4 = reinterpret_cast<const float4*>(row + (int64_t)col * 4);
            const float4 b0 = __ldg(p4), b1 = __ldg(p4 + 1);
            v[0] += b0.x, v[1] += b0.y, v[2] += b0.z, v[3] += b0.w, v[4] += b1.x, v[5] += b1.y, v[6] += b1.z, v[7] += b1.w;
          };
          if (col + 8 <= (int)ep.N) {
            if (g1row) {
              if (g1h) add_bf16(gq1[g]); else add_f32(g1row);
            }
            if (g2row) {
              if (g2h) add_bf16(gq2[g]); else add_f32(g2row);
            }
          } else if (col < (int)ep.N) {  // ragged last 8-column group of the matrix
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (col + j < (int)ep.N) {
                if (g1row) v[j] += g1h ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g1row)[col + j]) : reinterpret_cast<const float*>(g1row)[col + j];
                if (g2row) v[j] += g2h ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g2row)[col + j]) : reinterpret_cast<const float*>(g2row)[col + j];
              }
            }
          }
        }
        if constexpr (GELU) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float2 g2 = gelu_erf_fast2(make_float2(v[j], v[j + 1]));
            v[j] = g2.x, v[j + 1] = g2.y;
          }
        }
        if constexpr (OUT_F32) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t addr = my_row + ((((uint32_t)(2 * g + hh)) ^ sw) << 4);
            if constexpr (RES) {
              const uint4 rv = ptx::lds128(addr);
              v[4 * hh] += __uint_as_float(rv.x), v[4 * hh + 1] += __uint_as_float(rv.y);
              v[4 * hh + 2] += __uint_as_float(rv.z), v[4 * hh + 3] += __uint_as_float(rv.w);
            }
            ptx::sts128(addr, __float_as_uint(v[4 * hh]), __float_as_uint(v[4 * hh + 1]), __float_as_uint(v[4 * hh + 2]),
                        __float_as_uint(v[4 * hh + 3]));
          }
        } else {
          const uint32_t addr = my_row + ((((uint32_t)g) ^ sw) << 4);
          if constexpr (RES) {
            const uint4 rv = ptx::lds128(addr);
            v[0] += __uint_as_float(rv.x << 16), v[1] += __uint_as_float(rv.x & 0xffff0000u);
            v[2] += __uint_as_float(rv.y << 16), v[3] += __uint_as_float(rv.y & 0xffff0000u);
            v[4] += __uint_as_float(rv.z << 16), v[5] += __uint_as_float(rv.z & 0xffff0000u);
            v[6] += __uint_as_float(rv.w << 16), v[7] += __uint_as_float(rv.w & 0xffff0000u);
          }
          const uint32_t pk[4] = {pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])};
          if constexpr (STATS) {
            if (col + 8 <= (int)ep.N) {  // stats producers have N % 8 == 0: a group of 8 columns is valid or clipped as a whole
              if (g == 0 && !(rd & 1)) st_x0 = __uint_as_float(pk[0] << 16);
              const float2 x02 = make_float2(-st_x0, -st_x0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 w2 = __fadd2_rn(make_float2(__uint_as_float(pk[j] << 16), __uint_as_float(pk[j] & 0xffff0000u)), x02);
                st_s = __fadd2_rn(st_s, w2);
                st_q = __ffma2_rn(w2, w2, st_q);
              }
            }
          }
          ptx::sts128(addr, pk[0], pk[1], pk[2], pk[3]);
        }
      }
      if constexpr (STATS) {
        static_assert(!STATS || (!OUT_F32 && kColsPerWarp % kStatsBlock == 0), "stats blocks are 64 bf16 columns = two rounds");
        if (rd & 1) {  // a 64-column block is complete
          const int64_t row = (int64_t)row0 + lane;
          const int blk = (col0 - CW) / kStatsBlock;
          const int parts = (int)((ep.N + kStatsBlock - 1) / kStatsBlock);
          if (row < ep.M && blk < parts) {
            const float nb = (float)min(kStatsBlock, (int)ep.N - blk * kStatsBlock), sh = st_s.x + st_s.y;
            reinterpret_cast<float2*>(ep.stats_out)[row * parts + blk] = make_float2(st_x0 + sh / nb, fmaxf(st_q.x + st_q.y - sh * sh / nb, 0.f));
          }
          st_s = make_float2(0.f, 0.f), st_q = make_float2(0.f, 0.f);
        }
      }
      ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
      __syncwarp();
      if (lane == 0 && !ABL(kAblNoStore)) {
        ptx::tma_store_2d(tmOut, buf, col0, row0);  // rows >= M / columns >= N are clipped by the hardware
        ptx::bulk_commit();
      }
      if (cx.q == 0 && cx.grp == 0 && lane == 0 && tile == cx.first_tile) TL(7 + rd);
    }
  }
  if (lane == 0) ptx::bulk_wait0();
  if (cx.q == 0 && cx.grp == 0 && lane == 0) TL(11);
}

// Slow element-wise epilogue: any alignment / dtype mix.  One accumulator row per thread, direct global accesses.
template <int BN>
__device__ __noinline__ void epilogue_generic(const EpiCtx& cx, const EpiParams& ep) {
  constexpr int kColsPerWarp = BN / kEpiGroups;
  const int lane = cx.lane;
  int it = 0;
  for (int tile = cx.first_tile; tile < cx.num_tiles; tile += cx.tile_stride, ++it) {
    const int tl = tile_at(tile, cx.num_tiles, cx.rev), m_blk = tl / cx.tiles_n, n_blk = tl - m_blk * cx.tiles_n;
    const int as = it & 1;
    const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
    ptx::mbar_wait(cx.tfull0 + 8u * as, aphase);
    ptx::tc_fence_after();
    const int64_t row = (int64_t)m_blk * cx.tile_m + cx.row_off + cx.q * 32 + lane;
    const bool row_ok = row < ep.M;
    const bool g1h = (ep.flags & ANEMOI_EPI_G1_BF16) != 0, g2h = (ep.flags & ANEMOI_EPI_G2_BF16) != 0;
    const char* g1row = (ep.g1 && row_ok) ? reinterpret_cast<const char*>(ep.g1) + (int64_t)ep.idx1[row] * ep.ldg * (g1h ? 2 : 4) : nullptr;
    const char* g2row = (ep.g2 && row_ok) ? reinterpret_cast<const char*>(ep.g2) + (int64_t)ep.idx2[row] * ep.ldg * (g2h ? 2 : 4) : nullptr;
    const float2 ln_st = (ep.ln_stats && row_ok) ? ln_row_mean_rstd(ep, row) : make_float2(0.f, 1.f);
#pragma unroll 1
    for (int c = 0; c < kColsPerWarp; c += 32) {
      uint32_t r[32];
      const int col_in_tile = cx.grp * kColsPerWarp + c;
      ptx::tmem_ld_32x32b_x32(cx.tmem_base + ((uint32_t)(cx.q * 32) << 16) + (uint32_t)(as * BN + col_in_tile), r);
      ptx::tmem_wait_ld();
      const int64_t col0 = (int64_t)n_blk * BN + col_in_tile;
      if (row_ok) {
#pragma unroll 1
        for (int j = 0; j < 32; ++j) {
          const int64_t n = col0 + j;
          if (n >= ep.N) break;
          float a = __uint_as_float(r[j]);
          if (ep.ln_stats) a = ln_st.y * (a - ln_st.x * ep.ln_colsum[n]);
          if (ep.bias) a += ep.bias[n];
          if (g1row) a += g1h ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g1row)[n]) : reinterpret_cast<const float*>(g1row)[n];
          if (g2row) a += g2h ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g2row)[n]) : reinterpret_cast<const float*>(g2row)[n];
          if (ep.flags & ANEMOI_EPI_GELU) a = gelu_erf_fast(a);
          if (ep.residual) a += load_as_f32(ep.residual, row * ep.ldr + n, ep.r_dtype);
          store_from_f32(ep.out, row * ep.ldo + n, ep.o_dtype, a);
        }
      }
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (cx.remote_empty)
        ptx::mbar_arrive_cluster(cx.tempty0 + 8u * as);
      else
        ptx::mbar_arrive(cx.tempty0 + 8u * as);
    }
  }
}

template <int BN, int CG, bool RESD>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes, int num_kb, int tiles_m,
                             int tiles_n, int epi_mode, const EpiParams ep) {
  using Cfg = GemmCfg<BN, CG, RESD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t bar_base = ptx::smem_u32(smem_raw);
  if (bar_base & 1023u) __trap();  // the 128B-swizzled operand tiles need 1024-byte alignment (no static shared memory in this kernel)
  const uint32_t smem_base = bar_base + Cfg::kHeadBytes;
  const uint32_t staging_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; the TMEM base address slot sits at byte 248; the
  // residual barriers (kStgBufs per epilogue warp) live at kResBarOffset of the head block
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  auto res_bar = [&](int w) { return bar_base + (uint32_t)Cfg::kResBarOffset + 8u * (uint32_t)(w * Cfg::kStgBufs); };  // kStgBufs barriers per warp
  const uint32_t tmem_slot = bar_base + 248u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TL(0);
  // Role map.  The SMSP arbiter favours the highest warp id (B300_MICROARCH.md "hi-wid-first"), so the two single-lane control warps
  // (TMA producer, MMA issuer: ~25 instructions per UMMA, and every cycle they lose delays the tensor pipe) sit ABOVE the epilogue
  // warps: ctrl = 0 producer, 1 MMA issuer, 2 TMEM allocator, 3 spare; ew = epilogue warp index (its TMEM lane quarter is ew % 4 =
  // warp % 4 in both layouts).
#if ANEMOI_GEMM_CTRL_HI
  const int ctrl = warp - kEpiWarps, ew = warp;
#else
  const int ctrl = warp < 4 ? warp : -1, ew = warp - 4;
#endif
  const bool is_epi = ew >= 0 && ew < kEpiWarps;
  const int num_tiles = tiles_m * tiles_n;  // tiles of (128*CG) x BN
  const bool rev = (ep.flags & ANEMOI_EPI_REVERSE) != 0;
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int first_tile = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (ctrl == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    if (epi_mode & kEpiFast) ptx::prefetch_tmap(&tmOut);
    if (epi_mode & kEpiRes) ptx::prefetch_tmap(&tmRes);
  }
  if (ctrl == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);  // CG = 2: the leader's expect_tx covers the bytes of BOTH CTAs' loads; the peer only issues loads
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);
      ptx::mbar_init(tempty_bar(s), kEpiWarps * CG);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int w = 0; w < kEpiWarps * Cfg::kStgBufs; ++w) ptx::mbar_init(res_bar(0) + 8u * w, 1);
    ptx::fence_barrier_init();
  }
  if (ctrl == 2) {
    if constexpr (CG == 2) {
      ptx::tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // PDL: everything above (descriptor prefetch, barrier init, TMEM allocation, cluster sync) may overlap the predecessor's tail; from here
  // on global memory is read (operands, bias, residual) and written
  if (threadIdx.x == 0) TL(1);
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) TL(2);

  if (ctrl == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs of a pair: own A rows, own half of the W rows) =====
      int stage = 0;
      uint32_t phase = 0;
      // L2 prefetch cursor for the A operand (streamed from DRAM): kL2Ahead k-blocks beyond the one being loaded into the ring, across tiles
      int pf_tile = first_tile, pf_kb = 0;
      auto pf_advance = [&]() {
        if (++pf_kb == num_kb) pf_kb = 0, pf_tile += tile_stride;
      };
      if constexpr (kL2Ahead > 0) {
        for (int i = 0; i < kL2Ahead + Cfg::kStages && pf_tile < num_tiles; ++i) pf_advance();
      }
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const int tl = tile_at(tile, num_tiles, rev), m_blk = tl / tiles_n, n_blk = tl - m_blk * tiles_n;
        const int a_row = m_blk * (kBM * CG) + (int)rank * kBM;
        const int w_row = n_blk * BN + (int)rank * (BN / CG);
        for (int kb = 0; kb < num_kb; ++kb) {
          if constexpr (kL2Ahead > 0) {
            if (pf_tile < num_tiles) {
              ptx::tma_prefetch_l2_2d(&tmA, pf_kb * kBK, (tile_at(pf_tile, num_tiles, rev) / tiles_n) * (kBM * CG) + (int)rank * kBM);
              pf_advance();
            }
          }
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (tile == first_tile && kb == 0) TL(3);
          const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
          if (ABLK(kAblNoLoad)) {
            if (CG == 1 || leader) ptx::mbar_arrive(full_bar(stage));
          } else if constexpr (CG == 2) {
            const uint32_t lbar = ptx::mapa(full_bar(stage), 0);  // completion bytes of both CTAs' loads land on the leader's barrier
            if (leader)
              ptx::mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            ptx::tma_load_2d_2sm(a_dst, &tmA, lbar, kb * kBK, a_row);
            ptx::tma_load_2d_2sm(a_dst + Cfg::kABytes, &tmW, lbar, kb * kBK, w_row);
          } else {
            ptx::mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            ptx::tma_load_2d(a_dst, &tmA, full_bar(stage), kb * kBK, a_row);
            ptx::tma_load_2d(a_dst + Cfg::kABytes, &tmW, full_bar(stage), kb * kBK, w_row);
          }
          if (++stage == Cfg::kStages) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (ctrl == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader CTA only when CG = 2) =====
      // (ptxas wraps every tcgen05.mma in an ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY waterfall, ~20 instructions per UMMA, also when
      // the whole warp walks this loop - tried; what matters is that this warp wins the issue arbitration, see the role map)
      // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((kBM * CG) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          if (tile == first_tile && kb == 0) TL(4);
          if (tile == first_tile && kb == num_kb - 1) TL(5);
          const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          const uint64_t a_desc = make_sw128_desc(a_addr);
          const uint64_t b_desc = make_sw128_desc(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            if (ABLK(kAblNoMma)) break;
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle span: +2 in 16-byte units
            if constexpr (CG == 2)
              ptx::umma_bf16_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            else
              ptx::umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees this smem slot (in both CTAs when CG = 2) once the MMAs above have read it
          if constexpr (CG == 2) ptx::umma_commit_2sm(empty_bar(stage)); else ptx::umma_commit(empty_bar(stage));
          if (++stage == Cfg::kStages) stage = 0, phase ^= 1u;
        }
        if constexpr (CG == 2) ptx::umma_commit_2sm(tfull_bar(as)); else ptx::umma_commit(tfull_bar(as));  // accumulator complete
      }
    }
  } else if (is_epi) {
    // ===== epilogue =====
    EpiCtx cx;
    cx.tmem_base = tmem_base, cx.lane = lane, cx.q = ew & 3, cx.grp = ew >> 2;
    cx.stg = staging_base + (uint32_t)ew * (uint32_t)(Cfg::kStgBufs * 2048), cx.res_bar = res_bar(ew);
    cx.tfull0 = tfull_bar(0);
    cx.remote_empty = CG == 2;
    cx.tempty0 = CG == 2 ? ptx::mapa(tempty_bar(0), 0) : tempty_bar(0);
    cx.bias_smem = bar_base + 256;
    cx.tile_m = kBM * CG, cx.row_off = (int)rank * kBM;
    cx.abl = epi_mode & kAblMask;
    cx.num_tiles = num_tiles, cx.tiles_n = tiles_n, cx.first_tile = first_tile, cx.tile_stride = tile_stride;
    cx.rev = rev;
    if (ABLK(kAblNoEpi)) {
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
        ptx::mbar_wait(cx.tfull0 + 8u * (it & 1), (uint32_t)(it >> 1) & 1u);
        ptx::tc_fence_after();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (cx.remote_empty) ptx::mbar_arrive_cluster(cx.tempty0 + 8u * (it & 1)); else ptx::mbar_arrive(cx.tempty0 + 8u * (it & 1));
        }
      }
    } else if (!(epi_mode & kEpiFast)) {
      epilogue_generic<BN>(cx, ep);
    } else {
#define ANEMOI_EPI_CASE(F32, GELU, RES, GATHER)                                                                          \
  case (F32 ? kEpiOutF32 : 0) | (GELU ? kEpiGelu : 0) | (RES ? kEpiRes : 0) | (GATHER ? kEpiGather : 0):                   \
    epilogue_fast<BN, Cfg::kStgBufs, F32, GELU, RES, GATHER, false>(cx, ep, &tmOut, &tmRes);                                              \
    break;
#define ANEMOI_EPI_CASE_LN(F32, GELU)                                                                                     \
  case (F32 ? kEpiOutF32 : 0) | (GELU ? kEpiGelu : 0) | kEpiLnFold:                                                        \
    epilogue_fast<BN, Cfg::kStgBufs, F32, GELU, false, false, true>(cx, ep, &tmOut, &tmRes);                                              \
    break;
      if constexpr (RESD) {  // the deep-staging kernel only exists for the residual epilogues
        switch (epi_mode & ~kEpiFast & ~kAblMask) {
          ANEMOI_EPI_CASE(false, false, true, false)
          ANEMOI_EPI_CASE(false, true, true, false)
          ANEMOI_EPI_CASE(true, false, true, false)
          ANEMOI_EPI_CASE(true, true, true, false)
          case kEpiStats | kEpiRes:
            epilogue_fast<BN, Cfg::kStgBufs, false, false, true, false, false, true>(cx, ep, &tmOut, &tmRes);
            break;
          default: __trap();
        }
      } else
      switch (epi_mode & ~kEpiFast & ~kAblMask) {
        ANEMOI_EPI_CASE(false, false, false, false)
        ANEMOI_EPI_CASE(false, true, false, false)
        ANEMOI_EPI_CASE(false, false, true, false)
        ANEMOI_EPI_CASE(false, true, true, false)
        ANEMOI_EPI_CASE(true, false, false, false)
        ANEMOI_EPI_CASE(true, true, false, false)
        ANEMOI_EPI_CASE(true, false, true, false)
        ANEMOI_EPI_CASE(true, true, true, false)
        ANEMOI_EPI_CASE(false, false, false, true)
        ANEMOI_EPI_CASE(false, true, false, true)
        ANEMOI_EPI_CASE(true, false, false, true)
        ANEMOI_EPI_CASE(true, true, false, true)
        case kEpiStats:  // bf16 output + row statistics for the next LayerNorm (producer side), without / with residual
          epilogue_fast<BN, Cfg::kStgBufs, false, false, false, false, false, true>(cx, ep, &tmOut, &tmRes);
          break;
        case kEpiStats | kEpiRes:
          epilogue_fast<BN, Cfg::kStgBufs, false, false, true, false, false, true>(cx, ep, &tmOut, &tmRes);
          break;
        ANEMOI_EPI_CASE_LN(false, false)
        ANEMOI_EPI_CASE_LN(false, true)
        ANEMOI_EPI_CASE_LN(true, false)
        ANEMOI_EPI_CASE_LN(true, true)
        default: __trap();  // host never selects another combination
      }
#undef ANEMOI_EPI_CASE
#undef ANEMOI_EPI_CASE_LN
    }
  }
  __syncwarp();  // re-converge the single-lane producer / MMA warps before the aligned barrier
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();  // the peer may still signal our barriers / read our smem until here
  if (threadIdx.x == 0) TL(12);
  if (ctrl == 2) {
    ptx::tc_fence_after();
    if constexpr (CG == 2) ptx::tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols); else ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host side: TMA descriptors -----------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int64_t rows, cols, ld;
  int box_rows, es, inner;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && es == o.es && inner == o.inner;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&](int64_t v) { h ^= std::hash<int64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.rows), mix(k.cols), mix(k.ld), mix(k.box_rows), mix(k.es), mix(k.inner);
    return h;
  }
};

// [rows, cols] row-major of element size es (2 = bf16, 4 = fp32), leading dimension ld (elements); box = box_rows x 128 bytes,
// 128-byte swizzle, zero OOB fill on loads / clipping on stores.
static int get_tensor_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out, int es = 2, int inner = 128) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, rows, cols, ld, box_rows, es, inner};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("linear(tcgen05): cuTensorMapEncodeTiled not available from the driver");
    return -2;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(inner / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("linear(tcgen05): cuTensorMapEncodeTiled failed with CUresult %d (ptr %p rows %lld cols %lld ld %lld box %d)", (int)r, ptr,
              (long long)rows, (long long)cols, (long long)ld, box_rows);
    return -2;
  }
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

template <int BN, int CG, bool RESD = false>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmOut, const CUtensorMap& tmRes, int64_t K, int epi_mode,
                  const EpiParams& ep, cudaStream_t s) {
  using Cfg = GemmCfg<BN, CG, RESD>;
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, CG, RESD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel)");
    attr_set = true;
  }
  const int tiles_m = (int)((ep.M + kBM * CG - 1) / (kBM * CG)), tiles_n = (int)((ep.N + BN - 1) / BN);
  const int num_kb = (int)((K + kBK - 1) / kBK);
  const int64_t tiles = (int64_t)tiles_m * tiles_n;
  int sms = num_sms();
  if (CG == 2) sms &= ~1;
  int grid = (int)(tiles * CG < sms ? tiles * CG : sms);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = Cfg::kSmemBytes, cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = (pdl_enabled() && !(ep.flags & ANEMOI_EPI_NOPDL)) ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, CG, RESD>, tmA, tmW, tmOut, tmRes, num_kb, tiles_m, tiles_n, epi_mode, ep);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_bf16_tcgen05_kernel)");
  return launch_status("gemm_bf16_tcgen05_kernel");
}

// cta_group::2 (256x256 per CTA pair) is the default for problems that fill the 74 pairs; ANEMOI_B200_GEMM_CG=1 forces the single-CTA
// kernel (A/B measurements: profiles/README.md).  History: the first 2-CTA version was 2x SLOWER than the single-CTA one because the peer
// producer's `mbarrier.arrive.release.cluster` compiled to MEMBAR.ALL.CTA + ERRBAR, which waits for the bulk copies just issued and
// serialised the ring (2080 cycles per k-block); with a single expect_tx on the leader covering both CTAs' bytes it is 10-25 % faster.
static int env_cta_group() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ANEMOI_B200_GEMM_CG");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

int linear_tcgen05(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t K, const EpiParams& ep, cudaStream_t s) {
  // 2-CTA 256x256 tiles when the problem is large enough to fill the 74 CTA pairs at least twice; else single-CTA 128 x {256,128}
  const int64_t tiles_m = (ep.M + kBM - 1) / kBM;
  int bn = 256, cg = 1;
  if (ep.N <= 128 || tiles_m * ((ep.N + 255) / 256) < 2 * (int64_t)num_sms()) bn = 128;
  if (bn == 256 && env_cta_group() == 2 && ((ep.M + 255) / 256) * ((ep.N + 255) / 256) >= (int64_t)num_sms()) cg = 2;
  {
    // opt-in experiment (ANEMOI_B200_GEMM_PAIR_BN128=1, not yet measured): 256 x 128 CTA-pair tiles for narrow outputs, where 256 x 256
    // tiles quantise badly (N = 512, M = 40 962: 322 tiles on 74 pairs = 5 waves at 87 %; 644 half tiles = 9 half waves at 96.7 %)
    static int pair128 = -1;
    if (pair128 < 0) {
      const char* e = getenv("ANEMOI_B200_GEMM_PAIR_BN128");
      pair128 = (e && e[0] == '1') ? 1 : 0;
    }
    if (pair128 && cg == 2 && ep.N <= 512) bn = 128;
  }
  CUtensorMap tmA, tmW, tmOut, tmRes;
  int rc = get_tensor_map(A, ep.M, K, lda, kBM, &tmA);
  if (rc) return rc;
  rc = get_tensor_map(W, ep.N, K, ldw, bn / cg, &tmW);
  if (rc) return rc;
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int os = ep.o_dtype == ANEMOI_BF16 ? 2 : 4;
  const bool gather = ep.g1 || ep.g2;
  // fast epilogue: TMA-storable output, 16-byte aligned bias / gather rows, residual (if any) of the output dtype and TMA-loadable
  bool fast = a16(ep.out) && (ep.ldo * os) % 16 == 0 && (!ep.bias || a16(ep.bias)) &&
              (!gather || ((!ep.g1 || a16(ep.g1)) && (!ep.g2 || a16(ep.g2)) && ep.ldg % ((ep.flags & (ANEMOI_EPI_G1_BF16 | ANEMOI_EPI_G2_BF16)) ? 8 : 4) == 0 && !ep.residual));
  if (ep.residual) fast = fast && ep.r_dtype == ep.o_dtype && a16(ep.residual) && (ep.ldr * os) % 16 == 0;
  // folded LayerNorm: fast path needs a (non-null) bias tile, no residual / gather, aligned stats and column sums
  if (ep.ln_stats) fast = fast && ep.bias && !ep.residual && !gather && a16(ep.ln_colsum) && (reinterpret_cast<uintptr_t>(ep.ln_stats) & 7) == 0;
  int epi_mode = 0;
  bool fused_stats = false;
  tmOut = tmA, tmRes = tmA;  // placeholders when unused
  if (fast) {
    epi_mode = kEpiFast | (os == 4 ? kEpiOutF32 : 0) | ((ep.flags & ANEMOI_EPI_GELU) ? kEpiGelu : 0) | (ep.residual ? kEpiRes : 0) |
               (gather ? kEpiGather : 0) | (ep.ln_stats ? kEpiLnFold : 0);
    // the epilogue itself produces stats_out for the plain / residual bf16 forms; every other form takes the separate pass below
    fused_stats = ep.stats_out && os == 2 && !(ep.flags & ANEMOI_EPI_GELU) && !gather && !ep.ln_stats && (reinterpret_cast<uintptr_t>(ep.stats_out) & 7) == 0;
    if (fused_stats) epi_mode |= kEpiStats;
    rc = get_tensor_map(ep.out, ep.M, ep.N, ep.ldo, 32, &tmOut, os, 64);
    if (rc) return rc;
    if (ep.residual) {
      rc = get_tensor_map(ep.residual, ep.M, ep.N, ep.ldr, 32, &tmRes, os, 64);
      if (rc) return rc;
    }
  }
  {
    static int abl = -1;
    if (abl < 0) {
      const char* e = getenv("ANEMOI_B200_GEMM_ABLATE");
      abl = e ? atoi(e) : 0;
    }
    epi_mode |= (abl << 8) & kAblMask;
  }
  // opt-in (ANEMOI_B200_GEMM_RESD=1): residual epilogues of the CTA-pair kernel with four staging buffers per warp and two residual loads in
  // flight (GemmCfg RESD)
  static int resd = -1;
  if (resd < 0) {
    const char* e = getenv("ANEMOI_B200_GEMM_RESD");
    resd = (e && e[0] == '1') ? 1 : 0;  // measured (profiles/r2/call23_ab_resd.txt): projection 43.0 vs 43.0 us, MLP-2 80.7 vs 78.7 us - no gain, the
                                        // 5-stage ring costs 2 us: the residual round trips are not what bounds these GEMMs (HBM + the tail wave are)
  }
  if (resd && cg == 2 && bn == 256 && (epi_mode & kEpiFast) && (epi_mode & kEpiRes) && !(epi_mode & (kEpiGather | kEpiLnFold)))
    rc = launch<256, 2, true>(tmA, tmW, tmOut, tmRes, K, epi_mode, ep, s);
  else
  rc = cg == 2     ? (bn == 128 ? launch<128, 2>(tmA, tmW, tmOut, tmRes, K, epi_mode, ep, s) : launch<256, 2>(tmA, tmW, tmOut, tmRes, K, epi_mode, ep, s))
       : bn == 128 ? launch<128, 1>(tmA, tmW, tmOut, tmRes, K, epi_mode, ep, s)
                   : launch<256, 1>(tmA, tmW, tmOut, tmRes, K, epi_mode, ep, s);
  if (rc == 0 && ep.stats_out && !fused_stats) rc = launch_partial_row_stats(ep.out, ep.ldo, ep.o_dtype, ep.M, ep.N, ep.stats_out, s);
  return rc;
}

}  // namespace anemoi

#ifdef ANEMOI_GEMM_TIMELINE
extern "C" __attribute__((visibility("default"))) int anemoi_b200_debug_timeline(long long* host32) {
  return cudaMemcpyFromSymbol(host32, anemoi::g_timeline, sizeof(long long) * 32) == cudaSuccess ? 0 : -2;
}
#endif
