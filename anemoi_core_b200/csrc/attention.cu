// attention.cu — GraphTransformer edge-softmax attention forward, one warp per destination node over the cached
// CSR (dst-sorted edges).  Replaces layers/conv.py:103-147 and triton/gt.py:81-179.
//
// Layout: q/out rows [H*Ch] per dst, k/v rows per src.  A lane owns VEC = H*Ch/32 contiguous channels, so a warp
// reads each gathered k / v row as one fully coalesced run of 16-byte vectors; a head spans LPH = Ch/VEC lanes
// and its score is finished with log2(LPH) xor-shuffles.  Online softmax in fp32 (running max / sum per head,
// as gt.py:121-160).
//
// Fused lin_edge (block.py:623-635): the reference materialises eproj = W_e a_e + b_e as an [E, H*Ch] tensor that
// is written by a GEMM and re-read here.  Both uses are linear in eproj, so per dst node and head
//     q.(k + W a + b) = q.k + (W^T q).a + q.b          (q.b is constant over the edges of a dst => drops out of the softmax)
//     sum_e alpha_e (v + W a_e + b) = sum_e alpha_e v + W (sum_e alpha_e a_e) + b
// i.e. an edge_dim-long dot product per edge and two [Ch x edge_dim] products per dst node; the [E, H*Ch] stream
// (the largest tensor of the layer) never exists.  Edge attributes stay fp32.
#include <cstdlib>

#include "attention.h"
#include "common.cuh"

namespace anemoi {

// ---- slab kernel ---------------------------------------------------------------------------------------------
// Work item = (range of kNodesPerRange consecutive dst nodes, channel slab).  A slab is NCH x 512 bytes of a node row:
// lane l owns the 16-byte chunks (j*32 + l), j < NCH, so every gathered k / v row segment is read by ONE fully
// coalesced 512-byte warp load per chunk.  A head spans LPH = Ch / EPC consecutive lanes of a chunk (EPC = elements per
// 16 bytes); a lane therefore serves NCH heads, one per chunk.
// Memory-level parallelism (the kernel is latency / L2 bound, not FLOP bound): colptr of the whole range and the src ids
// of the range's (contiguous) edge run are fetched with coalesced loads ahead of use and handed out by shuffles; the
// gathers of kBatch = 4 edges (k and v, all chunks) are issued back to back before any arithmetic; q / qw of the next
// node are prefetched during the current one.
// MODE 1: materialised eproj (may be null = no edge term).
// MODE 2: fused lin_edge.  The caller supplies qw[d,h,:] = W_e,h^T q[d,h,:] (folded into the q GEMM) and receives
//         abar[d,h,:] = sum_e alpha_e a_e (W_e is applied to it inside the projection GEMM); the attribute index space is
//         split over the LPH lanes of a head (lane o owns attributes o, o+LPH, ...), so the per-edge cost of the edge term
//         is ceil(D/LPH) loads and 2*ceil(D/LPH) FMAs per lane and rides on the score's shuffle reduction.
#ifndef ATTN_NODES_PER_RANGE
#define ATTN_NODES_PER_RANGE 8
#endif
#ifndef ATTN_BATCH
#define ATTN_BATCH 4
#endif
#ifndef ATTN_MINBLOCKS
#define ATTN_MINBLOCKS 3
#endif
[[maybe_unused]] constexpr int kNodesPerRange = ATTN_NODES_PER_RANGE;
[[maybe_unused]] constexpr int kBatch = ATTN_BATCH;

template <typename T>
struct ChunkT {
  static constexpr int EPC = 16 / (int)sizeof(T);
  __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[EPC]) {
    if constexpr (sizeof(T) == 4) {
      f[0] = __uint_as_float(u.x), f[1] = __uint_as_float(u.y), f[2] = __uint_as_float(u.z), f[3] = __uint_as_float(u.w);
    } else {
      f[0] = __uint_as_float(u.x << 16), f[1] = __uint_as_float(u.x & 0xffff0000u);
      f[2] = __uint_as_float(u.y << 16), f[3] = __uint_as_float(u.y & 0xffff0000u);
      f[4] = __uint_as_float(u.z << 16), f[5] = __uint_as_float(u.z & 0xffff0000u);
      f[6] = __uint_as_float(u.w << 16), f[7] = __uint_as_float(u.w & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float (&f)[EPC]) {
    if constexpr (sizeof(T) == 4) {
      return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    } else {
      return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
  }
};

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <typename T, int NCH, int LPH, int MODE>
__global__ void __launch_bounds__(128, ATTN_MINBLOCKS) gt_attention_slab_kernel(const AttnParams p, int n_slabs, int active_lanes) {
  using CT = ChunkT<T>;
  constexpr int EPC = CT::EPC;
  constexpr int NA = (kMaxEdgeDim + LPH - 1) / LPH;  // attribute slots per lane (upper bound; `na` of them are live)
  constexpr int SLAB = NCH * 32 * EPC;               // channels per slab
  const int lane = threadIdx.x & 31;
  const bool active = lane < active_lanes;  // narrow rows (< 512 bytes): the idle lanes shadow lane 0 and never store
  const int lane_eff = active ? lane : 0;
  const int sub = lane & (LPH - 1);
  const int na = MODE == 2 ? (p.edge_dim + LPH - 1) / LPH : 0;
  const char* __restrict__ qp = reinterpret_cast<const char*>(p.q);
  const char* __restrict__ kp = reinterpret_cast<const char*>(p.k);
  const char* __restrict__ vp = reinterpret_cast<const char*>(p.v);
  const char* __restrict__ ep = reinterpret_cast<const char*>(p.e);
  const int64_t ldq_b = p.ldq * (int64_t)sizeof(T), ldk_b = p.ldk * (int64_t)sizeof(T), ldv_b = p.ldv * (int64_t)sizeof(T);
  const int64_t lde_b = p.lde_proj * (int64_t)sizeof(T);
  const int64_t n_ranges = (p.n_dst + kNodesPerRange - 1) / kNodesPerRange;
  const int64_t items = n_ranges * n_slabs;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float qscale = p.scale * 1.4426950408889634f;  // scores in the log2 domain -> ex2

  for (int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < items; item += warps_total) {
    const int slab = (int)(item % n_slabs);
    const int64_t n0 = (item / n_slabs) * kNodesPerRange;
    const int nn = (int)min((int64_t)kNodesPerRange, p.n_dst - n0);
    const int64_t lane_off = ((int64_t)slab * SLAB + (int64_t)lane_eff * EPC) * (int64_t)sizeof(T);  // byte offset of chunk 0 within a row
    const int cp = __ldg(p.colptr + n0 + min(lane, nn));  // lanes 0..nn hold colptr[n0 + lane]
    int blk_base = __shfl_sync(0xffffffffu, cp, 0);
    const int E1 = __shfl_sync(0xffffffffu, cp, nn);
    int src_cur = (blk_base + lane < E1) ? __ldg(p.src + blk_base + lane) : 0;
    int src_nxt = (blk_base + 32 + lane < E1) ? __ldg(p.src + blk_base + 32 + lane) : 0;
    // head index of each of this lane's chunks and the lane's attribute offsets inside qw / abar rows
    int qw_off[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) qw_off[j] = ((slab * SLAB + (j * 32 + lane_eff) * EPC) / p.ch) * p.dp + sub;
    uint4 q_raw[NCH];
    float qw_raw[NCH][NA];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      q_raw[j] = ldg16(qp + n0 * ldq_b + lane_off + j * 512);
      if constexpr (MODE == 2) {
#pragma unroll
        for (int t = 0; t < NA; ++t)
          qw_raw[j][t] = (t < na && sub + t * LPH < p.dp) ? to_f32<T>(reinterpret_cast<const T*>(p.qw)[n0 * p.ldqw + qw_off[j] + t * LPH]) : 0.f;
      }
    }

    for (int nd = 0; nd < nn; ++nd) {
      const int64_t d = n0 + nd;
      const int e0 = __shfl_sync(0xffffffffu, cp, nd), e1 = __shfl_sync(0xffffffffu, cp, nd + 1);
      float q[NCH][EPC], acc[NCH][EPC], m_i[NCH], l_i[NCH], qw[NCH][NA], abar[NCH][NA];
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        CT::unpack(q_raw[j], q[j]);
        m_i[j] = -INFINITY, l_i[j] = 0.f;
#pragma unroll
        for (int i = 0; i < EPC; ++i) q[j][i] *= qscale, acc[j][i] = 0.f;
#pragma unroll
        for (int t = 0; t < NA; ++t) qw[j][t] = MODE == 2 ? qw_raw[j][t] * qscale : 0.f, abar[j][t] = 0.f;
      }
      if (nd + 1 < nn) {  // prefetch the next node's q / qw
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          q_raw[j] = ldg16(qp + (d + 1) * ldq_b + lane_off + j * 512);
          if constexpr (MODE == 2) {
#pragma unroll
            for (int t = 0; t < NA; ++t)
              if (t < na && sub + t * LPH < p.dp) qw_raw[j][t] = to_f32<T>(reinterpret_cast<const T*>(p.qw)[(d + 1) * p.ldqw + qw_off[j] + t * LPH]);
          }
        }
      }
      for (int eb = e0; eb < e1; eb += kBatch) {
        // ---- src ids of this batch from the prefetched blocks ----
        if (eb - blk_base >= 32) {
          blk_base += 32;
          src_cur = src_nxt;
          src_nxt = (blk_base + 32 + lane < E1) ? __ldg(p.src + blk_base + 32 + lane) : 0;
        }
        int sid[kBatch], eid[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          eid[b] = min(eb + b, e1 - 1);
          const int off = eid[b] - blk_base;
          const int a0 = __shfl_sync(0xffffffffu, src_cur, off & 31), a1 = __shfl_sync(0xffffffffu, src_nxt, off & 31);
          sid[b] = off < 32 ? a0 : a1;
        }
        // ---- issue all gathers of the batch ----
        uint4 k_raw[kBatch][NCH], v_raw[kBatch][NCH], e_raw[kBatch][NCH];
        float at[kBatch][NA];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const char* kr = kp + (int64_t)sid[b] * ldk_b + lane_off;
          const char* vr = vp + (int64_t)sid[b] * ldv_b + lane_off;
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            k_raw[b][j] = ldg16(kr + j * 512);
            v_raw[b][j] = ldg16(vr + j * 512);
            if constexpr (MODE == 1) e_raw[b][j] = ep ? ldg16(ep + (int64_t)eid[b] * lde_b + lane_off + j * 512) : make_uint4(0, 0, 0, 0);
          }
          if constexpr (MODE == 2) {
            const float* ar = p.edge_attr + (int64_t)eid[b] * p.lde + sub;  // rows are zero-padded to >= na*LPH floats (host contract)
#pragma unroll
            for (int t = 0; t < NA; ++t) at[b][t] = t < na ? __ldg(ar + t * LPH) : 0.f;
          }
        }
        // ---- scores ----
        float sc[kBatch][NCH];
#pragma unroll
        for (int b = 0; b < kBatch; ++b)
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            float kf[EPC];
            CT::unpack(k_raw[b][j], kf);
            float t = 0.f;
            if constexpr (MODE == 1) {
              float ef[EPC];
              CT::unpack(e_raw[b][j], ef);
#pragma unroll
              for (int i = 0; i < EPC; ++i) t += q[j][i] * (kf[i] + ef[i]);
            } else {
#pragma unroll
              for (int i = 0; i < EPC; ++i) t += q[j][i] * kf[i];
            }
            if constexpr (MODE == 2) {
#pragma unroll
              for (int tt = 0; tt < NA; ++tt)
                if (tt < na) t += qw[j][tt] * at[b][tt];
            }
#pragma unroll
            for (int o = 1; o < LPH; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            sc[b][j] = (eb + b < e1) ? t : -INFINITY;
          }
        // ---- online softmax update ----
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          float mx = m_i[j];
#pragma unroll
          for (int b = 0; b < kBatch; ++b) mx = fmaxf(mx, sc[b][j]);
          const float corr = exp2f(m_i[j] - mx);
          float w[kBatch], wsum = 0.f;
#pragma unroll
          for (int b = 0; b < kBatch; ++b) w[b] = exp2f(sc[b][j] - mx), wsum += w[b];
          l_i[j] = l_i[j] * corr + wsum;
          m_i[j] = mx;
#pragma unroll
          for (int i = 0; i < EPC; ++i) acc[j][i] *= corr;
#pragma unroll
          for (int b = 0; b < kBatch; ++b) {
            float vf[EPC];
            CT::unpack(v_raw[b][j], vf);
            if constexpr (MODE == 1) {
              float ef[EPC];
              CT::unpack(e_raw[b][j], ef);
#pragma unroll
              for (int i = 0; i < EPC; ++i) acc[j][i] += w[b] * (vf[i] + ef[i]);
            } else {
#pragma unroll
              for (int i = 0; i < EPC; ++i) acc[j][i] += w[b] * vf[i];
            }
          }
          if constexpr (MODE == 2) {
#pragma unroll
            for (int t = 0; t < NA; ++t)
              if (t < na) {
                float s2 = abar[j][t] * corr;
#pragma unroll
                for (int b = 0; b < kBatch; ++b) s2 += w[b] * at[b][t];
                abar[j][t] = s2;
              }
          }
        }
      }
      // ---- finalise node d ----
      const bool has_edges = e1 > e0;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float o[EPC];
        const float inv = has_edges ? 1.0f / l_i[j] : 0.f;
#pragma unroll
        for (int i = 0; i < EPC; ++i) o[i] = acc[j][i] * inv;
        if constexpr (MODE == 2) {
          if (has_edges && p.b_edge) {  // + b_e * sum(alpha) = + b_e
            const float* bp = p.b_edge + slab * SLAB + (j * 32 + lane_eff) * EPC;
#pragma unroll
            for (int i = 0; i < EPC; i += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(bp + i));
              o[i] += bv.x, o[i + 1] += bv.y, o[i + 2] += bv.z, o[i + 3] += bv.w;
            }
          }
          if (active) {
#pragma unroll
            for (int t = 0; t < NA; ++t)
              if (t < na && sub + t * LPH < p.dp)
                reinterpret_cast<T*>(p.abar)[d * p.ldabar + qw_off[j] + t * LPH] = from_f32<T>(abar[j][t] * inv);
          }
        }
        if (active) {
          if (p.add) {
            float r[EPC];
            CT::unpack(ldg16(reinterpret_cast<const char*>(p.add) + (d * p.ldadd) * (int64_t)sizeof(T) + lane_off + j * 512), r);
#pragma unroll
            for (int i = 0; i < EPC; ++i) o[i] += r[i];
          }
          *reinterpret_cast<uint4*>(reinterpret_cast<char*>(p.out) + (d * p.ldo) * (int64_t)sizeof(T) + lane_off + j * 512) = CT::pack(o);
        }
      }
    }
  }
}

// ---- pipelined kernel (cp.async ring) ----------------------------------------------------------------------------
// Same lane/chunk layout as the slab kernel, but the gathered rows no longer land in registers: every warp owns a ring of
// kSlots shared-memory slots and streams the k | v row segments (and the 64-byte attribute row, or the materialised e segment)
// of its edges into it with 16-byte cp.async (LDGSTS), kSlots - kBatch .. kSlots edges ahead of the arithmetic and ACROSS
// destination-node boundaries: the edges of a contiguous dst range are one contiguous run of the dst-sorted edge list.
// Bytes in flight are bounded by shared memory (kSlots x 2 KB per warp) instead of the register file, so the warp never waits
// on a gather it has just issued; one cp.async group per edge (empty groups past the end keep the count uniform) and
// cp.async.wait_group<kSlots - kBatch> before a batch is consumed.
// Work split: each warp takes a contiguous range of dst nodes holding ~E / (#warps) edges (boundaries by a warp-wide 32-ary
// search of colptr), so there is no per-range pipeline restart and the load is balanced by edges.
#ifndef ATTN_SLOTS
#define ATTN_SLOTS 8
#endif
#ifndef ATTN_PIPE_MINBLOCKS
#define ATTN_PIPE_MINBLOCKS 3
#endif
constexpr int kSlots = ATTN_SLOTS;

__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

#ifndef ATTN_PIPE_BATCH
#define ATTN_PIPE_BATCH 3
#endif
constexpr int kPBatch = ATTN_PIPE_BATCH;  // 3 divides every in-degree of the icosahedral multi-scale mesh (6, 12, ... 36): no padded slots

template <typename T, int NCH, int LPH, int MODE>
__global__ void __launch_bounds__(128, ATTN_PIPE_MINBLOCKS) gt_attention_pipe_kernel(const AttnParams p, int n_slabs, int active_lanes) {
  using CT = ChunkT<T>;
  constexpr int EPC = CT::EPC;
  constexpr int NA = (kMaxEdgeDim + LPH - 1) / LPH;
  constexpr int SLAB = NCH * 32 * EPC;
  constexpr int kKV = NCH * 512;                                          // bytes of one k (or v) row segment
  constexpr int kSlotBytes = 2 * kKV + (MODE == 2 ? 64 : kKV);             // k | v | attributes-or-e
  static_assert(kSlots > kPBatch, "ring must be deeper than a batch");
  extern __shared__ __align__(16) uint8_t smem_ring[];
  pdl_wait();  // PDL (common.cuh): q | k | v come from the GEMM just before
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const bool active = lane < active_lanes;
  const int lane_eff = active ? lane : 0;
  const int sub = lane & (LPH - 1);
  const int na = MODE == 2 ? (p.edge_dim + LPH - 1) / LPH : 0;
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem_ring) + (uint32_t)(wib * kSlots * kSlotBytes);
  const char* __restrict__ qp = reinterpret_cast<const char*>(p.q);
  const char* __restrict__ kp = reinterpret_cast<const char*>(p.k);
  const char* __restrict__ vp = reinterpret_cast<const char*>(p.v);
  const char* __restrict__ ep = reinterpret_cast<const char*>(p.e);
  const char* __restrict__ addp = reinterpret_cast<const char*>(p.add);
  const T* __restrict__ qwp = reinterpret_cast<const T*>(p.qw);
  const int64_t ldq_b = p.ldq * (int64_t)sizeof(T), ldk_b = p.ldk * (int64_t)sizeof(T), ldv_b = p.ldv * (int64_t)sizeof(T);
  const int64_t lde_b = p.lde_proj * (int64_t)sizeof(T), ldadd_b = p.ldadd * (int64_t)sizeof(T);
  const float qscale = p.scale * 1.4426950408889634f;

  // ---- this warp's (slab, dst range) ----
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int w = blockIdx.x * (blockDim.x >> 5) + wib;
  const int slab = w % n_slabs, R = warps_total / n_slabs;
  int r = w / n_slabs;
  if (r >= R) return;
  if (p.reverse) r = R - 1 - r;
  const int n_edges = __ldg(p.colptr + p.n_dst);
  const int lo_e = (int)((int64_t)n_edges * r / R), hi_e = (int)((int64_t)n_edges * (r + 1) / R);
  const int n_lo = r == 0 ? 0 : colptr_lower_bound(p.colptr, (int)p.n_dst, lo_e, lane);
  const int n_hi = r == R - 1 ? (int)p.n_dst : colptr_lower_bound(p.colptr, (int)p.n_dst, hi_e, lane);
  if (n_lo >= n_hi) return;
  const int64_t lane_off = ((int64_t)slab * SLAB + (int64_t)lane_eff * EPC) * (int64_t)sizeof(T);
  const int E0 = __ldg(p.colptr + n_lo), E1 = __ldg(p.colptr + n_hi);

  const char* const k_lane = kp + lane_off;
  const char* const v_lane = vp + lane_off;
  const char* const a_lane = reinterpret_cast<const char*>(p.edge_attr) + (lane & 3) * 16;
  const uint32_t ldk_b32 = (uint32_t)ldk_b, ldv_b32 = (uint32_t)ldv_b, lde_b32 = (uint32_t)p.lde * 4u;  // < 2^31 (checked by the launcher)
  // ---- producer state: src ids streamed in blocks of 32, one cp.async group per edge, running ring position ----
  int pe = E0;  // next edge to issue
  int ppos = 0;  // its ring slot
  int pblk = E0;
  int psrc = (pblk + lane < E1) ? __ldg(p.src + pblk + lane) : 0;
  int psrc_n = (pblk + 32 + lane < E1) ? __ldg(p.src + pblk + 32 + lane) : 0;
  auto issue = [&]() {
    if (pe < E1) {
      if (pe - pblk >= 32) {
        pblk += 32;
        psrc = psrc_n;
        psrc_n = (pblk + 32 + lane < E1) ? __ldg(p.src + pblk + 32 + lane) : 0;
      }
      const int sid = __shfl_sync(0xffffffffu, psrc, pe - pblk);
      const uint32_t slot = ring + (uint32_t)(ppos * kSlotBytes) + lane * 16;
      // 32-bit row pitch x 32-bit id on a per-lane base pointer: one IMAD.WIDE.U32 per gather address (the 64-bit form cost ~10)
      const char* kr = k_lane + (uint64_t)(uint32_t)sid * ldk_b32;
      const char* vr = v_lane + (uint64_t)(uint32_t)sid * ldv_b32;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        cp_async16(slot + j * 512, kr + j * 512);
        cp_async16(slot + kKV + j * 512, vr + j * 512);
      }
      if constexpr (MODE == 2) {
        if (lane < 4) cp_async16(slot + 2 * kKV, a_lane + (uint64_t)(uint32_t)pe * lde_b32);
      } else {
        if (ep) {
#pragma unroll
          for (int j = 0; j < NCH; ++j) cp_async16(slot + 2 * kKV + j * 512, ep + (int64_t)pe * lde_b + lane_off + j * 512);
        }
      }
      ++pe;
      if (++ppos == kSlots) ppos = 0;
    }
    cp_async_commit();  // empty past the end: keeps "groups issued = edges consumed + kSlots" uniform
  };
#pragma unroll 1
  for (int i = 0; i < kSlots; ++i) issue();

  int qw_off[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) qw_off[j] = ((slab * SLAB + (j * 32 + lane_eff) * EPC) / p.ch) * p.dp + sub;

  // node-level operands (q, qw, add) of the NEXT node are fetched one node ahead into registers (L1 path; the gathers use cp.async.cg)
  uint4 q_nx[NCH], add_nx[NCH];
  T qw_nx[NCH][NA];
  auto fetch_node = [&](int d) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      q_nx[j] = ldg16(qp + (int64_t)d * ldq_b + lane_off + j * 512);
      if (addp) add_nx[j] = ldg16(addp + (int64_t)d * ldadd_b + lane_off + j * 512);
      if constexpr (MODE == 2) {
#pragma unroll
        for (int t = 0; t < NA; ++t)
          if (t < na && sub + t * LPH < p.dp) qw_nx[j][t] = qwp[(int64_t)d * p.ldqw + qw_off[j] + t * LPH];
      }
    }
  };
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    add_nx[j] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int t = 0; t < NA; ++t) qw_nx[j][t] = from_f32<T>(0.f);
  }
  fetch_node(n_lo);

  int cpos = 0;  // ring slot of the next edge to consume
  int cbase = n_lo;  // colptr block: lanes hold colptr[cbase + lane]
  int cp = __ldg(p.colptr + min(cbase + lane, n_hi));
  for (int d = n_lo; d < n_hi; ++d) {
    if (d - cbase >= 31) {
      cbase = d;
      cp = __ldg(p.colptr + min(cbase + lane, n_hi));
    }
    const int e0 = __shfl_sync(0xffffffffu, cp, d - cbase), e1 = __shfl_sync(0xffffffffu, cp, d - cbase + 1);
    float q[NCH][EPC], acc[NCH][EPC], m_i[NCH], l_i[NCH], qw[NCH][NA], abar[NCH][NA];
    uint4 add_cur[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      CT::unpack(q_nx[j], q[j]);
      add_cur[j] = add_nx[j];
      m_i[j] = -INFINITY, l_i[j] = 0.f;
#pragma unroll
      for (int i = 0; i < EPC; ++i) q[j][i] *= qscale, acc[j][i] = 0.f;
#pragma unroll
      for (int t = 0; t < NA; ++t) abar[j][t] = 0.f, qw[j][t] = MODE == 2 ? to_f32<T>(qw_nx[j][t]) * qscale : 0.f;
    }
    if (d + 1 < n_hi) fetch_node(d + 1);
    for (int eb = e0; eb < e1; eb += kPBatch) {
      const int nb = min(kPBatch, e1 - eb);
      cp_async_wait<kSlots - kPBatch>();  // edges eb .. eb+kPBatch-1 have landed (this lane's copies)
      __syncwarp();                       // ... and every other lane's
      uint32_t sl[kPBatch];               // slot addresses of the batch (entries past the node's last edge alias slot 0: valid data, weight 0)
#pragma unroll
      for (int b = 0; b < kPBatch; ++b) {
        int pos = cpos + (b < nb ? b : 0);
        if (pos >= kSlots) pos -= kSlots;
        sl[b] = ring + (uint32_t)(pos * kSlotBytes) + lane * 16;
      }
      float sc[kPBatch][NCH], at[kPBatch][NA];
#pragma unroll
      for (int b = 0; b < kPBatch; ++b) {
        if constexpr (MODE == 2) {
#pragma unroll
          for (int t = 0; t < NA; ++t) at[b][t] = t < na ? lds32f(sl[b] - lane * 16 + 2 * kKV + (sub + t * LPH) * 4) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          float kf[EPC];
          CT::unpack(lds16(sl[b] + j * 512), kf);
          if constexpr (MODE == 1) {
            if (ep) {
              float ef[EPC];
              CT::unpack(lds16(sl[b] + 2 * kKV + j * 512), ef);
#pragma unroll
              for (int i = 0; i < EPC; ++i) kf[i] += ef[i];
            }
          }
          // packed fp32x2 FMAs (FFMA2): the kernel is issue-bound, two lanes of work per issue slot
          float2 t2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < EPC; i += 2) t2 = __ffma2_rn(make_float2(q[j][i], q[j][i + 1]), make_float2(kf[i], kf[i + 1]), t2);
          float t = t2.x + t2.y;
          if constexpr (MODE == 2) {
#pragma unroll
            for (int tt = 0; tt < NA; ++tt)
              if (tt < na) t += qw[j][tt] * at[b][tt];
          }
#pragma unroll
          for (int o = 1; o < LPH; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          sc[b][j] = b < nb ? t : -INFINITY;
        }
      }
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float mx = m_i[j];
#pragma unroll
        for (int b = 0; b < kPBatch; ++b) mx = fmaxf(mx, sc[b][j]);
        const float corr = exp2f(m_i[j] - mx);
        float wgt[kPBatch], wsum = 0.f;
#pragma unroll
        for (int b = 0; b < kPBatch; ++b) wgt[b] = exp2f(sc[b][j] - mx), wsum += wgt[b];
        l_i[j] = l_i[j] * corr + wsum;
        m_i[j] = mx;
        const float2 corr2 = make_float2(corr, corr);
#pragma unroll
        for (int i = 0; i < EPC; i += 2) {
          const float2 a2 = __fmul2_rn(make_float2(acc[j][i], acc[j][i + 1]), corr2);
          acc[j][i] = a2.x, acc[j][i + 1] = a2.y;
        }
#pragma unroll
        for (int b = 0; b < kPBatch; ++b) {
          float vf[EPC];
          CT::unpack(lds16(sl[b] + kKV + j * 512), vf);
          if constexpr (MODE == 1) {
            if (ep) {
              float ef[EPC];
              CT::unpack(lds16(sl[b] + 2 * kKV + j * 512), ef);
#pragma unroll
              for (int i = 0; i < EPC; ++i) vf[i] += ef[i];
            }
          }
          const float2 w2 = make_float2(wgt[b], wgt[b]);
#pragma unroll
          for (int i = 0; i < EPC; i += 2) {
            const float2 a2 = __ffma2_rn(w2, make_float2(vf[i], vf[i + 1]), make_float2(acc[j][i], acc[j][i + 1]));
            acc[j][i] = a2.x, acc[j][i + 1] = a2.y;
          }
        }
        if constexpr (MODE == 2) {
#pragma unroll
          for (int t = 0; t < NA; ++t)
            if (t < na) {
              float s2 = abar[j][t] * corr;
#pragma unroll
              for (int b = 0; b < kPBatch; ++b) s2 += wgt[b] * at[b][t];
              abar[j][t] = s2;
            }
        }
      }
      cpos += nb;
      if (cpos >= kSlots) cpos -= kSlots;
      __syncwarp();  // all lanes are done reading these slots before they are refilled
      for (int i = 0; i < nb; ++i) issue();
    }
    // ---- finalise node d ----
    const bool has_edges = e1 > e0;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float o[EPC];
      const float inv = has_edges ? 1.0f / l_i[j] : 0.f;
      if (p.lse && active && sub == 0)  // scores live in the log2 domain here: back to natural logarithms
        p.lse[(int64_t)d * p.heads + (slab * SLAB + (j * 32 + lane_eff) * EPC) / p.ch] = has_edges ? (m_i[j] + log2f(l_i[j])) * 0.6931471805599453f : 0.f;
#pragma unroll
      for (int i = 0; i < EPC; ++i) o[i] = acc[j][i] * inv;
      if constexpr (MODE == 2) {
        if (has_edges && p.b_edge) {
          const float* bp = p.b_edge + slab * SLAB + (j * 32 + lane_eff) * EPC;
#pragma unroll
          for (int i = 0; i < EPC; i += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bp + i));
            o[i] += bv.x, o[i + 1] += bv.y, o[i + 2] += bv.z, o[i + 3] += bv.w;
          }
        }
        if (active) {
#pragma unroll
          for (int t = 0; t < NA; ++t)
            if (t < na && sub + t * LPH < p.dp)
              reinterpret_cast<T*>(p.abar)[(int64_t)d * p.ldabar + qw_off[j] + t * LPH] = from_f32<T>(abar[j][t] * inv);
        }
      }
      if (active) {
        if (addp) {
          float rr[EPC];
          CT::unpack(add_cur[j], rr);
#pragma unroll
          for (int i = 0; i < EPC; ++i) o[i] += rr[i];
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<char*>(p.out) + ((int64_t)d * p.ldo) * (int64_t)sizeof(T) + lane_off + j * 512) = CT::pack(o);
      }
    }
  }
  cp_async_wait<0>();
}

// Generic shapes (any H, Ch <= 256, any alignment): one warp per (dst, head), lanes stride over the head's channels.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) gt_attention_generic_kernel(const AttnParams p) {
  constexpr int MAXV = 8;
  const int lane = threadIdx.x & 31;
  const int D = p.edge_dim;
  const T* __restrict__ qp = reinterpret_cast<const T*>(p.q);
  const T* __restrict__ kp = reinterpret_cast<const T*>(p.k);
  const T* __restrict__ vp = reinterpret_cast<const T*>(p.v);
  const T* __restrict__ ep = reinterpret_cast<const T*>(p.e);
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t items = p.n_dst * p.heads;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < items; w += warps_total) {
    const int64_t d = w / p.heads;
    const int h = (int)(w - d * p.heads);
    const int e0 = p.colptr[d], e1 = p.colptr[d + 1];
    const int base = h * p.ch;
    float q[MAXV], acc[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      q[i] = c < p.ch ? to_f32<T>(qp[d * p.ldq + base + c]) * p.scale : 0.f;
      acc[i] = 0.f;
    }
    float m_i = -INFINITY, l_i = 0.f;
    for (int eidx = e0; eidx < e1; ++eidx) {
      const int s = p.src[eidx];
      float vv[MAXV];
      float sc = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        float kk = 0.f;
        vv[i] = 0.f;
        if (c < p.ch) {
          kk = to_f32<T>(kp[(int64_t)s * p.ldk + base + c]);
          vv[i] = to_f32<T>(vp[(int64_t)s * p.ldv + base + c]);
          float ee = 0.f;
          if constexpr (MODE == 1) ee = to_f32<T>(ep[(int64_t)eidx * p.lde_proj + base + c]);
          if constexpr (MODE == 2) {
            ee = p.b_edge ? p.b_edge[base + c] : 0.f;
            for (int j = 0; j < D; ++j) ee += p.w_edge[(int64_t)(base + c) * p.ldw_e + j] * p.edge_attr[(int64_t)eidx * p.lde + j];
          }
          kk += ee, vv[i] += ee;
        }
        sc += q[i] * kk;
      }
      sc = warp_sum(sc);
      const float m_new = fmaxf(m_i, sc);
      const float corr = __expf(m_i - m_new), wgt = __expf(sc - m_new);
      l_i = l_i * corr + wgt;
      m_i = m_new;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) acc[i] = acc[i] * corr + wgt * vv[i];
    }
    const float inv = e1 > e0 ? 1.0f / l_i : 0.f;
    if (p.lse && lane == 0) p.lse[d * p.heads + h] = e1 > e0 ? m_i + logf(l_i) : 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < p.ch) {
        float o = acc[i] * inv;
        if (p.add) o += to_f32<T>(reinterpret_cast<const T*>(p.add)[d * p.ldadd + base + c]);
        reinterpret_cast<T*>(p.out)[d * p.ldo + base + c] = from_f32<T>(o);
      }
    }
  }
}

template <typename T, int NCH, int LPH, int MODE>
static int launch_pipe_mode(const AttnParams& p, int n_slabs, int active_lanes, cudaStream_t s) {
  constexpr int kSlotBytes = 2 * NCH * 512 + (MODE == 2 ? 64 : NCH * 512);
  constexpr int smem = 4 * kSlots * kSlotBytes;  // 4 warps per CTA
  const int64_t max_pitch = (int64_t)1 << 31;  // gather addresses are formed from 32-bit byte pitches
  if (p.ldk * (int64_t)sizeof(T) >= max_pitch || p.ldv * (int64_t)sizeof(T) >= max_pitch || p.lde * 4 >= max_pitch) return 1;
  static int blocks_per_sm_dev[kMaxDevices] = {};
  int& blocks_per_sm = blocks_per_sm_dev[current_device()];
  if (blocks_per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(gt_attention_pipe_kernel<T, NCH, LPH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attention)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, gt_attention_pipe_kernel<T, NCH, LPH, MODE>, 128, smem);
    if (e != cudaSuccess || blocks_per_sm < 1) blocks_per_sm = 1;
  }
  // one persistent wave; every warp gets ~E/#warps edges; #warps must be a multiple of n_slabs (4 | warps, n_slabs in {1,2,4,...})
  int64_t blocks = (int64_t)num_sms() * blocks_per_sm;
  const int64_t max_useful = (p.n_dst * n_slabs + 3) / 4;
  if (blocks > max_useful) blocks = max_useful;
  while ((blocks * 4) % n_slabs) ++blocks;
  cudaError_t le = launch_pdl(gt_attention_pipe_kernel<T, NCH, LPH, MODE>, dim3((unsigned)blocks), dim3(128), smem, s, p, n_slabs, active_lanes);
  if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(gt_attention_pipe_kernel)");
  return launch_status("gt_attention_pipe_kernel");
}

template <typename T, int NCH, int LPH>
static int launch_slab(const AttnParams& p, int mode, int n_slabs, int active_lanes, cudaStream_t s) {
#ifdef ATTN_USE_SLAB_KERNEL
  const int64_t items = ((p.n_dst + kNodesPerRange - 1) / kNodesPerRange) * n_slabs;
  const int warps_per_block = 4;
  int64_t blocks = (items + warps_per_block - 1) / warps_per_block;
  const int64_t cap = (int64_t)num_sms() * ATTN_MINBLOCKS;  // resident CTAs per SM (register-limited), persistent over the items
  if (blocks > cap) blocks = cap;
  if (mode == 2)
    gt_attention_slab_kernel<T, NCH, LPH, 2><<<(unsigned)blocks, 128, 0, s>>>(p, n_slabs, active_lanes);
  else
    gt_attention_slab_kernel<T, NCH, LPH, 1><<<(unsigned)blocks, 128, 0, s>>>(p, n_slabs, active_lanes);
  return launch_status("gt_attention_slab_kernel");
#else
  return mode == 2 ? launch_pipe_mode<T, NCH, LPH, 2>(p, n_slabs, active_lanes, s)
                   : launch_pipe_mode<T, NCH, LPH, 1>(p, n_slabs, active_lanes, s);
#endif
}

template <typename T, int NCH>
static int launch_slab_lph(const AttnParams& p, int mode, int lph, int n_slabs, int active_lanes, cudaStream_t s) {
  switch (lph) {
    case 2: return launch_slab<T, NCH, 2>(p, mode, n_slabs, active_lanes, s);
    case 4: return launch_slab<T, NCH, 4>(p, mode, n_slabs, active_lanes, s);
    case 8: return launch_slab<T, NCH, 8>(p, mode, n_slabs, active_lanes, s);
    case 16: return launch_slab<T, NCH, 16>(p, mode, n_slabs, active_lanes, s);
    default: return 1;  // not a fast-path shape
  }
}

template <typename T>
static int dispatch(const AttnParams& p, int mode, bool aligned, bool folded, cudaStream_t s) {
  const int C = p.heads * p.ch;
  constexpr int EPC = ChunkT<T>::EPC;
  if (aligned && C % EPC == 0 && p.ch % EPC == 0) {
    const int lph = p.ch / EPC;
    const int chunks = C / EPC;  // 16-byte chunks per row
    int rc = 1;
    if (chunks < 32) {
      rc = launch_slab_lph<T, 1>(p, mode, lph, 1, chunks, s);  // narrow rows: part of the warp idles
    } else if (chunks % 32 == 0 && (32 % lph) == 0) {
      const int full = chunks / 32;
      rc = (full % 2 == 0) ? launch_slab_lph<T, 2>(p, mode, lph, full / 2, 32, s) : launch_slab_lph<T, 1>(p, mode, lph, full, 32, s);
    }
    if (rc <= 0) return rc;
  }
  if (folded) {
    set_error("gt_attention: the folded lin_edge form needs a slab-kernel shape (16-byte aligned rows, Ch/(16/elemsize) in {2,4,8,16}, <= 16 "
              "attributes padded to 16 floats); use the w_edge form for H=%d Ch=%d",
              p.heads, p.ch);
    return -3;
  }
  if (p.ch > 256) {
    set_error("gt_attention: channels per head %d > 256 unsupported", p.ch);
    return -3;
  }
  const int64_t items = p.n_dst * p.heads;
  int64_t blocks = (items + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (mode == 0)
    gt_attention_generic_kernel<T, 0><<<(unsigned)blocks, 256, 0, s>>>(p);
  else if (mode == 1)
    gt_attention_generic_kernel<T, 1><<<(unsigned)blocks, 256, 0, s>>>(p);
  else
    gt_attention_generic_kernel<T, 2><<<(unsigned)blocks, 256, 0, s>>>(p);
  return launch_status("gt_attention_generic_kernel");
}

}  // namespace anemoi

using namespace anemoi;

extern "C" int anemoi_b200_gt_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* e,
                                            int64_t lde_proj, const float* edge_attr, int64_t lde, int64_t edge_dim, const float* w_edge,
                                            int64_t ldw_e, const float* b_edge, const void* qw, int64_t ldqw, void* abar, int64_t ldabar,
                                            int64_t dp, const int32_t* src32, const int32_t* colptr32, const void* add, int64_t ldadd, void* out,
                                            int64_t ldo, float* lse, int64_t n_dst, int64_t heads, int64_t ch, int dtype, int flags, void* stream) {
  ANEMOI_CHECK_ARG(n_dst >= 0 && heads >= 1 && ch >= 1, "gt_attention: bad shape");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "gt_attention: bad dtype %d", dtype);
  if (n_dst == 0) return 0;
  ANEMOI_CHECK_ARG(q && k && v && out && src32 && colptr32, "gt_attention: null pointer");
  ANEMOI_CHECK_ARG(!(e && edge_attr), "gt_attention: pass either a materialised edge projection or raw edge attributes, not both");
  const int64_t C = heads * ch;
  ANEMOI_CHECK_ARG(ldq >= C && ldk >= C && ldv >= C && ldo >= C, "gt_attention: leading dimension < heads*ch");
  const int mode = e ? 1 : (edge_attr ? 2 : 0);
  const bool folded = mode == 2 && qw != nullptr;
  if (mode == 2) {
    ANEMOI_CHECK_ARG(edge_dim >= 1 && lde >= edge_dim, "gt_attention: bad edge attribute shape");
    ANEMOI_CHECK_ARG(folded || (w_edge && ldw_e >= edge_dim), "gt_attention: fused lin_edge needs either qw/abar (folded form) or w_edge");
    ANEMOI_CHECK_ARG(!folded || (abar && dp >= edge_dim && ldqw >= heads * dp && ldabar >= heads * dp), "gt_attention: bad folded lin_edge arguments");
  }
  AttnParams p;
  p.q = q, p.k = k, p.v = v, p.e = e, p.add = add, p.out = out;
  p.ldq = ldq, p.ldk = ldk, p.ldv = ldv, p.lde_proj = lde_proj, p.ldadd = ldadd, p.ldo = ldo;
  p.edge_attr = edge_attr, p.lde = lde, p.edge_dim = (int)edge_dim, p.w_edge = w_edge, p.ldw_e = ldw_e, p.b_edge = b_edge;
  p.qw = qw, p.abar = abar, p.ldqw = ldqw, p.ldabar = ldabar, p.dp = (int)dp;
  p.src = src32, p.colptr = colptr32, p.n_dst = n_dst, p.heads = (int)heads, p.ch = (int)ch;
  p.scale = 1.0f / sqrtf((float)ch);
  p.lse = lse;
  p.reverse = (flags & ANEMOI_EPI_REVERSE) ? 1 : 0;
  const int es = dtype == ANEMOI_BF16 ? 2 : 4;
  auto al = [&](const void* ptr, int64_t ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * es) % 16 == 0); };
  bool slab_ok = al(q, ldq) && al(k, ldk) && al(v, ldv) && al(e, lde_proj) && al(add, ldadd) && al(out, ldo);
  if (mode == 2) {
    // slab (fast) path needs the folded form, <= 16 attributes, rows zero-padded to 16 floats, 16-byte aligned bias
    slab_ok = slab_ok && folded && edge_dim <= kMaxEdgeDim && lde >= kMaxEdgeDim && (!b_edge || (reinterpret_cast<uintptr_t>(b_edge) & 15) == 0);
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ANEMOI_BF16 && folded && slab_ok && (ch == 32 || ch == 64)) {
    // mma.sync formulation (attention_mma.cu), opt-in with ANEMOI_B200_ATTN_MMA=1: numerically equivalent and 3x fewer arithmetic
    // instructions, but measured SLOWER than the FP32-pipe kernel (235 vs 135 us at cfg2; profiles/README.md): per-edge issue and
    // per-node bookkeeping dominate both kernels, and at 200+ registers only 8 warps per SM hide the ldmatrix -> HMMA chains.
    static int use_mma = -1;
    if (use_mma < 0) {
      const char* ev = getenv("ANEMOI_B200_ATTN_MMA");
      use_mma = (ev && ev[0] == '1') ? 1 : 0;
    }
    if (use_mma && !lse) {
      const int rc_mma = launch_gt_attention_mma(p, s);
      if (rc_mma <= 0) return rc_mma;
    }
  }
  const int rc = dtype == ANEMOI_BF16 ? dispatch<__nv_bfloat16>(p, mode, slab_ok, folded, s) : dispatch<float>(p, mode, slab_ok, folded, s);
  return rc;
}
