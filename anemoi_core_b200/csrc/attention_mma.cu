// attention_mma.cu — GraphTransformer edge-softmax attention, folded lin_edge form (MODE 2), bf16, on the warp-level tensor cores.
//
// Why tensor cores here at all: per destination node the products are matrix-VECTOR shaped (one q row against ~8 gathered k rows), so a
// tcgen05 tile (M >= 64 rows of ONE operand pair) has nothing to chew on; the FP32-pipe kernel (attention.cu) is issue-bound instead -
// 252 warp-instructions per edge, a third of them bf16 -> fp32 unpacks (profiles/r1_ncu_attention_v8_sass_summary.txt).  mma.sync with
// the 8 gathered edges of a node as the N dimension does the same arithmetic straight from the bf16 rows in shared memory (ldmatrix, no
// unpack) at ~1/8 tensor utilisation, which is irrelevant (the kernel is bound by the L2 gather) but cuts the instruction count ~5x.
//
// A warp owns a channel slab of 256 channels = HPW = 256 / Ch heads and a contiguous dst range holding ~E / #warps edges (same split and
// the same cp.async ring as the pipe kernel: one group per edge, k | v | 16 fp32 attributes per slot, slots 1104 bytes apart so the eight
// 16-byte rows of an ldmatrix hit eight different bank groups).  Per tile of <= 8 edges of one destination node:
//   S^T[head, edge]  = sum_c Q'[head, c] K[edge, c]     m16n8k16 bf16: A = q placed in its head's row (zero elsewhere), B = ldmatrix(K)
//                    + sum_a QW[head, a] attr[edge, a]   m16n8k8 tf32 x 2: A = W_e^T q per head, B = the fp32 attributes as they are
//   online softmax per head in the log2 domain (a head's 8 scores live in one 4-lane group: 2 xor-shuffles per reduction)
//   P[edge, head] as bf16 is, register for register, the B fragment of the next products (the accumulator layout of S^T matches it):
//   D[c, head]      += sum_e V[e, c] P[e, head]          m16n8k8 bf16: A = ldmatrix.trans(V); only the column of c's own head is used
//   abar^T[a, head] += sum_e attr[e, a] P[e, head]       m16n8k8 bf16
// Finalise: out = D / l + b_edge + self term, staged through shared memory so that the row leaves as one coalesced 512-byte store.
#include "attention.h"

namespace anemoi {
namespace {

constexpr int kCW = 256;                          // channels per warp
constexpr int kSlotBytes = 512 + 512 + 64 + 16;   // k | v | attributes | pad: 276 words = 20 (mod 32) -> conflict-free ldmatrix rows
constexpr int kMSlots = 24;                       // ring depth in edges: a tile of 8 is consumed while 16 more are in flight
constexpr int kTile = 8;
constexpr int kWarpSmem = kMSlots * kSlotBytes + kCW * 4;  // ring + fp32 staging row

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16_k8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mma_tf32_k8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

template <int CH>
__global__ void __launch_bounds__(128, 2) gt_attention_mma_kernel(const AttnParams p, int n_slabs) {
  constexpr int HPW = kCW / CH;   // heads per warp (8 for Ch = 32, 4 for Ch = 64)
  constexpr int KSPH = CH / 16;   // k-steps (16 channels) per head
  static_assert(HPW <= 8 && HPW >= 1 && CH % 32 == 0, "one head per accumulator row, 32-channel ldmatrix groups inside one head");
  extern __shared__ __align__(16) uint8_t smem_mma[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem_mma) + (uint32_t)(wib * kWarpSmem);
  float* stage_g = reinterpret_cast<float*>(smem_mma + wib * kWarpSmem + kMSlots * kSlotBytes);
  // never-written slots are read (and weighted by exactly 0) when a tile is padded past the end of the range: keep them finite
  for (int i = lane; i < kMSlots * kSlotBytes / 16; i += 32) reinterpret_cast<uint4*>(smem_mma + wib * kWarpSmem)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();

  const char* __restrict__ qp = reinterpret_cast<const char*>(p.q);
  const char* __restrict__ kp = reinterpret_cast<const char*>(p.k);
  const char* __restrict__ vp = reinterpret_cast<const char*>(p.v);
  const char* __restrict__ addp = reinterpret_cast<const char*>(p.add);
  const __nv_bfloat16* __restrict__ qwp = reinterpret_cast<const __nv_bfloat16*>(p.qw);
  // row pitches in bytes as 32-bit values (checked on the host): every gather address is ONE IMAD.WIDE.U32 on a per-lane base pointer
  const uint32_t ldq_b = (uint32_t)p.ldq * 2u, ldk_b = (uint32_t)p.ldk * 2u, ldv_b = (uint32_t)p.ldv * 2u, ldadd_b = (uint32_t)p.ldadd * 2u;
  const uint32_t lde_b = (uint32_t)p.lde * 4u, ldqw_b = (uint32_t)p.ldqw * 2u, ldab_b = (uint32_t)p.ldabar * 2u, ldo_b = (uint32_t)p.ldo * 2u;
  const float qscale = p.scale * 1.4426950408889634f;

  // ---- this warp's (slab, dst range): contiguous nodes holding ~E / R edges ----
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int w = blockIdx.x * (blockDim.x >> 5) + wib;
  const int slab = w % n_slabs, r = w / n_slabs, R = warps_total / n_slabs;
  if (r >= R) return;
  const int n_edges = __ldg(p.colptr + p.n_dst);
  const int lo_e = (int)((int64_t)n_edges * r / R), hi_e = (int)((int64_t)n_edges * (r + 1) / R);
  const int n_lo = r == 0 ? 0 : colptr_lower_bound(p.colptr, (int)p.n_dst, lo_e, lane);
  const int n_hi = r == R - 1 ? (int)p.n_dst : colptr_lower_bound(p.colptr, (int)p.n_dst, hi_e, lane);
  if (n_lo >= n_hi) return;
  const uint32_t slab_off = (uint32_t)slab * kCW * 2u;  // byte offset of the slab inside a node row
  const char* const k_lane = kp + slab_off + lane * 16;
  const char* const v_lane = vp + slab_off + lane * 16;
  const char* const a_lane = reinterpret_cast<const char*>(p.edge_attr) + (lane & 3) * 16;
  const int E0 = __ldg(p.colptr + n_lo), E1 = __ldg(p.colptr + n_hi);

  // ---- producer: one cp.async group per edge (k | v | attributes), src ids streamed 32 at a time ----
  int pe = E0, ppos = 0, pblk = E0;
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
      cp_async16(slot, k_lane + (uint64_t)(uint32_t)sid * ldk_b);
      cp_async16(slot + 512, v_lane + (uint64_t)(uint32_t)sid * ldv_b);
      if (lane < 4) cp_async16(slot + 1024, a_lane + (uint64_t)(uint32_t)pe * lde_b);
      ++pe;
      if (++ppos == kMSlots) ppos = 0;
    }
    cp_async_commit();  // empty past the end: "groups issued = edges consumed + kMSlots" stays uniform
  };
#pragma unroll 1
  for (int i = 0; i < kMSlots; ++i) issue();

  // ---- node-level operands, fetched one node ahead: the q fragment of this lane's head (row g) and its W_e^T q ----
  const bool head_ok = g < HPW;
  const int head = slab * HPW + g;  // global head index of accumulator row g
  const char* const q_lane = qp + slab_off + (g * CH + 2 * t) * 2;
  const char* const qw_lane = reinterpret_cast<const char*>(qwp) + (head * p.dp + t) * 2;
  bool qw_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) qw_ok[i] = head_ok && (t + 4 * i) < p.dp;
  uint32_t q_nx[KSPH][2];
  unsigned short qw_nx[4];  // raw bf16 bits; widened at use so that the loads stay in flight across the previous node's arithmetic
  auto fetch_node = [&](int d) {
    const char* qb = q_lane + (uint64_t)(uint32_t)d * ldq_b;
    const char* wb = qw_lane + (uint64_t)(uint32_t)d * ldqw_b;
#pragma unroll
    for (int ks = 0; ks < KSPH; ++ks) {
      q_nx[ks][0] = head_ok ? __ldg(reinterpret_cast<const uint32_t*>(qb + ks * 32)) : 0u;
      q_nx[ks][1] = head_ok ? __ldg(reinterpret_cast<const uint32_t*>(qb + ks * 32 + 16)) : 0u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) qw_nx[i] = qw_ok[i] ? __ldg(reinterpret_cast<const unsigned short*>(wb + i * 8)) : (unsigned short)0;
  };
  fetch_node(n_lo);
  // abar destinations of this lane: rows = attributes g, g + 8; columns = heads 2t, 2t + 1 of the slab
  const uint32_t ab_off0 = (uint32_t)(((slab * HPW + 2 * t) * p.dp + g) * 2), ab_off1 = ab_off0 + (uint32_t)(p.dp * 2);
  const bool ab_h0 = 2 * t < HPW, ab_h1 = 2 * t + 1 < HPW, ab_lo = g < p.dp, ab_hi = g + 8 < p.dp;
  const float4 be0 = p.b_edge ? __ldg(reinterpret_cast<const float4*>(p.b_edge + slab * kCW + lane * 8)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 be1 = p.b_edge ? __ldg(reinterpret_cast<const float4*>(p.b_edge + slab * kCW + lane * 8) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);

  int cpos = 0;
  int e_next = E0;
  for (int d = n_lo; d < n_hi; ++d) {
    const int e0 = e_next, e1 = __ldg(p.colptr + d + 1);
    e_next = e1;
    uint32_t qa[KSPH][2], qwa[4];
#pragma unroll
    for (int ks = 0; ks < KSPH; ++ks) qa[ks][0] = q_nx[ks][0], qa[ks][1] = q_nx[ks][1];
#pragma unroll
    for (int i = 0; i < 4; ++i) qwa[i] = (uint32_t)qw_nx[i] << 16;  // attributes t, t+4, t+8, t+12 of head g as fp32 bit patterns (bf16 -> tf32: exact)
    if (d + 1 < n_hi) fetch_node(d + 1);

    float D[kCW / 16][4], Dab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < kCW / 16; ++i) D[i][0] = D[i][1] = D[i][2] = D[i][3] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int eb = e0; eb < e1; eb += kTile) {
      const int nb = min(kTile, e1 - eb);
      cp_async_wait<kMSlots - kTile>();  // the oldest 8 groups (a superset of this tile's edges) have landed
      __syncwarp();
      // ldmatrix row address of this lane: matrix lane/8 (16-byte column group), row lane%8 (edge slot)
      int pos = cpos + (lane & 7);
      if (pos >= kMSlots) pos -= kMSlots;
      const uint32_t row_k = ring + (uint32_t)(pos * kSlotBytes) + (uint32_t)((lane >> 3) * 16);
      // ---- scores.  One accumulator per head h: Sh[h] = Q_all . K_h^T, where row g of Q_all is head g's q at the channel offsets of
      // the k-step (every row multiplies head h's keys; only row h of Sh[h] is a score, the others are discarded).  No operand
      // masking, and HPW independent MMA chains instead of one 16 deep.
      float Sh[HPW][4], Sa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int h = 0; h < HPW; ++h) Sh[h][0] = Sh[h][1] = Sh[h][2] = Sh[h][3] = 0.f;
#pragma unroll
      for (int jj = 0; jj < kCW / 32; ++jj) {
        uint32_t b[4];
        ldmatrix_x4(b, row_k + jj * 64);
        const int hj = (jj * 32) / CH;    // head of these two k-steps
        const int ks0 = (2 * jj) % KSPH;  // their index inside the head
        mma_bf16_k16(Sh[hj], qa[ks0][0], 0u, qa[ks0][1], 0u, b[0], b[1]);
        mma_bf16_k16(Sh[hj], qa[ks0 + 1][0], 0u, qa[ks0 + 1][1], 0u, b[2], b[3]);
      }
      {  // edge-attribute term (every row is its own head's W_e^T q): B[k = attribute, n = edge g] from the fp32 attribute row of slot g
        int pg = cpos + g;
        if (pg >= kMSlots) pg -= kMSlots;
        const uint32_t arow = ring + (uint32_t)(pg * kSlotBytes) + 1024u + (uint32_t)(t * 4);
        mma_tf32_k8(Sa, qwa[0], 0u, qwa[1], 0u, lds32(arow), lds32(arow + 16));
        mma_tf32_k8(Sa, qwa[2], 0u, qwa[3], 0u, lds32(arow + 32), lds32(arow + 48));
      }
      float sx = Sh[0][0], sy = Sh[0][1];  // row g of Sh[g]
#pragma unroll
      for (int h = 1; h < HPW; ++h) sx = g == h ? Sh[h][0] : sx, sy = g == h ? Sh[h][1] : sy;
      // ---- online softmax of head g over edge columns 2t, 2t+1 (log2 domain) ----
      const float s0 = (2 * t < nb) ? (sx + Sa[0]) * qscale : -INFINITY;
      const float s1 = (2 * t + 1 < nb) ? (sy + Sa[1]) * qscale : -INFINITY;
      float mx = fmaxf(s0, s1);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);  // finite: nb >= 1
      const float corr = exp2f(m_run - m_new);
      const uint32_t pb = pack_bf16x2(exp2f(s0 - m_new), exp2f(s1 - m_new));  // B fragment of the P products: (edges 2t, 2t+1; head g)
      float ps = __uint_as_float(pb << 16) + __uint_as_float(pb & 0xffff0000u);  // normaliser from the ROUNDED weights the numerator uses
      ps += __shfl_xor_sync(0xffffffffu, ps, 1);
      ps += __shfl_xor_sync(0xffffffffu, ps, 2);
      l_run = l_run * corr + ps;
      m_run = m_new;
      if (eb != e0) {  // rescale what earlier tiles of this node accumulated: D columns are heads 2t, 2t+1, whose factors live in rows 2t, 2t+1
        const float ca = __shfl_sync(0xffffffffu, corr, 8 * t), cb = __shfl_sync(0xffffffffu, corr, 8 * t + 4);
#pragma unroll
        for (int i = 0; i < kCW / 16; ++i) D[i][0] *= ca, D[i][1] *= cb, D[i][2] *= ca, D[i][3] *= cb;
        Dab[0] *= ca, Dab[1] *= cb, Dab[2] *= ca, Dab[3] *= cb;
      }
      // ---- D[channel, head] += V^T P ----
      const uint32_t row_v = row_k + 512u;
#pragma unroll
      for (int ii = 0; ii < kCW / 32; ++ii) {
        uint32_t a[4];
        ldmatrix_x4_trans(a, row_v + ii * 64);
        mma_bf16_k8(D[2 * ii], a[0], a[1], pb);
        mma_bf16_k8(D[2 * ii + 1], a[2], a[3], pb);
      }
      {  // abar^T[attribute, head] += attr^T P: A[row = attribute g (+8), k = edges 2t, 2t+1]
        int p0 = cpos + 2 * t, p1 = cpos + 2 * t + 1;
        if (p0 >= kMSlots) p0 -= kMSlots;
        if (p1 >= kMSlots) p1 -= kMSlots;
        const uint32_t r0 = ring + (uint32_t)(p0 * kSlotBytes) + 1024u + (uint32_t)(g * 4), r1 = ring + (uint32_t)(p1 * kSlotBytes) + 1024u + (uint32_t)(g * 4);
        const uint32_t a0 = pack_bf16x2(__uint_as_float(lds32(r0)), __uint_as_float(lds32(r1)));
        const uint32_t a1 = pack_bf16x2(__uint_as_float(lds32(r0 + 32)), __uint_as_float(lds32(r1 + 32)));
        mma_bf16_k8(Dab, a0, a1, pb);
      }
      cpos += nb;
      if (cpos >= kMSlots) cpos -= kMSlots;
      __syncwarp();  // every lane is done with these slots before they are refilled
      for (int i = 0; i < nb; ++i) issue();
    }

    // ---- finalise node d ----
    const bool has_edges = e1 > e0;
    const float il = has_edges ? 1.0f / l_run : 0.f;                     // head g
    const float ila = __shfl_sync(0xffffffffu, il, 8 * t), ilb = __shfl_sync(0xffffffffu, il, 8 * t + 4);  // heads 2t, 2t+1
    __syncwarp();  // the previous node's staging row has been read by every lane
    // channel tile i belongs to head hi; its sums sit in column hi of D[i]: lanes with t == hi / 2, register parity hi % 2 (unnormalised)
#pragma unroll
    for (int i = 0; i < kCW / 16; ++i) {
      const int hi = (i * 16) / CH;
      if ((hi >> 1) == t) {
        stage_g[i * 16 + g] = D[i][hi & 1];
        stage_g[i * 16 + g + 8] = D[i][2 + (hi & 1)];
      }
    }
    {  // abar[d, head, attribute]: rows = attributes g, g + 8; columns = heads 2t, 2t + 1
      char* ab = reinterpret_cast<char*>(p.abar) + (uint64_t)(uint32_t)d * ldab_b;
      if (ab_h0 && ab_lo) *reinterpret_cast<__nv_bfloat16*>(ab + ab_off0) = __float2bfloat16(Dab[0] * ila);
      if (ab_h0 && ab_hi) *reinterpret_cast<__nv_bfloat16*>(ab + ab_off0 + 16) = __float2bfloat16(Dab[2] * ila);
      if (ab_h1 && ab_lo) *reinterpret_cast<__nv_bfloat16*>(ab + ab_off1) = __float2bfloat16(Dab[1] * ilb);
      if (ab_h1 && ab_hi) *reinterpret_cast<__nv_bfloat16*>(ab + ab_off1 + 16) = __float2bfloat16(Dab[3] * ilb);
    }
    __syncwarp();
    {  // coalesced row write: lane owns channels lane*8 .. lane*8+7 of the slab, all of head (lane * 8) / CH
      const float4 o0 = *reinterpret_cast<const float4*>(stage_g + lane * 8), o1 = *reinterpret_cast<const float4*>(stage_g + lane * 8 + 4);
      const float ilo = __shfl_sync(0xffffffffu, il, ((lane * 8) / CH) * 4);  // normaliser of that head (row (lane*8)/CH lives in lanes 4*row..)
      float o[8] = {o0.x * ilo, o0.y * ilo, o0.z * ilo, o0.w * ilo, o1.x * ilo, o1.y * ilo, o1.z * ilo, o1.w * ilo};
      if (has_edges) {
        o[0] += be0.x, o[1] += be0.y, o[2] += be0.z, o[3] += be0.w, o[4] += be1.x, o[5] += be1.y, o[6] += be1.z, o[7] += be1.w;
      }
      if (addp) {
        const uint4 av = __ldg(reinterpret_cast<const uint4*>(addp + (uint64_t)(uint32_t)d * ldadd_b + slab_off + lane * 16));
        o[0] += __uint_as_float(av.x << 16), o[1] += __uint_as_float(av.x & 0xffff0000u);
        o[2] += __uint_as_float(av.y << 16), o[3] += __uint_as_float(av.y & 0xffff0000u);
        o[4] += __uint_as_float(av.z << 16), o[5] += __uint_as_float(av.z & 0xffff0000u);
        o[6] += __uint_as_float(av.w << 16), o[7] += __uint_as_float(av.w & 0xffff0000u);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<char*>(p.out) + (uint64_t)(uint32_t)d * ldo_b + slab_off + lane * 16) =
          make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
  }
  cp_async_wait<0>();
}

template <int CH>
int launch_ch(const AttnParams& p, int n_slabs, cudaStream_t s) {
  constexpr int smem = 4 * kWarpSmem;
  static int blocks_per_sm_dev[kMaxDevices] = {};
  int& blocks_per_sm = blocks_per_sm_dev[current_device()];
  if (blocks_per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(gt_attention_mma_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gt_attention_mma_kernel)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, gt_attention_mma_kernel<CH>, 128, smem);
    if (e != cudaSuccess || blocks_per_sm < 1) blocks_per_sm = 1;
  }
  int64_t blocks = (int64_t)num_sms() * blocks_per_sm;
  const int64_t max_useful = (p.n_dst * n_slabs + 3) / 4;
  if (blocks > max_useful) blocks = max_useful;
  while ((blocks * 4) % n_slabs) ++blocks;
  gt_attention_mma_kernel<CH><<<(unsigned)blocks, 128, smem, s>>>(p, n_slabs);
  return launch_status("gt_attention_mma_kernel");
}

}  // namespace

int launch_gt_attention_mma(const AttnParams& p, cudaStream_t s) {
  const int C = p.heads * p.ch;
  // folded form only (raw attributes padded to 16 floats, qw / abar given), bf16 rows in whole 256-channel slabs
  if (C % kCW != 0 || p.dp > kMaxEdgeDim || p.dp < 1 || !p.qw || !p.abar || !p.edge_attr) return 1;
  const int64_t max_ld = (int64_t)1 << 29;  // byte pitches are formed in 32 bits inside the kernel
  if (p.ldq >= max_ld || p.ldk >= max_ld || p.ldv >= max_ld || p.ldo >= max_ld || p.ldadd >= max_ld || p.lde >= max_ld || p.ldqw >= max_ld ||
      p.ldabar >= max_ld)
    return 1;
  if (p.ch == 32) return launch_ch<32>(p, C / kCW, s);
  if (p.ch == 64) return launch_ch<64>(p, C / kCW, s);
  return 1;
}

}  // namespace anemoi
