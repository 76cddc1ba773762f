// graphconv_fused.cu — the WHOLE GraphConv operator (layers/conv.py:66-81) as ONE kernel, for the widths where it is HBM-bound
// (C = 16 / 32 / 64; SURVEY.md 8d: 10 C^2 flop per edge against ~4 C b bytes):
//     e'[i]  = LayerNorm(edge_mlp([x_dst[dst_i] ; x_src[src_i] ; e[i]])) + e[i]          out[d] = sum over the edges i into d of e'[i]
// gather -> edge MLP (3C -> C -> ... -> C, GELU between) -> LayerNorm -> residual -> dst-segmented sum without any intermediate in HBM:
// per edge the kernel reads one e row and writes one e' row (the only compulsory traffic), the two gathered node rows come from L2.
//
// A CTA owns a contiguous, node-aligned share of the dst-sorted edge list and walks it in tiles of TE consecutive edges:
//   * the tile's operand rows [x_dst[dst] | x_src[src] | e] are gathered by cp.async into a shared-memory tile [TE, 3C] one tile AHEAD
//     of the arithmetic (two tile buffers; the indices of the tile after that are already in registers),
//   * a warp owns 16 edges through ALL layers.  bf16: mma.sync m16n8k16 with the A fragments of layer 1 from the tile (ldmatrix), the
//     weights of every layer resident in shared memory for the whole kernel, and the accumulators of layer l re-packed in registers as the
//     A fragments of layer l+1 (bias + GELU applied on the way) - the hidden activations never leave the register file;  fp32 (parity
//     mode): plain FFMA, one output column per lane, hidden rows through the warp's own tile rows,
//   * LayerNorm statistics by quad shuffles (a row lives in the four lanes of a quad), e' = LN(h) + e written IN PLACE over the e
//     columns of the tile, then copied out with 16-byte coalesced stores,
//   * the segmented sum reads e' back from the tile: one warp per destination, fp32 accumulation in edge order of the ROUNDED e' (the
//     reference scatters the stored tensor, conv.py:79), no atomics; a destination whose edge run crosses a tile boundary is carried in
//     a shared-memory fp32 row to the next tile (any in-degree works).
// tcgen05 is deliberately not used here: at these widths a 128-edge tile's MMAs are ~1 % of the tensor peak's worth of work and the
// kernel is bound by HBM and instruction issue; the chained UMMA formulation for C >= 512 does not fit on chip.
#include <type_traits>

#include "common.cuh"

namespace anemoi {
namespace gcf {

constexpr int kMaxLayers = 6;

struct Params {
  const void* x_src;
  const void* x_dst;
  const void* e;
  int64_t lds, ldd, lde;  // elements
  const void* w;          // packed weights of dtype T: layer 0 [C, 3C] row-major, then (n_layers - 1) x [C, C]
  const float* b;         // [n_layers, C] fp32
  const float* gamma;     // [C] or null
  const float* beta;      // [C] or null
  void* e_new;
  int64_t ldn;
  const int32_t* src;
  const int32_t* dst;
  const int32_t* colptr;
  void* out;
  int64_t ldo;
  int n_dst, n_edges, n_layers;
  float eps;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// Shared-memory geometry.  Row pitches are (row bytes + 16): an odd number of 16-byte units, so the eight row addresses of an ldmatrix
// (bf16) or of a quarter-warp LDS.128 (fp32) fall into eight different 16-byte bank groups.
template <typename T, int C, int TE>
struct Geo {
  static constexpr int ES = (int)sizeof(T);
  static constexpr int kTilePitch = 3 * C * ES + 16;
  static constexpr int kW0Pitch = 3 * C * ES + 16;
  static constexpr int kWhPitch = C * ES + 16;
  static constexpr int kTileBytes = TE * kTilePitch;
  static constexpr int kThreads = TE * 2;                    // one warp per 16 edges
  static constexpr int kChunksPerRow = 3 * C * ES / 16;      // 16-byte chunks of one operand row [x_dst | x_src | e]
  static constexpr int kChunksPerPart = C * ES / 16;
  // gather / copy-out mapping: kChunksPerPart consecutive threads cover one part of one row, so a pass of the CTA covers kRowsPerPass rows
  // and a tile takes kPasses passes; per pass a thread moves its 16-byte piece of each of the three parts (all indices compile-time)
  static constexpr int kRowsPerPass = kThreads / kChunksPerPart;
  static constexpr int kPasses = TE / kRowsPerPass;
  static_assert(TE % 16 == 0 && C % 16 == 0 && (kChunksPerPart & (kChunksPerPart - 1)) == 0 && kPasses >= 1 && TE % kRowsPerPass == 0, "tile geometry");
  __host__ __device__ static constexpr int weight_bytes(int n_layers) { return C * kW0Pitch + (n_layers - 1) * C * kWhPitch; }
  // layout: [weights][bias n_layers*C f32][gamma C][beta C][carry 2*C f32][dst ids 2*(TE+4) i32][tile 0][tile 1]
  static constexpr int kDstBytes = (TE + 4) * 4;  // dst id of every row of a tile; entry TE = id of the row after the tile (or -1)
  __host__ __device__ static constexpr int smem_bytes(int n_layers) {
    return weight_bytes(n_layers) + (n_layers + 4) * C * 4 + 2 * kDstBytes + 2 * kTileBytes;
  }
};

template <typename T, int C, int TE>
__global__ void __launch_bounds__(TE * 2) graphconv_fused_kernel(const Params p) {
  using G = Geo<T, C, TE>;
  constexpr bool kBf16 = !std::is_same<T, float>::value;
  constexpr int ES = G::ES;
  constexpr int NT = G::kThreads;
  extern __shared__ __align__(16) uint8_t gcf_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = p.n_layers;
  const uint32_t s_w = (uint32_t)__cvta_generic_to_shared(gcf_smem);
  const uint32_t s_b = s_w + (uint32_t)G::weight_bytes(L);
  const uint32_t s_gamma = s_b + (uint32_t)(L * C * 4);
  const uint32_t s_beta = s_gamma + C * 4;
  const uint32_t s_carry = s_beta + C * 4;
  const uint32_t s_dst = s_carry + 2 * C * 4;
  const uint32_t s_tile = s_dst + 2 * G::kDstBytes;
  __shared__ int s_range[2];
  __shared__ int s_carry_dst[2];  // destination the carry row of each parity belongs to

  // ---- this CTA's node-aligned share of the edge list --------------------------------------------------------------------
  if (warp == 0) {
    const int R = gridDim.x, r = blockIdx.x;
    auto lower_bound = [&](int x) {  // first node n with colptr[n] >= x (lane 0's result is used)
      int lo = 0, hi = p.n_dst;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p.colptr + mid) >= x) hi = mid; else lo = mid + 1;
      }
      return lo;
    };
    if (lane == 0) {
      s_range[0] = r == 0 ? 0 : lower_bound((int)((int64_t)p.n_edges * r / R));
      s_range[1] = r == R - 1 ? p.n_dst : lower_bound((int)((int64_t)p.n_edges * (r + 1) / R));
    }
  }
  // ---- resident operands: weights (re-pitched), biases, LayerNorm affine --------------------------------------------------
  {
    const char* wg = reinterpret_cast<const char*>(p.w);
    constexpr int c0 = 3 * C * ES / 16;  // 16-byte chunks per row of layer 0
    constexpr int ch = C * ES / 16;
    for (int q = tid; q < C * c0; q += NT) cp_async16(s_w + (q / c0) * G::kW0Pitch + (q % c0) * 16, wg + (int64_t)q * 16);
    const char* wh = wg + (int64_t)C * 3 * C * ES;
    const uint32_t s_wh = s_w + C * G::kW0Pitch;
    for (int q = tid; q < (L - 1) * C * ch; q += NT) cp_async16(s_wh + (q / ch) * G::kWhPitch + (q % ch) * 16, wh + (int64_t)q * 16);
    for (int q = tid; q < L * C; q += NT) sts32f(s_b + q * 4, __ldg(p.b + q));
    for (int q = tid; q < C; q += NT) {
      sts32f(s_gamma + q * 4, p.gamma ? __ldg(p.gamma + q) : 1.f);
      sts32f(s_beta + q * 4, p.beta ? __ldg(p.beta + q) : 0.f);
    }
    cp_commit();
  }
  __syncthreads();
  const int n_lo = s_range[0], n_hi = s_range[1];
  if (n_lo >= n_hi) {
    cp_wait_all();
    return;
  }
  const int E0 = __ldg(p.colptr + n_lo), E1 = __ldg(p.colptr + n_hi);
  T* const outp = reinterpret_cast<T*>(p.out);
  // destinations without edges: out = 0 (scatter into zeros, conv.py:79); 32 destinations per warp step
  for (int d0 = n_lo + warp * 32; d0 < n_hi; d0 += NT) {
    const int d = d0 + lane;
    const bool empty = d < n_hi && __ldg(p.colptr + d) == __ldg(p.colptr + d + 1);
    unsigned m = __ballot_sync(0xffffffffu, empty);
    while (m) {
      const int dz = d0 + __ffs(m) - 1;
      m &= m - 1;
      for (int c = lane; c < C; c += 32) outp[(int64_t)dz * p.ldo + c] = from_f32<T>(0.f);
    }
  }
  if (tid == 0) s_carry_dst[0] = s_carry_dst[1] = -1;
  const int n_tiles = (E1 - E0 + TE - 1) / TE;

  // ---- gather pipeline ----------------------------------------------------------------------------------------------------------
  // chunk q = tid + i * NT of a tile: row q / kChunksPerRow, part (x_dst | x_src | e) and 16-byte piece inside the part
  const char* const xs = reinterpret_cast<const char*>(p.x_src);
  const char* const xd = reinterpret_cast<const char*>(p.x_dst);
  const char* const eg = reinterpret_cast<const char*>(p.e);
  const int64_t lds_b = p.lds * ES, ldd_b = p.ldd * ES, lde_b = p.lde * ES;
  constexpr int PP = G::kChunksPerPart;
  const int g_row = tid / PP, g_piece = tid % PP;
  int gidx[G::kPasses][2];  // (dst, src) ids of this thread's rows of the tile that is gathered NEXT
  auto load_idx = [&](int t) {
    const int te0 = E0 + t * TE;
#pragma unroll
    for (int h = 0; h < G::kPasses; ++h) {
      const int ei = te0 + g_row + h * G::kRowsPerPass;
      gidx[h][0] = gidx[h][1] = 0;
      if (ei < E1) gidx[h][0] = __ldg(p.dst + ei), gidx[h][1] = __ldg(p.src + ei);
    }
  };
  auto gather = [&](int t) {
    const int te0 = E0 + t * TE;
    const uint32_t tile = s_tile + (uint32_t)((t & 1) * G::kTileBytes) + (uint32_t)(g_row * G::kTilePitch + g_piece * 16);
#pragma unroll
    for (int h = 0; h < G::kPasses; ++h) {
      const int r = g_row + h * G::kRowsPerPass;
      const int ei = te0 + r;
      if (ei < E1) {
        const uint32_t dstp = tile + (uint32_t)(h * G::kRowsPerPass * G::kTilePitch);
        cp_async16(dstp, xd + (int64_t)gidx[h][0] * ldd_b + g_piece * 16);
        cp_async16(dstp + PP * 16, xs + (int64_t)gidx[h][1] * lds_b + g_piece * 16);
        cp_async16(dstp + 2 * PP * 16, eg + (int64_t)ei * lde_b + g_piece * 16);
      }
      if (g_piece == 0) sts32(s_dst + (uint32_t)((t & 1) * G::kDstBytes + r * 4), ei < E1 ? (uint32_t)gidx[h][0] : 0xffffffffu);
    }
    cp_commit();
  };
  load_idx(0);
  gather(0);
  if (n_tiles > 1) load_idx(1);

  const int g = lane >> 2, tq = lane & 3;  // mma fragment coordinates: row group / thread in quad
  for (int t = 0; t < n_tiles; ++t) {
    const int te0 = E0 + t * TE, te1 = min(te0 + TE, E1);
    const uint32_t tile = s_tile + (uint32_t)((t & 1) * G::kTileBytes);
    cp_wait_all();
    __syncthreads();  // tile t (and, first time, the weights) has landed; everybody is done with tile t-1's buffer
    if (t + 1 < n_tiles) {
      gather(t + 1);
      if (t + 2 < n_tiles) load_idx(t + 2);
    }
    const int r0 = warp * 16;
    if (te0 + r0 < te1) {  // warps whose 16 rows lie past the end of the CTA's last tile have nothing to do
      if constexpr (kBf16) {
        // ================= bf16: mma.sync, hidden activations in registers =================
        float acc[C / 8][4];
        uint32_t afr[C / 16][4];
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
          const uint32_t wl = l == 0 ? s_w : s_w + C * G::kW0Pitch + (uint32_t)((l - 1) * C * G::kWhPitch);
          const int wp = l == 0 ? G::kW0Pitch : G::kWhPitch;
          const uint32_t wlane = wl + (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * wp + ((lane >> 3) & 1) * 16);
          const uint32_t bl = s_b + (uint32_t)(l * C * 4) + tq * 8;
#pragma unroll
          for (int j = 0; j < C / 8; ++j) {  // the accumulators start from the bias (no add in the epilogue)
            const float2 bb = lds64f(bl + j * 32);
            acc[j][0] = acc[j][2] = bb.x, acc[j][1] = acc[j][3] = bb.y;
          }
          if (l == 0) {
            const uint32_t alane = tile + (uint32_t)((r0 + (lane & 15)) * G::kTilePitch + (lane >> 4) * 16);
#pragma unroll
            for (int kk = 0; kk < 3 * C / 16; ++kk) {
              uint32_t a[4];
              ldsm_x4(a, alane + kk * 32);
#pragma unroll
              for (int jp = 0; jp < C / 16; ++jp) {
                uint32_t b[4];
                ldsm_x4(b, wlane + (uint32_t)(jp * 16 * G::kW0Pitch + kk * 32));
                mma_bf16(acc[2 * jp], a, b[0], b[1]);
                mma_bf16(acc[2 * jp + 1], a, b[2], b[3]);
              }
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < C / 16; ++kk) {
#pragma unroll
              for (int jp = 0; jp < C / 16; ++jp) {
                uint32_t b[4];
                ldsm_x4(b, wlane + (uint32_t)(jp * 16 * G::kWhPitch + kk * 32));
                mma_bf16(acc[2 * jp], afr[kk], b[0], b[1]);
                mma_bf16(acc[2 * jp + 1], afr[kk], b[2], b[3]);
              }
            }
          }
          if (l + 1 < L) {  // GELU, re-packed as the next layer's A fragments
#pragma unroll
            for (int j = 0; j < C / 8; ++j) {
              const float2 lo = gelu_erf_fast2(make_float2(acc[j][0], acc[j][1]));
              const float2 hi = gelu_erf_fast2(make_float2(acc[j][2], acc[j][3]));
              afr[j >> 1][(j & 1) * 2] = pack_bf16x2(lo.x, lo.y);
              afr[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(hi.x, hi.y);
            }
          } else {  // rounding to the storage dtype (the LayerNorm of the reference reads the bf16 Linear output)
#pragma unroll
            for (int j = 0; j < C / 8; ++j) {
              const float2 lo = unpack_bf16x2(pack_bf16x2(acc[j][0], acc[j][1]));
              const float2 hi = unpack_bf16x2(pack_bf16x2(acc[j][2], acc[j][3]));
              acc[j][0] = lo.x, acc[j][1] = lo.y, acc[j][2] = hi.x, acc[j][3] = hi.y;
            }
          }
        }
        // LayerNorm over the row (two-pass, the row lives in the 4 lanes of a quad), + e, written in place over the e columns
        // (packed fp32x2 arithmetic throughout: the kernel is bound by instruction issue, two columns per issue slot)
        float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < C / 8; ++j) {
          sa = __fadd2_rn(sa, make_float2(acc[j][0], acc[j][1]));
          sb = __fadd2_rn(sb, make_float2(acc[j][2], acc[j][3]));
        }
        float s0 = sa.x + sa.y, s1 = sb.x + sb.y;
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1), s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2), s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float m0 = s0 * (1.0f / C), m1 = s1 * (1.0f / C);
        const float2 nm0 = make_float2(-m0, -m0), nm1 = make_float2(-m1, -m1);
        float2 qa = make_float2(0.f, 0.f), qb = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < C / 8; ++j) {
          const float2 da = __fadd2_rn(make_float2(acc[j][0], acc[j][1]), nm0), db = __fadd2_rn(make_float2(acc[j][2], acc[j][3]), nm1);
          qa = __ffma2_rn(da, da, qa), qb = __ffma2_rn(db, db, qb);
        }
        float q0 = qa.x + qa.y, q1 = qb.x + qb.y;
        q0 += __shfl_xor_sync(0xffffffffu, q0, 1), q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
        q0 += __shfl_xor_sync(0xffffffffu, q0, 2), q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
        const float rs0 = rsqrtf(q0 * (1.0f / C) + p.eps), rs1 = rsqrtf(q1 * (1.0f / C) + p.eps);
        const float2 r0v = make_float2(rs0, rs0), r1v = make_float2(rs1, rs1);
        const uint32_t erow0 = tile + (uint32_t)((r0 + g) * G::kTilePitch + 2 * C * ES + tq * 4), erow1 = erow0 + 8 * G::kTilePitch;
#pragma unroll
        for (int j = 0; j < C / 8; ++j) {
          // (x - m) rstd gamma + beta + e  =  x * (rstd gamma) + (beta + e - m rstd gamma): four packed operations per column pair and row
          const float2 gm = lds64f(s_gamma + tq * 8 + j * 32), bt = lds64f(s_beta + tq * 8 + j * 32);
          const float2 e0 = unpack_bf16x2(lds32(erow0 + j * 16)), e1 = unpack_bf16x2(lds32(erow1 + j * 16));
          const float2 sc0 = __fmul2_rn(gm, r0v), sc1 = __fmul2_rn(gm, r1v);
          const float2 c0 = __ffma2_rn(nm0, sc0, __fadd2_rn(bt, e0)), c1 = __ffma2_rn(nm1, sc1, __fadd2_rn(bt, e1));
          const float2 y0 = __ffma2_rn(make_float2(acc[j][0], acc[j][1]), sc0, c0), y1 = __ffma2_rn(make_float2(acc[j][2], acc[j][3]), sc1, c1);
          sts32(erow0 + j * 16, pack_bf16x2(y0.x, y0.y));
          sts32(erow1 + j * 16, pack_bf16x2(y1.x, y1.y));
        }
      } else {
        // ================= fp32 parity mode: FFMA, lane = output column =================
        constexpr int NPL = C >= 32 ? C / 32 : 1;  // columns per lane
        constexpr int RPL = C >= 32 ? 16 : 8;      // rows per lane (C = 16: the two half-warps split the 16 rows)
        const int rb = C >= 32 ? 0 : (lane >> 4) * 8;
        const int n0 = C >= 32 ? lane : (lane & 15);
        float acc[NPL][RPL];
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
          const uint32_t wl = l == 0 ? s_w : s_w + C * G::kW0Pitch + (uint32_t)((l - 1) * C * G::kWhPitch);
          const int wp = l == 0 ? G::kW0Pitch : G::kWhPitch;
          const int K4 = (l == 0 ? 3 * C : C) / 4;
#pragma unroll
          for (int i = 0; i < NPL; ++i) {
            const float bb = lds32f(s_b + (uint32_t)((l * C + n0 + 32 * i) * 4));
#pragma unroll
            for (int r = 0; r < RPL; ++r) acc[i][r] = bb;
          }
          const uint32_t arow = tile + (uint32_t)((r0 + rb) * G::kTilePitch);
#pragma unroll 2
          for (int k4 = 0; k4 < K4; ++k4) {
            uint4 wv[NPL];
#pragma unroll
            for (int i = 0; i < NPL; ++i) wv[i] = lds128(wl + (uint32_t)((n0 + 32 * i) * wp + k4 * 16));
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
              const uint4 av = lds128(arow + (uint32_t)(r * G::kTilePitch + k4 * 16));
#pragma unroll
              for (int i = 0; i < NPL; ++i) {
                acc[i][r] = fmaf(__uint_as_float(av.x), __uint_as_float(wv[i].x), acc[i][r]);
                acc[i][r] = fmaf(__uint_as_float(av.y), __uint_as_float(wv[i].y), acc[i][r]);
                acc[i][r] = fmaf(__uint_as_float(av.z), __uint_as_float(wv[i].z), acc[i][r]);
                acc[i][r] = fmaf(__uint_as_float(av.w), __uint_as_float(wv[i].w), acc[i][r]);
              }
            }
          }
          if (l + 1 < L) {  // hidden row -> columns [0, C) of the warp's own tile rows (x_dst is no longer needed)
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NPL; ++i)
#pragma unroll
              for (int r = 0; r < RPL; ++r) sts32f(arow + (uint32_t)(r * G::kTilePitch + (n0 + 32 * i) * 4), gelu_erf(acc[i][r]));
            __syncwarp();
          }
        }
        float gm[NPL], bt[NPL];
#pragma unroll
        for (int i = 0; i < NPL; ++i) gm[i] = lds32f(s_gamma + (n0 + 32 * i) * 4), bt[i] = lds32f(s_beta + (n0 + 32 * i) * 4);
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < NPL; ++i) s += acc[i][r];
          if (C >= 32) s += __shfl_xor_sync(0xffffffffu, s, 16);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = s * (1.0f / C);
          float qq = 0.f;
#pragma unroll
          for (int i = 0; i < NPL; ++i) qq += (acc[i][r] - mean) * (acc[i][r] - mean);
          if (C >= 32) qq += __shfl_xor_sync(0xffffffffu, qq, 16);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
          const float rstd = rsqrtf(qq * (1.0f / C) + p.eps);
          const uint32_t erow = tile + (uint32_t)((r0 + rb + r) * G::kTilePitch + 2 * C * 4);
#pragma unroll
          for (int i = 0; i < NPL; ++i) {
            const uint32_t a = erow + (uint32_t)((n0 + 32 * i) * 4);
            sts32f(a, (acc[i][r] - mean) * rstd * gm[i] + bt[i] + lds32f(a));
          }
        }
      }
    }
    __syncthreads();  // e' of the whole tile is in place
    // ---- copy e' out (16-byte coalesced stores, same thread -> (row, piece) map as the gather) ----
    {
      const int nvalid = te1 - te0;
      char* const eo = reinterpret_cast<char*>(p.e_new) + (int64_t)(te0 + g_row) * p.ldn * ES + g_piece * 16;
      const uint32_t from = tile + (uint32_t)(g_row * G::kTilePitch + 2 * C * ES + g_piece * 16);
#pragma unroll
      for (int h = 0; h < G::kPasses; ++h) {
        if (g_row + h * G::kRowsPerPass < nvalid)
          *reinterpret_cast<uint4*>(eo + (int64_t)h * G::kRowsPerPass * p.ldn * ES) = lds128(from + (uint32_t)(h * G::kRowsPerPass * G::kTilePitch));
      }
    }
    // ---- segmented sum ----
    // Runs of equal dst ids are found from the tile's ids in shared memory (no global loads on this path).  Warp w sums the runs that START in
    // its 16 rows, to their end inside the tile; a run that reaches the tile's end and continues in the next tile (the id of the row after
    // the tile is the first id of the other buffer, already gathered) is parked in the fp32 carry row of the next parity with its owner id.
    {
      const uint32_t ids = s_dst + (uint32_t)((t & 1) * G::kDstBytes);
      const int nvalid = te1 - te0;
      const int next_first = t + 1 < n_tiles ? (int)lds32(s_dst + (uint32_t)(((t + 1) & 1) * G::kDstBytes)) : -1;
      const uint32_t carry_in = s_carry + (uint32_t)((t & 1) * C * 4), carry_out = s_carry + (uint32_t)(((t + 1) & 1) * C * 4);
      const int row = warp * 16 + (lane & 15);
      const int my = row < nvalid ? (int)lds32(ids + row * 4) : -2;
      const int before = row == 0 ? -3 : (row <= nvalid ? (int)lds32(ids + (row - 1) * 4) : -2);
      unsigned starts = __ballot_sync(0xffffffffu, lane < 16 && row < nvalid && my != before);
      while (starts) {
        const int rs = warp * 16 + __ffs(starts) - 1;
        starts &= starts - 1;
        const int d = (int)lds32(ids + rs * 4);
        int re = rs + 1;  // first row after the run
        for (;;) {
          const int rr = re + lane;
          const unsigned diff = __ballot_sync(0xffffffffu, rr >= nvalid || (int)lds32(ids + rr * 4) != d);
          if (diff) {
            re += __ffs(diff) - 1;
            break;
          }
          re += 32;
        }
        const bool from_carry = rs == 0 && s_carry_dst[t & 1] == d;
        const bool to_carry = re == nvalid && next_first == d;
        // LPR lanes cover a row's column pairs, the warp's 32 / LPR lane groups take every (32 / LPR)-th row of the run; the group sums are
        // then added across groups (fixed order: deterministic, fp32)
        constexpr int LPR = C / 2 < 32 ? C / 2 : 32, RW = 32 / LPR;
        const int cp = lane % LPR, grp = lane / LPR;
        float sx = 0.f, sy = 0.f;
        if (from_carry && grp == 0) {
          const float2 cv = lds64f(carry_in + cp * 8);
          sx = cv.x, sy = cv.y;
        }
        const uint32_t col = tile + (uint32_t)(2 * C * ES + cp * 2 * ES);
#pragma unroll 2
        for (int r = rs + grp; r < re; r += RW) {
          const uint32_t addr = col + (uint32_t)(r * G::kTilePitch);
          if constexpr (kBf16) {
            const float2 v = unpack_bf16x2(lds32(addr));
            sx += v.x, sy += v.y;
          } else {
            const float2 v = lds64f(addr);
            sx += v.x, sy += v.y;
          }
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o), sy += __shfl_xor_sync(0xffffffffu, sy, o);
        if (grp == 0) {
          if (to_carry) {
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(carry_out + cp * 8), "f"(sx), "f"(sy) : "memory");
          } else {
            T* o = outp + (int64_t)d * p.ldo + cp * 2;
            if constexpr (kBf16)
              *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(sx, sy);
            else
              *reinterpret_cast<float2*>(o) = make_float2(sx, sy);
          }
        }
        if (to_carry && lane == 0) s_carry_dst[(t + 1) & 1] = d;
      }
    }
  }
}

template <typename T, int C, int TE>
int launch(const Params& p, cudaStream_t s) {
  using G = Geo<T, C, TE>;
  const int smem = G::smem_bytes(p.n_layers);
  if (smem > 227 * 1024) {
    set_error("graphconv_fused: %d layers of width %d need %d bytes of shared memory", p.n_layers, C, smem);
    return -3;
  }
  static int cached_smem[kMaxDevices] = {};
  static int blocks_per_sm[kMaxDevices] = {};
  const int dev = current_device();
  if (cached_smem[dev] != smem) {
    ANEMOI_CUDA(cudaFuncSetAttribute(graphconv_fused_kernel<T, C, TE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ANEMOI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[dev], graphconv_fused_kernel<T, C, TE>, G::kThreads, smem));
    if (blocks_per_sm[dev] < 1) blocks_per_sm[dev] = 1;
    cached_smem[dev] = smem;
  }
  // every CTA should see at least ~4 tiles, so that the resident weights are amortised and the pipeline has something to overlap
  const int64_t want = (p.n_edges + 4 * TE - 1) / (4 * TE);
  int64_t blocks = (int64_t)num_sms() * blocks_per_sm[dev];
  if (blocks > want) blocks = want;
  if (blocks < 1) blocks = 1;
  graphconv_fused_kernel<T, C, TE><<<(unsigned)blocks, G::kThreads, smem, s>>>(p);
  return launch_status("graphconv_fused_kernel");
}

}  // namespace gcf
}  // namespace anemoi

using namespace anemoi;

extern "C" int anemoi_b200_graphconv_fused(const void* x_src, int64_t lds, const void* x_dst, int64_t ldd, const void* e, int64_t lde,
                                           const void* weights, const float* biases, int64_t n_layers, const float* gamma, const float* beta,
                                           void* e_new, int64_t ldn, const int32_t* src32, const int32_t* dst32, const int32_t* colptr32, void* out,
                                           int64_t ldo, int64_t n_dst, int64_t n_edges, int64_t C, float eps, int dtype, void* stream) {
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "graphconv_fused: bad dtype %d", dtype);
  ANEMOI_CHECK_ARG(n_dst >= 0 && n_edges >= 0 && n_edges < (int64_t)1 << 31 && n_dst < (int64_t)1 << 31, "graphconv_fused: bad sizes");
  if (C != 16 && C != 32 && C != 64) {
    set_error("graphconv_fused: C = %lld is not one of 16 / 32 / 64 (use the decomposed path)", (long long)C);
    return -3;
  }
  if (n_layers < 2 || n_layers > gcf::kMaxLayers) {
    set_error("graphconv_fused: %lld layers (supported: 2 .. %d)", (long long)n_layers, gcf::kMaxLayers);
    return -3;
  }
  if (n_dst == 0) return 0;
  ANEMOI_CHECK_ARG(colptr32 && out && weights && biases, "graphconv_fused: null pointer");
  ANEMOI_CHECK_ARG(ldo >= C, "graphconv_fused: ldo too small");
  const int es = dtype == ANEMOI_BF16 ? 2 : 4;
  auto al = [&](const void* ptr, int64_t ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * es) % 16 == 0; };
  if (n_edges > 0) {
    ANEMOI_CHECK_ARG(x_src && x_dst && e && e_new && src32 && dst32, "graphconv_fused: null pointer");
    ANEMOI_CHECK_ARG(lds >= C && ldd >= C && lde >= C && ldn >= C, "graphconv_fused: leading dimension too small");
    if (!(al(x_src, lds) && al(x_dst, ldd) && al(e, lde) && al(e_new, ldn) && al(weights, 8))) {
      set_error("graphconv_fused: operands must be 16-byte aligned with 16-byte row pitches");
      return -3;
    }
  }
  if ((reinterpret_cast<uintptr_t>(out) & 7) != 0 || (ldo * es) % 8 != 0) {
    set_error("graphconv_fused: out must be 8-byte aligned with an 8-byte row pitch");
    return -3;
  }
  gcf::Params p{x_src, x_dst, e, lds, ldd, lde, weights, biases, gamma, beta, e_new, ldn, src32, dst32, colptr32, out, ldo,
                (int)n_dst, (int)n_edges, (int)n_layers, eps};
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ANEMOI_BF16) {
    if (C == 16) return gcf::launch<__nv_bfloat16, 16, 128>(p, s);
    if (C == 32) return gcf::launch<__nv_bfloat16, 32, 128>(p, s);
    return gcf::launch<__nv_bfloat16, 64, 128>(p, s);
  }
  if (C == 16) return gcf::launch<float, 16, 128>(p, s);
  if (C == 32) return gcf::launch<float, 32, 128>(p, s);
  return gcf::launch<float, 64, 64>(p, s);
}
