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
#include "common.cuh"

namespace anemoi {

constexpr int kMaxEdgeDim = 16;

struct AttnParams {
  const void *q, *k, *v, *e, *add;
  void* out;
  int64_t ldq, ldk, ldv, lde_proj, ldadd, ldo;
  const float* edge_attr;  // [E, lde] fp32
  int64_t lde;
  int edge_dim;
  const float* w_edge;  // [H*Ch, ldw_e] fp32
  int64_t ldw_e;
  const float* b_edge;  // [H*Ch] or null
  const int32_t* src;
  const int32_t* colptr;
  int64_t n_dst;
  int heads, ch;
  float scale;
};

// MODE 0: no edge term; 1: materialised eproj; 2: fused lin_edge from raw attributes.
template <typename T, int VEC, int MODE>
__global__ void __launch_bounds__(256) gt_attention_warp_kernel(const AttnParams p) {
  extern __shared__ float s_w[];  // MODE 2: W_e re-laid out as [(j*VEC + c)*32 + lane] (bank-conflict free), then b_e as [c*32+lane]
  const int lane = threadIdx.x & 31;
  const int D = p.edge_dim;
  if constexpr (MODE == 2) {
    const int C = 32 * VEC;
    for (int i = threadIdx.x; i < C * D; i += blockDim.x) {
      const int ch = i / D, j = i - ch * D;
      s_w[(j * VEC + (ch % VEC)) * 32 + ch / VEC] = p.w_edge[(int64_t)ch * p.ldw_e + j];
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_w[C * D + (i % VEC) * 32 + i / VEC] = p.b_edge ? p.b_edge[i] : 0.f;
    __syncthreads();
  }
  const int lph = p.ch / VEC;  // lanes per head (power of two)
  const T* __restrict__ qp = reinterpret_cast<const T*>(p.q);
  const T* __restrict__ kp = reinterpret_cast<const T*>(p.k);
  const T* __restrict__ vp = reinterpret_cast<const T*>(p.v);
  const T* __restrict__ ep = reinterpret_cast<const T*>(p.e);
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t d = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < p.n_dst; d += warps_total) {
    const int e0 = p.colptr[d], e1 = p.colptr[d + 1];
    float q[VEC], acc[VEC];
    load_vec_f32<T, VEC>(qp + d * p.ldq + lane * VEC, q);
#pragma unroll
    for (int c = 0; c < VEC; ++c) q[c] *= p.scale, acc[c] = 0.f;
    float qw[kMaxEdgeDim], abar[kMaxEdgeDim];
    if constexpr (MODE == 2) {
#pragma unroll
      for (int j = 0; j < kMaxEdgeDim; ++j) {
        qw[j] = 0.f, abar[j] = 0.f;
        if (j < D) {
          float t = 0.f;
#pragma unroll
          for (int c = 0; c < VEC; ++c) t += q[c] * s_w[(j * VEC + c) * 32 + lane];
          for (int o = 1; o < lph; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          qw[j] = t;
        }
      }
    }
    float m_i = -INFINITY, l_i = 0.f;
    for (int eb = e0; eb < e1; eb += 2) {
      // two edges per iteration: independent gathers in flight, one rescale of the accumulators
      const bool two = eb + 1 < e1;
      const int s0 = p.src[eb], s1 = two ? p.src[eb + 1] : s0;
      float k0[VEC], k1[VEC], v0[VEC], v1[VEC];
      load_vec_f32<T, VEC>(kp + (int64_t)s0 * p.ldk + lane * VEC, k0);
      load_vec_f32<T, VEC>(kp + (int64_t)s1 * p.ldk + lane * VEC, k1);
      load_vec_f32<T, VEC>(vp + (int64_t)s0 * p.ldv + lane * VEC, v0);
      load_vec_f32<T, VEC>(vp + (int64_t)s1 * p.ldv + lane * VEC, v1);
      float a0[kMaxEdgeDim], a1[kMaxEdgeDim];
      if constexpr (MODE == 1) {
        float t0[VEC], t1[VEC];
        load_vec_f32<T, VEC>(ep + (int64_t)eb * p.lde_proj + lane * VEC, t0);
        load_vec_f32<T, VEC>(ep + (int64_t)(two ? eb + 1 : eb) * p.lde_proj + lane * VEC, t1);
#pragma unroll
        for (int c = 0; c < VEC; ++c) k0[c] += t0[c], v0[c] += t0[c], k1[c] += t1[c], v1[c] += t1[c];
      }
      if constexpr (MODE == 2) {
        const float* ap0 = p.edge_attr + (int64_t)eb * p.lde;
        const float* ap1 = p.edge_attr + (int64_t)(two ? eb + 1 : eb) * p.lde;
#pragma unroll
        for (int j = 0; j < kMaxEdgeDim; j += 4) {
          if (j < D) {  // lde is a multiple of 4 and rows are zero-padded (host contract)
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(ap0 + j));
            const float4 t1 = __ldg(reinterpret_cast<const float4*>(ap1 + j));
            a0[j] = t0.x, a0[j + 1] = t0.y, a0[j + 2] = t0.z, a0[j + 3] = t0.w;
            a1[j] = t1.x, a1[j + 1] = t1.y, a1[j + 2] = t1.z, a1[j + 3] = t1.w;
          } else {
            a0[j] = a0[j + 1] = a0[j + 2] = a0[j + 3] = 0.f;
            a1[j] = a1[j + 1] = a1[j + 2] = a1[j + 3] = 0.f;
          }
        }
      }
      float sc0 = 0.f, sc1 = 0.f;
#pragma unroll
      for (int c = 0; c < VEC; ++c) sc0 += q[c] * k0[c], sc1 += q[c] * k1[c];
      for (int o = 1; o < lph; o <<= 1) {
        sc0 += __shfl_xor_sync(0xffffffffu, sc0, o);
        sc1 += __shfl_xor_sync(0xffffffffu, sc1, o);
      }
      if constexpr (MODE == 2) {
#pragma unroll
        for (int j = 0; j < kMaxEdgeDim; ++j)
          if (j < D) sc0 += qw[j] * a0[j], sc1 += qw[j] * a1[j];
      }
      if (!two) sc1 = -INFINITY;
      const float m_new = fmaxf(m_i, fmaxf(sc0, sc1));
      const float corr = __expf(m_i - m_new);  // exp(-inf) = 0 on the first iteration
      const float w0 = __expf(sc0 - m_new), w1 = __expf(sc1 - m_new);
      l_i = l_i * corr + w0 + w1;
      m_i = m_new;
#pragma unroll
      for (int c = 0; c < VEC; ++c) acc[c] = acc[c] * corr + w0 * v0[c] + w1 * v1[c];
      if constexpr (MODE == 2) {
#pragma unroll
        for (int j = 0; j < kMaxEdgeDim; ++j)
          if (j < D) abar[j] = abar[j] * corr + w0 * a0[j] + w1 * a1[j];
      }
    }
    float o[VEC];
    if (e1 > e0) {
      const float inv = 1.0f / l_i;
#pragma unroll
      for (int c = 0; c < VEC; ++c) o[c] = acc[c] * inv;
      if constexpr (MODE == 2) {
        const int C = 32 * VEC;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          float t = s_w[C * D + c * 32 + lane];
#pragma unroll
          for (int j = 0; j < kMaxEdgeDim; ++j)
            if (j < D) t += s_w[(j * VEC + c) * 32 + lane] * (abar[j] * inv);
          o[c] += t;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < VEC; ++c) o[c] = 0.f;  // zero in-degree (gt.py:112-119)
    }
    if (p.add) {
      float r[VEC];
      load_vec_f32<T, VEC>(reinterpret_cast<const T*>(p.add) + d * p.ldadd + lane * VEC, r);
#pragma unroll
      for (int c = 0; c < VEC; ++c) o[c] += r[c];
    }
    store_vec_f32<T, VEC>(reinterpret_cast<T*>(p.out) + d * p.ldo + lane * VEC, o);
  }
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

template <typename T, int VEC>
static int launch_warp(const AttnParams& p, int mode, cudaStream_t s) {
  const int warps_per_block = 8;
  int64_t blocks = (p.n_dst + warps_per_block - 1) / warps_per_block;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = mode == 2 ? (size_t)(32 * VEC) * (p.edge_dim + 1) * sizeof(float) : 0;
  if (mode == 0) {
    gt_attention_warp_kernel<T, VEC, 0><<<(unsigned)blocks, 256, 0, s>>>(p);
  } else if (mode == 1) {
    gt_attention_warp_kernel<T, VEC, 1><<<(unsigned)blocks, 256, 0, s>>>(p);
  } else {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(gt_attention_warp_kernel<T, VEC, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attention)");
    }
    gt_attention_warp_kernel<T, VEC, 2><<<(unsigned)blocks, 256, smem, s>>>(p);
  }
  return launch_status("gt_attention_warp_kernel");
}

template <typename T>
static int dispatch(const AttnParams& p, int mode, bool aligned, cudaStream_t s) {
  const int C = p.heads * p.ch;
  if (aligned && C % 32 == 0) {
    const int vec = C / 32;
    const bool ok = p.ch % vec == 0 && ((p.ch / vec) & (p.ch / vec - 1)) == 0 && p.ch / vec <= 32;
    if (ok) {
      switch (vec) {
        case 2: return launch_warp<T, 2>(p, mode, s);
        case 4: return launch_warp<T, 4>(p, mode, s);
        case 8: return launch_warp<T, 8>(p, mode, s);
        case 16: return launch_warp<T, 16>(p, mode, s);
        case 32: return launch_warp<T, 32>(p, mode, s);
        default: break;
      }
    }
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
                                            int64_t ldw_e, const float* b_edge, const int32_t* src32, const int32_t* colptr32, const void* add,
                                            int64_t ldadd, void* out, int64_t ldo, int64_t n_dst, int64_t heads, int64_t ch, int dtype,
                                            void* stream) {
  ANEMOI_CHECK_ARG(n_dst >= 0 && heads >= 1 && ch >= 1, "gt_attention: bad shape");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "gt_attention: bad dtype %d", dtype);
  if (n_dst == 0) return 0;
  ANEMOI_CHECK_ARG(q && k && v && out && src32 && colptr32, "gt_attention: null pointer");
  ANEMOI_CHECK_ARG(!(e && edge_attr), "gt_attention: pass either a materialised edge projection or raw edge attributes, not both");
  const int64_t C = heads * ch;
  ANEMOI_CHECK_ARG(ldq >= C && ldk >= C && ldv >= C && ldo >= C, "gt_attention: leading dimension < heads*ch");
  int mode = e ? 1 : (edge_attr ? 2 : 0);
  if (mode == 2) {
    ANEMOI_CHECK_ARG(w_edge && edge_dim >= 1 && lde >= edge_dim && ldw_e >= edge_dim, "gt_attention: bad fused lin_edge arguments");
  }
  AttnParams p;
  p.q = q, p.k = k, p.v = v, p.e = e, p.add = add, p.out = out;
  p.ldq = ldq, p.ldk = ldk, p.ldv = ldv, p.lde_proj = lde_proj, p.ldadd = ldadd, p.ldo = ldo;
  p.edge_attr = edge_attr, p.lde = lde, p.edge_dim = (int)edge_dim, p.w_edge = w_edge, p.ldw_e = ldw_e, p.b_edge = b_edge;
  p.src = src32, p.colptr = colptr32, p.n_dst = n_dst, p.heads = (int)heads, p.ch = (int)ch;
  p.scale = 1.0f / sqrtf((float)ch);
  const int es = dtype == ANEMOI_BF16 ? 2 : 4;
  auto al = [&](const void* ptr, int64_t ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * es) % 16 == 0); };
  bool aligned = al(q, ldq) && al(k, ldk) && al(v, ldv) && al(e, lde_proj) && al(add, ldadd) && al(out, ldo);
  if (mode == 2) {
    // fast fused path contract: <= 16 attributes, rows padded to a multiple of 4 floats and 16-byte aligned
    const bool fused_ok = edge_dim <= kMaxEdgeDim && lde % 4 == 0 && lde >= ((edge_dim + 3) / 4) * 4 &&
                          (reinterpret_cast<uintptr_t>(edge_attr) & 15) == 0;
    aligned = aligned && fused_ok;
  }
  cudaStream_t s = (cudaStream_t)stream;
  return dtype == ANEMOI_BF16 ? dispatch<__nv_bfloat16>(p, mode, aligned, s) : dispatch<float>(p, mode, aligned, s);
}
