// backward.cu — backward of the fused forward ops, so that the drop-in modules train under anemoi-training.
//
//   gt_attention_bwd        replaces triton/gt.py:182-376 (`_gt_bwd_dst_pass`, `_gt_bwd_src_pass`) and :451-556 (autograd wiring): two
//                           deterministic passes over the cached CSR, no atomics.
//                             pass 1 (dst-major, one warp per (dst, head)): with D = dout . out and alpha_e = exp(s_e - lse),
//                                 ds_e = scale * alpha_e (dout . (v_s + e_e) - D),  dq = sum_e ds_e (k_s + e_e),  de_e = ds_e q + alpha_e dout,
//                                 and (alpha_e, ds_e) per (edge, head) to a scratch buffer;
//                             pass 2 (src-major over the reverse CSR, one warp per (src, head)): dk_s = sum ds_e q_dst(e), dv_s = sum alpha_e dout_dst(e).
//   layer_norm_bwd          dx, dgamma, dbeta of LayerNorm; with the optional gathered second cotangent (g = dy[i] + dz[idx[i]]) it is ALSO the
//                           backward of the GraphConv tail e' = LN(h) + e, out[d] = sum e' (layers/conv.py:73-81): dz = d out, idx = dst.
//   gelu fwd / bwd          the exact-erf GELU as a separate element-wise op (training keeps the pre-activation).
// All accumulation in fp32.
#include "common.cuh"

namespace anemoi {
namespace {

constexpr int kMaxV = 8;  // channels per lane: heads up to 256 channels

template <typename T>
__global__ void __launch_bounds__(256) gt_attention_bwd_dst_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
                                                                   const T* __restrict__ v, int64_t ldv, const T* __restrict__ e, int64_t lde,
                                                                   const T* __restrict__ out, int64_t ldo, const T* __restrict__ dout, int64_t lddo,
                                                                   const float* __restrict__ lse, const int32_t* __restrict__ src,
                                                                   const int32_t* __restrict__ colptr, T* __restrict__ dq, int64_t lddq,
                                                                   T* __restrict__ de, int64_t ldde, float* __restrict__ alpha,
                                                                   float* __restrict__ ds, int64_t n_dst, int heads, int ch, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t items = n_dst * heads;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < items; w += warps_total) {
    const int64_t d = w / heads;
    const int h = (int)(w - d * heads);
    const int base = h * ch;
    const int e0 = colptr[d], e1 = colptr[d + 1];
    float qv[kMaxV], gv[kMaxV], acc[kMaxV];
    float dsum = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      qv[i] = gv[i] = acc[i] = 0.f;
      if (c < ch) {
        qv[i] = to_f32<T>(q[d * ldq + base + c]);
        gv[i] = to_f32<T>(dout[d * lddo + base + c]);
        dsum += gv[i] * to_f32<T>(out[d * ldo + base + c]);
      }
    }
    dsum = warp_sum(dsum);
    const float l = lse[d * heads + h];
    for (int ei = e0; ei < e1; ++ei) {
      const int s = src[ei];
      float kk[kMaxV], sc = 0.f, da = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxV; ++i) {
        const int c = lane + 32 * i;
        kk[i] = 0.f;
        if (c < ch) {
          const float ee = e ? to_f32<T>(e[(int64_t)ei * lde + base + c]) : 0.f;
          kk[i] = to_f32<T>(k[(int64_t)s * ldk + base + c]) + ee;
          sc += qv[i] * kk[i];
          da += gv[i] * (to_f32<T>(v[(int64_t)s * ldv + base + c]) + ee);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o), da += __shfl_xor_sync(0xffffffffu, da, o);
      const float a = __expf(sc * scale - l);
      const float g = scale * a * (da - dsum);
#pragma unroll
      for (int i = 0; i < kMaxV; ++i) {
        const int c = lane + 32 * i;
        if (c < ch) {
          acc[i] += g * kk[i];
          if (de) de[(int64_t)ei * ldde + base + c] = from_f32<T>(g * qv[i] + a * gv[i]);
        }
      }
      if (lane == 0) alpha[(int64_t)ei * heads + h] = a, ds[(int64_t)ei * heads + h] = g;
    }
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < ch) dq[d * lddq + base + c] = from_f32<T>(acc[i]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gt_attention_bwd_src_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ dout, int64_t lddo,
                                                                   const float* __restrict__ alpha, const float* __restrict__ ds,
                                                                   const int32_t* __restrict__ rev_ptr, const int32_t* __restrict__ rev_eid,
                                                                   const int32_t* __restrict__ dst, T* __restrict__ dk, int64_t lddk,
                                                                   T* __restrict__ dv, int64_t lddv, int64_t n_src, int heads, int ch) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t items = n_src * heads;
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < items; w += warps_total) {
    const int64_t s = w / heads;
    const int h = (int)(w - s * heads);
    const int base = h * ch;
    float ak[kMaxV], av[kMaxV];
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) ak[i] = av[i] = 0.f;
    for (int j = rev_ptr[s]; j < rev_ptr[s + 1]; ++j) {
      const int ei = rev_eid[j];
      const int64_t d = dst[ei];
      const float a = alpha[(int64_t)ei * heads + h], g = ds[(int64_t)ei * heads + h];
#pragma unroll
      for (int i = 0; i < kMaxV; ++i) {
        const int c = lane + 32 * i;
        if (c < ch) {
          ak[i] += g * to_f32<T>(q[d * ldq + base + c]);
          av[i] += a * to_f32<T>(dout[d * lddo + base + c]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < ch) dk[s * lddk + base + c] = from_f32<T>(ak[i]), dv[s * lddv + base + c] = from_f32<T>(av[i]);
    }
  }
}

// ---- fast forms (rows of C = NCH * 32 * (16 / sizeof(T)) channels exactly, heads of LPH * (16 / sizeof(T)) channels) ---------------------
// Same two passes, laid out like the forward kernel: ONE warp owns a node for ALL heads, lane l holds the 16-byte chunks l, l + 32, ... of a
// row, a head spans LPH consecutive lanes (scores finish with log2(LPH) shuffles).  Every row access is NCH fully coalesced 512-byte warp
// loads (the per-(node, head) kernels above read 64 bytes per warp access and re-walk the edge list once per head): cfg2 layer, bf16:
// dst pass 1.73 ms -> ~0.3 ms, src pass 1.2 ms -> ~0.2 ms.
template <typename T>
struct BwdChunk {
  static constexpr int EPC = 16 / (int)sizeof(T);
  __device__ static __forceinline__ void load(const T* p, float (&f)[EPC]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    if constexpr (sizeof(T) == 4) {
      f[0] = __uint_as_float(u.x), f[1] = __uint_as_float(u.y), f[2] = __uint_as_float(u.z), f[3] = __uint_as_float(u.w);
    } else {
      f[0] = __uint_as_float(u.x << 16), f[1] = __uint_as_float(u.x & 0xffff0000u), f[2] = __uint_as_float(u.y << 16), f[3] = __uint_as_float(u.y & 0xffff0000u);
      f[4] = __uint_as_float(u.z << 16), f[5] = __uint_as_float(u.z & 0xffff0000u), f[6] = __uint_as_float(u.w << 16), f[7] = __uint_as_float(u.w & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store(T* p, const float (&f)[EPC]) {
    if constexpr (sizeof(T) == 4)
      *reinterpret_cast<uint4*>(p) = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    else
      *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
};

template <typename T, int NCH, int LPH, bool HAS_E>
__global__ void __launch_bounds__(256) gt_attention_bwd_dst_fast_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk,
                                                                        const T* __restrict__ v, int64_t ldv, const T* __restrict__ e, int64_t lde,
                                                                        const T* __restrict__ out, int64_t ldo, const T* __restrict__ dout,
                                                                        int64_t lddo, const float* __restrict__ lse, const int32_t* __restrict__ src,
                                                                        const int32_t* __restrict__ colptr, T* __restrict__ dq, int64_t lddq,
                                                                        T* __restrict__ de, int64_t ldde, float* __restrict__ alpha,
                                                                        float* __restrict__ ds, int64_t n_dst, int heads, float scale) {
  using CT = BwdChunk<T>;
  constexpr int EPC = CT::EPC;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  int col[NCH], hd[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) col[j] = (j * 32 + lane) * EPC, hd[j] = (j * 32 + lane) / LPH;
  for (int64_t d = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < n_dst; d += warps_total) {
    const int e0 = __ldg(colptr + d), e1 = __ldg(colptr + d + 1);
    float qf[NCH][EPC], gf[NCH][EPC], acc[NCH][EPC], D[NCH], l[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float of[EPC];
      CT::load(q + d * ldq + col[j], qf[j]);
      CT::load(dout + d * lddo + col[j], gf[j]);
      CT::load(out + d * ldo + col[j], of);
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < EPC; ++i) t += gf[j][i] * of[i], acc[j][i] = 0.f;
#pragma unroll
      for (int o = 1; o < LPH; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      D[j] = t;
      l[j] = __ldg(lse + d * heads + hd[j]);
    }
    int s_next = e0 < e1 ? __ldg(src + e0) : 0;
    for (int ei = e0; ei < e1; ++ei) {
      const int64_t s = s_next;
      if (ei + 1 < e1) s_next = __ldg(src + ei + 1);
      float kk[NCH][EPC], vv[NCH][EPC];
#pragma unroll
      for (int j = 0; j < NCH; ++j) {  // all loads of the edge first: 3 * NCH independent 16-byte loads in flight per lane
        CT::load(k + s * ldk + col[j], kk[j]);
        CT::load(v + s * ldv + col[j], vv[j]);
      }
      if constexpr (HAS_E) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          float ef[EPC];
          CT::load(e + (int64_t)ei * lde + col[j], ef);
#pragma unroll
          for (int i = 0; i < EPC; ++i) kk[j][i] += ef[i], vv[j][i] += ef[i];
        }
      }
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float sc = 0.f, da = 0.f;
#pragma unroll
        for (int i = 0; i < EPC; ++i) sc += qf[j][i] * kk[j][i], da += gf[j][i] * vv[j][i];
#pragma unroll
        for (int o = 1; o < LPH; o <<= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o), da += __shfl_xor_sync(0xffffffffu, da, o);
        const float a = __expf(sc * scale - l[j]);
        const float g = scale * a * (da - D[j]);
        float dev[EPC];
#pragma unroll
        for (int i = 0; i < EPC; ++i) acc[j][i] += g * kk[j][i], dev[i] = g * qf[j][i] + a * gf[j][i];
        if constexpr (HAS_E) CT::store(de + (int64_t)ei * ldde + col[j], dev);
        if ((lane & (LPH - 1)) == 0) alpha[(int64_t)ei * heads + hd[j]] = a, ds[(int64_t)ei * heads + hd[j]] = g;
      }
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) CT::store(dq + d * lddq + col[j], acc[j]);
  }
}

template <typename T, int NCH, int LPH>
__global__ void __launch_bounds__(256) gt_attention_bwd_src_fast_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ dout, int64_t lddo,
                                                                        const float* __restrict__ alpha, const float* __restrict__ ds,
                                                                        const int32_t* __restrict__ rev_ptr, const int32_t* __restrict__ rev_eid,
                                                                        const int32_t* __restrict__ dst, T* __restrict__ dk, int64_t lddk,
                                                                        T* __restrict__ dv, int64_t lddv, int64_t n_src, int heads) {
  using CT = BwdChunk<T>;
  constexpr int EPC = CT::EPC;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  int col[NCH], hd[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) col[j] = (j * 32 + lane) * EPC, hd[j] = (j * 32 + lane) / LPH;
  for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < n_src; s += warps_total) {
    float ak[NCH][EPC], av[NCH][EPC];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int i = 0; i < EPC; ++i) ak[j][i] = av[j][i] = 0.f;
    const int j0 = __ldg(rev_ptr + s), j1 = __ldg(rev_ptr + s + 1);
    for (int jj = j0; jj < j1; ++jj) {
      const int ei = __ldg(rev_eid + jj);
      const int64_t d = __ldg(dst + ei);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float qf[EPC], gf[EPC];
        CT::load(q + d * ldq + col[j], qf);
        CT::load(dout + d * lddo + col[j], gf);
        const float a = __ldg(alpha + (int64_t)ei * heads + hd[j]), g = __ldg(ds + (int64_t)ei * heads + hd[j]);
#pragma unroll
        for (int i = 0; i < EPC; ++i) ak[j][i] += g * qf[i], av[j][i] += a * gf[i];
      }
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      CT::store(dk + s * lddk + col[j], ak[j]);
      CT::store(dv + s * lddv + col[j], av[j]);
    }
  }
}

template <typename T, int NCH, int LPH>
int launch_bwd_fast(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* e, int64_t lde, const void* out,
                    int64_t ldo, const void* dout, int64_t lddo, const float* lse, const int32_t* src32, const int32_t* colptr32,
                    const int32_t* dst32, const int32_t* rev_ptr32, const int32_t* rev_eid32, void* dq, int64_t lddq, void* dk, int64_t lddk,
                    void* dv, int64_t lddv, void* de, int64_t ldde, float* alpha, float* ds, int64_t n_src, int64_t n_dst, int heads, float scale,
                    cudaStream_t s) {
  const int64_t cap = (int64_t)num_sms() * 8;
  if (n_dst > 0) {
    int64_t blocks = (n_dst + 7) / 8;
    if (blocks > cap) blocks = cap;
    if (e)
      gt_attention_bwd_dst_fast_kernel<T, NCH, LPH, true><<<(unsigned)blocks, 256, 0, s>>>(
          (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, (const T*)e, lde, (const T*)out, ldo, (const T*)dout, lddo, lse, src32, colptr32, (T*)dq,
          lddq, (T*)de, ldde, alpha, ds, n_dst, heads, scale);
    else
      gt_attention_bwd_dst_fast_kernel<T, NCH, LPH, false><<<(unsigned)blocks, 256, 0, s>>>(
          (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, (const T*)e, lde, (const T*)out, ldo, (const T*)dout, lddo, lse, src32, colptr32, (T*)dq,
          lddq, (T*)de, ldde, alpha, ds, n_dst, heads, scale);
    const int rc = launch_status("gt_attention_bwd_dst_fast_kernel");
    if (rc) return rc;
  }
  if (n_src > 0) {
    int64_t blocks = (n_src + 7) / 8;
    if (blocks > cap) blocks = cap;
    gt_attention_bwd_src_fast_kernel<T, NCH, LPH><<<(unsigned)blocks, 256, 0, s>>>((const T*)q, ldq, (const T*)dout, lddo, alpha, ds, rev_ptr32, rev_eid32,
                                                                                   dst32, (T*)dk, lddk, (T*)dv, lddv, n_src, heads);
    return launch_status("gt_attention_bwd_src_fast_kernel");
  }
  return 0;
}

// LayerNorm backward over rows of C channels (C <= 1024, lane owns channels lane + 32 i); row r of `groups` per matrix row lives at
// x + (r / groups) * ldx + (r % groups) * C (qk_norm: one LayerNorm per head).  Cotangent g = dy[r] (nullable) + dz[idx[r]] (nullable).
// dgamma / dbeta: per-block partial sums [gridDim.x, 2, C] (summed by the caller: deterministic).  dres (nullable) receives g itself (the
// residual branch e' = LN(h) + e of the GraphConv tail).
constexpr int kLnV = 32;
template <typename T>
__global__ void __launch_bounds__(256) layer_norm_bwd_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                                                             const T* __restrict__ dy, int64_t lddy, const T* __restrict__ dz, int64_t lddz,
                                                             const int32_t* __restrict__ idx, T* __restrict__ dx, int64_t lddx,
                                                             T* __restrict__ dres, int64_t lddr, float* __restrict__ partial, int64_t rows,
                                                             int groups, int C, float eps) {
  extern __shared__ float red[];  // [warps][2][C]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = (C + 31) / 32;
  float dg[kLnV], db[kLnV];
#pragma unroll
  for (int i = 0; i < kLnV; ++i) dg[i] = db[i] = 0.f;
  const float invC = 1.0f / (float)C;
  for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < rows; r += (int64_t)gridDim.x * wpb) {
    const int64_t mr = r / groups;
    const int gr = (int)(r - mr * groups);
    const T* xr = x + mr * ldx + (int64_t)gr * C;
    const T* dyr = dy ? dy + mr * lddy + (int64_t)gr * C : nullptr;
    const T* dzr = dz ? dz + (int64_t)(idx ? idx[r] : r) * lddz + (int64_t)gr * C : nullptr;
    float xv[kLnV], gv[kLnV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnV; ++i) {
      const int c = lane + 32 * i;
      xv[i] = gv[i] = 0.f;
      if (i < nv && c < C) {
        xv[i] = to_f32<T>(xr[c]);
        gv[i] = (dyr ? to_f32<T>(dyr[c]) : 0.f) + (dzr ? to_f32<T>(dzr[c]) : 0.f);
        s += xv[i];
      }
    }
    const float mean = warp_sum(s) * invC;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kLnV; ++i) {
      const int c = lane + 32 * i;
      if (i < nv && c < C) {
        const float t = xv[i] - mean;
        var += t * t;
      }
    }
    const float rstd = rsqrtf(warp_sum(var) * invC + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnV; ++i) {
      const int c = lane + 32 * i;
      if (i < nv && c < C) {
        xv[i] = (xv[i] - mean) * rstd;  // x-hat
        const float dxh = gv[i] * (gamma ? gamma[c] : 1.f);
        m1 += dxh, m2 += dxh * xv[i];
        dg[i] += gv[i] * xv[i], db[i] += gv[i];
      }
    }
    m1 = warp_sum(m1) * invC, m2 = warp_sum(m2) * invC;
#pragma unroll
    for (int i = 0; i < kLnV; ++i) {
      const int c = lane + 32 * i;
      if (i < nv && c < C) {
        const float dxh = gv[i] * (gamma ? gamma[c] : 1.f);
        dx[mr * lddx + (int64_t)gr * C + c] = from_f32<T>(rstd * (dxh - m1 - xv[i] * m2));
        if (dres) dres[mr * lddr + (int64_t)gr * C + c] = from_f32<T>(gv[i]);
      }
    }
  }
  // block partials of dgamma / dbeta
#pragma unroll
  for (int i = 0; i < kLnV; ++i) {
    const int c = lane + 32 * i;
    if (i < nv && c < C) red[(wib * 2) * C + c] = dg[i], red[(wib * 2 + 1) * C + c] = db[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const int which = c / C, cc = c - which * C;
    float t = 0.f;
    for (int w2 = 0; w2 < wpb; ++w2) t += red[(w2 * 2 + which) * C + cc];
    partial[((int64_t)blockIdx.x * 2 + which) * C + cc] = t;
  }
}

// mode 0: y = gelu(x); mode 1 (and 2 here: the two-MUFU form only exists in the 16-byte kernel): y = dy * gelu'(x)   (exact erf GELU)
template <typename T>
__global__ void gelu_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy, T* __restrict__ y, int64_t ldy, int64_t M,
                            int64_t N, int mode) {
  const int64_t total = M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N, n = i - m * N;
    const float v = to_f32<T>(x[m * ldx + n]);
    float r;
    if (mode == 0) {
      r = gelu_erf(v);
    } else {
      const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * v * v);
      r = to_f32<T>(dy[m * lddy + n]) * (cdf + v * pdf);
    }
    y[m * ldy + n] = from_f32<T>(r);
  }
}

// 16 bytes per thread (rows of N % EPC == 0 elements, 16-byte aligned): the scalar kernel above spends its time on per-element index
// arithmetic and 2-byte accesses (0.30 ms for the 168 MB hidden tensor of a cfg2 MLP; this one is bound by the three streams).
template <typename T>
__global__ void __launch_bounds__(256) gelu_vec_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy, T* __restrict__ y,
                                                       int64_t ldy, int64_t M, int64_t N, int mode) {
  using CT = BwdChunk<T>;
  constexpr int EPC = CT::EPC;
  const int64_t q = N / EPC, total = M * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / q, n = (i - m * q) * EPC;
    float v[EPC], r[EPC];
    CT::load(x + m * ldx + n, v);
    if (mode == 0) {
      // the same one-MUFU form as the GEMM epilogues (|error| <= 5e-7): with erff the forward pass was issue-bound (ncu: issue slots 80 %
      // busy, DRAM 33 %, 107 us for a [40 962, 2048] bf16 tensor)
#pragma unroll
      for (int j = 0; j < EPC; ++j) r[j] = gelu_erf_fast(v[j]);
    } else if (mode == 2) {
      // two MUFU per element: Phi(-|x|) = exp2(P5(|x|)) as in gelu_erf_fast (|error| <= 1.3e-5 in Phi: the fit is weighted for gelu itself)
      // and phi(x) = exp2(-x^2 / (2 ln 2)) / sqrt(2 pi).  With erff + expf (~45 instructions per element) the backward pass was bound by
      // instruction issue, not by its three streams; meant for bf16 cotangents (ulp 4e-3), the exact form stays mode 1.
      constexpr float c[6] = {ANEMOI_GELU_P5};
      float g[EPC];
      CT::load(dy + m * lddy + n, g);
#pragma unroll
      for (int j = 0; j < EPC; ++j) {
        const float t = fminf(fabsf(v[j]), 10.f);
        float p = fmaf(c[5], t, c[4]);
        p = fmaf(p, t, c[3]), p = fmaf(p, t, c[2]), p = fmaf(p, t, c[1]), p = fmaf(p, t, c[0]);
        float ex, pd;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(p));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pd) : "f"(-0.72134752044448170368f * v[j] * v[j]));
        const float cdf = v[j] >= 0.f ? 1.0f - ex : ex;
        r[j] = g[j] * fmaf(v[j], 0.39894228040143267794f * pd, cdf);
      }
    } else {
      float g[EPC];
      CT::load(dy + m * lddy + n, g);
#pragma unroll
      for (int j = 0; j < EPC; ++j) {
        const float cdf = 0.5f * (1.0f + erff(v[j] * 0.70710678118654752440f));
        const float pdf = 0.39894228040143267794f * __expf(-0.5f * v[j] * v[j]);
        r[j] = g[j] * (cdf + v[j] * pdf);
      }
    }
    CT::store(y + m * ldy + n, r);
  }
}

// LayerNorm backward, one row per warp in the 16-byte chunk layout (C = NCH * 32 * EPC exactly, groups == 1): every access is a coalesced
// 512-byte warp load / store.  Same arithmetic and the same per-block dgamma / dbeta partials as layer_norm_bwd_kernel.
template <typename T, int NCH>
__global__ void __launch_bounds__(256) layer_norm_bwd_fast_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                                                                  const T* __restrict__ dy, int64_t lddy, const T* __restrict__ dz, int64_t lddz,
                                                                  const int32_t* __restrict__ idx, T* __restrict__ dx, int64_t lddx,
                                                                  T* __restrict__ dres, int64_t lddr, float* __restrict__ partial, int64_t rows,
                                                                  float eps) {
  using CT = BwdChunk<T>;
  constexpr int EPC = CT::EPC;
  constexpr int C = NCH * 32 * EPC;
  extern __shared__ float red[];  // [warps][2][C]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float dg[NCH][EPC], db[NCH][EPC], gm[NCH][EPC];
#pragma unroll
  for (int j = 0; j < NCH; ++j)
#pragma unroll
    for (int i = 0; i < EPC; ++i) dg[j][i] = db[j][i] = 0.f, gm[j][i] = gamma ? __ldg(gamma + (j * 32 + lane) * EPC + i) : 1.f;
  const float invC = 1.0f / (float)C;
  for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < rows; r += (int64_t)gridDim.x * wpb) {
    const int64_t zr = dz ? (int64_t)(idx ? __ldg(idx + r) : r) : 0;
    float xv[NCH][EPC], gv[NCH][EPC];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int c = (j * 32 + lane) * EPC;
      CT::load(x + r * ldx + c, xv[j]);
      if (dy) {
        CT::load(dy + r * lddy + c, gv[j]);
      } else {
#pragma unroll
        for (int i = 0; i < EPC; ++i) gv[j][i] = 0.f;
      }
      if (dz) {
        float t[EPC];
        CT::load(dz + zr * lddz + c, t);
#pragma unroll
        for (int i = 0; i < EPC; ++i) gv[j][i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < EPC; ++i) s += xv[j][i];
    }
    const float mean = warp_sum(s) * invC;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int i = 0; i < EPC; ++i) {
        const float t = xv[j][i] - mean;
        var += t * t;
      }
    const float rstd = rsqrtf(warp_sum(var) * invC + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int i = 0; i < EPC; ++i) {
        xv[j][i] = (xv[j][i] - mean) * rstd;
        const float dxh = gv[j][i] * gm[j][i];
        m1 += dxh, m2 += dxh * xv[j][i];
        dg[j][i] += gv[j][i] * xv[j][i], db[j][i] += gv[j][i];
      }
    m1 = warp_sum(m1) * invC, m2 = warp_sum(m2) * invC;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int c = (j * 32 + lane) * EPC;
      float o[EPC];
#pragma unroll
      for (int i = 0; i < EPC; ++i) o[i] = rstd * (gv[j][i] * gm[j][i] - m1 - xv[j][i] * m2);
      CT::store(dx + r * lddx + c, o);
      if (dres) CT::store(dres + r * lddr + c, gv[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < NCH; ++j)
#pragma unroll
    for (int i = 0; i < EPC; ++i) {
      const int c = (j * 32 + lane) * EPC + i;
      red[(wib * 2) * C + c] = dg[j][i], red[(wib * 2 + 1) * C + c] = db[j][i];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const int which = c / C, cc = c - which * C;
    float t = 0.f;
    for (int w2 = 0; w2 < wpb; ++w2) t += red[(w2 * 2 + which) * C + cc];
    partial[((int64_t)blockIdx.x * 2 + which) * C + cc] = t;
  }
}

template <typename T, int NCH>
int launch_ln_bwd_fast(const void* x, int64_t ldx, const float* gamma, const void* dy, int64_t lddy, const void* dz, int64_t lddz, const int32_t* idx,
                       void* dx, int64_t lddx, void* dres, int64_t lddr, float* partial, int64_t blocks, int64_t rows, float eps, cudaStream_t s) {
  constexpr int C = NCH * 32 * (16 / (int)sizeof(T));
  const size_t smem = (size_t)8 * 2 * C * sizeof(float);
  static bool attr_dev[kMaxDevices] = {};
  if (smem > 48 * 1024 && !attr_dev[current_device()]) {
    ANEMOI_CUDA(cudaFuncSetAttribute(layer_norm_bwd_fast_kernel<T, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev[current_device()] = true;
  }
  layer_norm_bwd_fast_kernel<T, NCH><<<(unsigned)blocks, 256, smem, s>>>((const T*)x, ldx, gamma, (const T*)dy, lddy, (const T*)dz, lddz, idx, (T*)dx, lddx,
                                                                         (T*)dres, lddr, partial, rows, eps);
  return launch_status("layer_norm_bwd_fast_kernel");
}

// out[n] = sum over j in [ptr[n], ptr[n+1]) of rows[eid ? eid[j] : j]  — the backward of a row gather table[idx] when the gathering index
// list is sorted (ptr = CSR offsets of the dst-sorted edges, eid = null) or comes with its reverse CSR (ptr / eid = edges grouped by source).
// Deterministic (fixed order, fp32 accumulation), one warp per output row, 16-byte chunks; replaces PyTorch's index_add_ (atomics:
// 3.6 ms per [327 600, 1024] bf16 cotangent on a B200, against ~0.2 ms for the read).
template <typename T>
__global__ void __launch_bounds__(256) segment_sum_kernel(const T* __restrict__ rows, int64_t ld, const int32_t* __restrict__ ptr,
                                                          const int32_t* __restrict__ eid, T* __restrict__ out, int64_t ldo, int64_t n_out, int C) {
  using CT = BwdChunk<T>;
  constexpr int EPC = CT::EPC;
  const int lane = threadIdx.x & 31;
  const int chunks = C / EPC;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < n_out; n += warps_total) {
    const int j0 = __ldg(ptr + n), j1 = __ldg(ptr + n + 1);
    for (int c = lane; c < chunks; c += 32) {
      float acc[EPC];
#pragma unroll
      for (int i = 0; i < EPC; ++i) acc[i] = 0.f;
      for (int j = j0; j < j1; ++j) {
        const int64_t r = eid ? (int64_t)__ldg(eid + j) : (int64_t)j;
        float f[EPC];
        CT::load(rows + r * ld + c * EPC, f);
#pragma unroll
        for (int i = 0; i < EPC; ++i) acc[i] += f[i];
      }
      CT::store(out + n * ldo + c * EPC, acc);
    }
  }
}


// out[c] = sum over rows of x[r, c] (fp32): the bias gradient of every Linear (db = column sums of the cotangent; PyTorch's sum(0) ran at
// ~1.2 TB/s on these [N, C] / [E, C] bf16 cotangents: 6.2 of the 47 ms of a cfg2 training step).  Deterministic two-stage reduction: stage 1
// writes one fp32 partial row per row chunk (grid = column tiles x row chunks, a thread keeps one 16-byte column group and walks its warp's
// rows with eight loads in flight), stage 2 is the same kernel over the partial rows.  No atomics; the order is fixed by the shape.
template <typename T>
__global__ void __launch_bounds__(256) col_sum_kernel(const T* __restrict__ x, int64_t ld, float* __restrict__ out, int64_t ldo, int64_t M,
                                                      int64_t N, int64_t rows_per_chunk) {
  constexpr int VEC = 16 / (int)sizeof(T);
  __shared__ float s_part[8][32 * VEC];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c0 = ((int64_t)blockIdx.x * 32 + lane) * VEC;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  auto add16 = [&](const uint4& u) {
    if constexpr (sizeof(T) == 2) {
      acc[0] += __uint_as_float(u.x << 16), acc[1] += __uint_as_float(u.x & 0xffff0000u), acc[2] += __uint_as_float(u.y << 16);
      acc[3] += __uint_as_float(u.y & 0xffff0000u), acc[4] += __uint_as_float(u.z << 16), acc[5] += __uint_as_float(u.z & 0xffff0000u);
      acc[6] += __uint_as_float(u.w << 16), acc[7] += __uint_as_float(u.w & 0xffff0000u);
    } else {
      acc[0] += __uint_as_float(u.x), acc[1] += __uint_as_float(u.y), acc[2] += __uint_as_float(u.z), acc[3] += __uint_as_float(u.w);
    }
  };
  if (c0 + VEC <= N) {
    const T* col = x + c0;
    int64_t r = r0 + w;
    for (; r + 56 < r1; r += 64) {  // eight 16-byte loads in flight per thread
      uint4 u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(col + (r + 8 * i) * ld));
#pragma unroll
      for (int i = 0; i < 8; ++i) add16(u[i]);
    }
    for (; r < r1; r += 8) add16(__ldg(reinterpret_cast<const uint4*>(col + r * ld)));
  } else if (c0 < N) {  // ragged last column group of the matrix
    for (int64_t r = r0 + w; r < r1; r += 8) {
#pragma unroll
      for (int i = 0; i < VEC; ++i)
        if (c0 + i < N) acc[i] += (float)x[r * ld + c0 + i];
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) s_part[w][lane * VEC + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 32 * VEC) {
    const int64_t c = (int64_t)blockIdx.x * 32 * VEC + threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += s_part[i][threadIdx.x];
    if (c < N) out[(int64_t)blockIdx.y * ldo + c] = s;
  }
}

}  // namespace
}  // namespace anemoi

using namespace anemoi;

extern "C" int anemoi_b200_gt_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* e,
                                            int64_t lde, const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                            const int32_t* src32, const int32_t* colptr32, const int32_t* dst32, const int32_t* rev_ptr32,
                                            const int32_t* rev_eid32, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                            void* de, int64_t ldde, float* alpha_scratch, float* ds_scratch, int64_t n_src, int64_t n_dst,
                                            int64_t heads, int64_t ch, int dtype, void* stream) {
  ANEMOI_CHECK_ARG(n_src >= 0 && n_dst >= 0 && heads >= 1 && ch >= 1 && ch <= 32 * kMaxV, "gt_attention_bwd: bad shape (channels per head <= 256)");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "gt_attention_bwd: bad dtype %d", dtype);
  ANEMOI_CHECK_ARG(q && k && v && out && dout && lse && colptr32 && rev_ptr32 && dq && dk && dv, "gt_attention_bwd: null pointer");
  ANEMOI_CHECK_ARG((e == nullptr) == (de == nullptr), "gt_attention_bwd: e and de go together");
  cudaStream_t s = (cudaStream_t)stream;
  const float scale = 1.0f / sqrtf((float)ch);
  {
    // fast forms: 16-byte aligned rows of exactly NCH * 32 chunks, heads of LPH chunks (cfg2: bf16, C = 512, Ch = 32 -> NCH 2, LPH 4)
    const int es = dtype == ANEMOI_BF16 ? 2 : 4, epc = 16 / es;
    const int64_t C = heads * ch;
    auto al = [&](const void* p, int64_t ld) { return !p || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * es) % 16 == 0); };
    const bool ok = ch % epc == 0 && C % (32 * epc) == 0 && al(q, ldq) && al(k, ldk) && al(v, ldv) && al(e, lde) && al(out, ldo) && al(dout, lddo) &&
                    al(dq, lddq) && al(dk, lddk) && al(dv, lddv) && al(de, ldde) && dst32 && rev_eid32;
    const int nch = ok ? (int)(C / (32 * epc)) : 0, lph = ok ? (int)(ch / epc) : 0;
#define ANEMOI_BWD_FAST(T, NCH, LPH)                                                                                                                  \
  return launch_bwd_fast<T, NCH, LPH>(q, ldq, k, ldk, v, ldv, e, lde, out, ldo, dout, lddo, lse, src32, colptr32, dst32, rev_ptr32, rev_eid32, dq, lddq, \
                                      dk, lddk, dv, lddv, de, ldde, alpha_scratch, ds_scratch, n_src, n_dst, (int)heads, scale, s)
    if (ok && dtype == ANEMOI_BF16) {
      if (nch == 1 && lph == 4) ANEMOI_BWD_FAST(__nv_bfloat16, 1, 4);
      if (nch == 2 && lph == 4) ANEMOI_BWD_FAST(__nv_bfloat16, 2, 4);
      if (nch == 4 && lph == 4) ANEMOI_BWD_FAST(__nv_bfloat16, 4, 4);
      if (nch == 1 && lph == 8) ANEMOI_BWD_FAST(__nv_bfloat16, 1, 8);
      if (nch == 2 && lph == 8) ANEMOI_BWD_FAST(__nv_bfloat16, 2, 8);
      if (nch == 4 && lph == 8) ANEMOI_BWD_FAST(__nv_bfloat16, 4, 8);
      if (nch == 1 && lph == 2) ANEMOI_BWD_FAST(__nv_bfloat16, 1, 2);
    } else if (ok) {
      if (nch == 1 && lph == 4) ANEMOI_BWD_FAST(float, 1, 4);
      if (nch == 2 && lph == 4) ANEMOI_BWD_FAST(float, 2, 4);
      if (nch == 4 && lph == 8) ANEMOI_BWD_FAST(float, 4, 8);
      if (nch == 2 && lph == 8) ANEMOI_BWD_FAST(float, 2, 8);
      if (nch == 4 && lph == 16) ANEMOI_BWD_FAST(float, 4, 16);
      if (nch == 8 && lph == 8) ANEMOI_BWD_FAST(float, 8, 8);
    }
#undef ANEMOI_BWD_FAST
  }
  const int64_t cap = (int64_t)num_sms() * 8;
  if (n_dst > 0) {
    int64_t blocks = (n_dst * heads + 7) / 8;
    if (blocks > cap) blocks = cap;
    if (dtype == ANEMOI_BF16)
      gt_attention_bwd_dst_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>(
          (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, (const __nv_bfloat16*)e, lde, (const __nv_bfloat16*)out,
          ldo, (const __nv_bfloat16*)dout, lddo, lse, src32, colptr32, (__nv_bfloat16*)dq, lddq, (__nv_bfloat16*)de, ldde, alpha_scratch, ds_scratch,
          n_dst, (int)heads, (int)ch, scale);
    else
      gt_attention_bwd_dst_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)q, ldq, (const float*)k, ldk, (const float*)v, ldv, (const float*)e,
                                                                          lde, (const float*)out, ldo, (const float*)dout, lddo, lse, src32, colptr32,
                                                                          (float*)dq, lddq, (float*)de, ldde, alpha_scratch, ds_scratch, n_dst,
                                                                          (int)heads, (int)ch, scale);
    const int rc = launch_status("gt_attention_bwd_dst_kernel");
    if (rc) return rc;
  }
  if (n_src > 0) {
    int64_t blocks = (n_src * heads + 7) / 8;
    if (blocks > cap) blocks = cap;
    if (dtype == ANEMOI_BF16)
      gt_attention_bwd_src_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)dout, lddo,
                                                                                  alpha_scratch, ds_scratch, rev_ptr32, rev_eid32, dst32,
                                                                                  (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, n_src, (int)heads,
                                                                                  (int)ch);
    else
      gt_attention_bwd_src_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)q, ldq, (const float*)dout, lddo, alpha_scratch, ds_scratch,
                                                                          rev_ptr32, rev_eid32, dst32, (float*)dk, lddk, (float*)dv, lddv, n_src,
                                                                          (int)heads, (int)ch);
    return launch_status("gt_attention_bwd_src_kernel");
  }
  return 0;
}

extern "C" int anemoi_b200_layer_norm_bwd(const void* x, int64_t ldx, const float* gamma, const void* dy, int64_t lddy, const void* dz, int64_t lddz,
                                          const int32_t* idx, void* dx, int64_t lddx, void* dres, int64_t lddr, float* partial, int64_t n_partial,
                                          int64_t M, int64_t groups, int64_t C, float eps, int dtype, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && groups >= 1 && C >= 1 && C <= 32 * kLnV, "layer_norm_bwd: bad shape (C <= 1024)");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "layer_norm_bwd: bad dtype %d", dtype);
  ANEMOI_CHECK_ARG(x && dx && partial && (dy || dz) && n_partial >= 1, "layer_norm_bwd: null pointer");
  const int64_t rows = M * groups;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > n_partial) blocks = n_partial;
  if (blocks < 1) blocks = 1;
  ANEMOI_CUDA(cudaMemsetAsync(partial, 0, (size_t)n_partial * 2 * C * sizeof(float), (cudaStream_t)stream));
  {
    const int es = dtype == ANEMOI_BF16 ? 2 : 4, epc = 16 / es;
    auto al = [&](const void* p, int64_t ld) { return !p || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * es) % 16 == 0); };
    if (groups == 1 && C % (32 * epc) == 0 && al(x, ldx) && al(dy, lddy) && al(dz, lddz) && al(dx, lddx) && al(dres, lddr)) {
      const int nch = (int)(C / (32 * epc));
      cudaStream_t st = (cudaStream_t)stream;
#define ANEMOI_LNB(T, NCH) return launch_ln_bwd_fast<T, NCH>(x, ldx, gamma, dy, lddy, dz, lddz, idx, dx, lddx, dres, lddr, partial, blocks, rows, eps, st)
      if (dtype == ANEMOI_BF16) {
        if (nch == 1) ANEMOI_LNB(__nv_bfloat16, 1);
        if (nch == 2) ANEMOI_LNB(__nv_bfloat16, 2);
        if (nch == 4) ANEMOI_LNB(__nv_bfloat16, 4);
      } else {
        if (nch == 1) ANEMOI_LNB(float, 1);
        if (nch == 2) ANEMOI_LNB(float, 2);
        if (nch == 4) ANEMOI_LNB(float, 4);
        if (nch == 8) ANEMOI_LNB(float, 8);
      }
#undef ANEMOI_LNB
    }
  }
  const size_t smem = (size_t)8 * 2 * C * sizeof(float);
  if (dtype == ANEMOI_BF16) {
    static bool attr_dev[kMaxDevices] = {};
    if (smem > 48 * 1024 && !attr_dev[current_device()]) {
      ANEMOI_CUDA(cudaFuncSetAttribute(layer_norm_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4));
      attr_dev[current_device()] = true;
    }
    layer_norm_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, ldx, gamma, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)dz, lddz, idx, (__nv_bfloat16*)dx, lddx,
        (__nv_bfloat16*)dres, lddr, partial, rows, (int)groups, (int)C, eps);
  } else {
    static bool attr_dev[kMaxDevices] = {};
    if (smem > 48 * 1024 && !attr_dev[current_device()]) {
      ANEMOI_CUDA(cudaFuncSetAttribute(layer_norm_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4));
      attr_dev[current_device()] = true;
    }
    layer_norm_bwd_kernel<float><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>((const float*)x, ldx, gamma, (const float*)dy, lddy,
                                                                                         (const float*)dz, lddz, idx, (float*)dx, lddx, (float*)dres,
                                                                                         lddr, partial, rows, (int)groups, (int)C, eps);
  }
  return launch_status("layer_norm_bwd_kernel");
}

extern "C" int anemoi_b200_gelu(const void* x, int64_t ldx, const void* dy, int64_t lddy, void* y, int64_t ldy, int64_t M, int64_t N, int mode,
                                int dtype, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && N >= 0 && mode >= 0 && mode <= 2, "gelu: bad argument");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "gelu: bad dtype %d", dtype);
  if (M * N == 0) return 0;
  ANEMOI_CHECK_ARG(x && y && (mode == 0 || dy), "gelu: null pointer");
  {
    const int es = dtype == ANEMOI_BF16 ? 2 : 4, epc = 16 / es;
    auto al = [&](const void* p, int64_t ld) { return !p || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * es) % 16 == 0); };
    if (N % epc == 0 && al(x, ldx) && al(dy, lddy) && al(y, ldy)) {
      int64_t vb = (M * (N / epc) + 255) / 256;
      const int64_t vcap = (int64_t)num_sms() * 32;
      if (vb > vcap) vb = vcap;
      if (dtype == ANEMOI_BF16)
        gelu_vec_kernel<__nv_bfloat16><<<(unsigned)vb, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)dy, lddy,
                                                                                        (__nv_bfloat16*)y, ldy, M, N, mode);
      else
        gelu_vec_kernel<float><<<(unsigned)vb, 256, 0, (cudaStream_t)stream>>>((const float*)x, ldx, (const float*)dy, lddy, (float*)y, ldy, M, N, mode);
      return launch_status("gelu_vec_kernel");
    }
  }
  int64_t blocks = (M * N + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == ANEMOI_BF16)
    gelu_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)dy, lddy,
                                                                                    (__nv_bfloat16*)y, ldy, M, N, mode);
  else
    gelu_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, ldx, (const float*)dy, lddy, (float*)y, ldy, M, N, mode);
  return launch_status("gelu_kernel");
}

extern "C" int anemoi_b200_segment_sum(const void* rows, int64_t ld, const int32_t* ptr32, const int32_t* eid32, void* out, int64_t ldo, int64_t n_out,
                                       int64_t C, int dtype, void* stream) {
  ANEMOI_CHECK_ARG(n_out >= 0 && C >= 1 && ld >= C && ldo >= C, "segment_sum: bad shape");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "segment_sum: bad dtype %d", dtype);
  if (n_out == 0) return 0;
  ANEMOI_CHECK_ARG(ptr32 && out, "segment_sum: null pointer");
  const int es = dtype == ANEMOI_BF16 ? 2 : 4;
  ANEMOI_CHECK_ARG(C % (16 / es) == 0 && (!rows || (reinterpret_cast<uintptr_t>(rows) & 15) == 0) && (ld * es) % 16 == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ldo * es) % 16 == 0,
                   "segment_sum: rows of whole 16-byte chunks, 16-byte aligned");
  int64_t blocks = (n_out + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == ANEMOI_BF16)
    segment_sum_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)rows, ld, ptr32, eid32, (__nv_bfloat16*)out,
                                                                                          ldo, n_out, (int)C);
  else
    segment_sum_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)rows, ld, ptr32, eid32, (float*)out, ldo, n_out, (int)C);
  return launch_status("segment_sum_kernel");
}

extern "C" int anemoi_b200_col_sum(const void* x, int64_t ld, float* out, float* partial, int64_t n_partial, int64_t M, int64_t N, int dtype,
                                   void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && N >= 0 && ld >= N && n_partial >= 0, "col_sum: bad shape");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "col_sum: bad dtype %d", dtype);
  if (N == 0) return 0;
  ANEMOI_CHECK_ARG(out, "col_sum: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (M == 0) {
    ANEMOI_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), s));
    return 0;
  }
  const int es = dtype == ANEMOI_BF16 ? 2 : 4, vec = 16 / es;
  ANEMOI_CHECK_ARG(x && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ld * es) % 16 == 0, "col_sum: 16-byte aligned rows");
  const int64_t col_tiles = (N + 32 * vec - 1) / (32 * vec);
  ANEMOI_CHECK_ARG(col_tiles <= 0x7fffffff, "col_sum: too many columns");
  // row chunks: ~8 CTAs per SM in total, at least 32 rows each, no more than the caller's partial rows (and the grid's y limit)
  int64_t chunks = ((int64_t)num_sms() * 8 + col_tiles - 1) / col_tiles;
  chunks = std::min(chunks, std::min((M + 31) / 32, std::min<int64_t>(partial ? n_partial : 1, 65535)));
  chunks = std::max<int64_t>(chunks, 1);
  const int64_t rpc = (M + chunks - 1) / chunks;
  chunks = (M + rpc - 1) / rpc;
  const int64_t ldp = (N + 3) / 4 * 4;  // pitch of the partial rows: whole 16-byte groups, so that stage 2 reads them like any fp32 matrix
  ANEMOI_CHECK_ARG(chunks == 1 || (reinterpret_cast<uintptr_t>(partial) & 15) == 0, "col_sum: partial must be 16-byte aligned");
  float* stage1 = chunks == 1 ? out : partial;
  const dim3 grid((unsigned)col_tiles, (unsigned)chunks);
  if (dtype == ANEMOI_BF16)
    col_sum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, ld, stage1, ldp, M, N, rpc);
  else
    col_sum_kernel<float><<<grid, 256, 0, s>>>((const float*)x, ld, stage1, ldp, M, N, rpc);
  if (chunks > 1) col_sum_kernel<float><<<dim3((unsigned)((N + 127) / 128), 1u), 256, 0, s>>>(partial, ldp, out, ldp, chunks, N, chunks);
  return launch_status("col_sum_kernel");
}
