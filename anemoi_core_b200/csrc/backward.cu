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

// mode 0: y = gelu(x); mode 1: y = dy * gelu'(x)   (exact erf GELU)
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
  ANEMOI_CHECK_ARG(M >= 0 && N >= 0 && (mode == 0 || mode == 1), "gelu: bad argument");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "gelu: bad dtype %d", dtype);
  if (M * N == 0) return 0;
  ANEMOI_CHECK_ARG(x && y && (mode == 0 || dy), "gelu: null pointer");
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
