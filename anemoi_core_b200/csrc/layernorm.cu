// layernorm.cu — row LayerNorm (+ optional residual), one warp per (row, group).  HBM-bound: one read of x
// (+ residual), one write of y; 16-byte vector accesses on the fast path; fp32 statistics (two-pass in registers).
#include "common.cuh"
#include "gemm.h"

namespace anemoi {

// Fast path: C % 8 == 0, C <= 256*ITERS, all row starts 16-byte aligned.  Lane owns 8 contiguous elements per iteration.
template <typename TI, typename TR, typename TO, int ITERS>
__global__ void __launch_bounds__(256) layer_norm_vec_kernel(const TI* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const TR* __restrict__ res, int64_t ldr,
                                                             TO* __restrict__ y, int64_t ldy, int64_t M, int64_t groups, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < M * groups; w += warps_total) {
    const int64_t m = w / groups, g = w - m * groups;
    const TI* xr = x + m * ldx + g * C;
    float v[ITERS][8];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * 256 + lane * 8;
      if (c < C) {
        load_vec_f32<TI, 8>(xr + c, v[it]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[it][i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[it][i] = 0.f;
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * 256 + lane * 8;
      if (c < C) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[it][i] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * 256 + lane * 8;
      if (c < C) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[it][i] - mean) * rstd;
        if (gamma) {
          float gm[8];
          load_vec_f32<float, 8>(gamma + c, gm);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] *= gm[i];
        }
        if (beta) {
          float bt[8];
          load_vec_f32<float, 8>(beta + c, bt);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += bt[i];
        }
        if (res) {
          float r[8];
          load_vec_f32<TR, 8>(res + m * ldr + g * C + c, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += r[i];
        }
        store_vec_f32<TO, 8>(y + m * ldy + g * C + c, o);
      }
    }
  }
}

// Row statistics only (mean, rstd): the LayerNorm itself is folded into the consuming GEMM (gemm_tcgen05.cu).  One warp per row,
// 16-byte loads, two-pass variance in registers; reads x once, writes 8 bytes per row.
template <typename TI, int ITERS>
__global__ void __launch_bounds__(256) row_stats_kernel(const TI* __restrict__ x, int64_t ldx, float2* __restrict__ stats, int64_t M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps_total) {
    const TI* xr = x + m * ldx;
    float v[ITERS][8];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * 256 + lane * 8;
      if (c < C) {
        load_vec_f32<TI, 8>(xr + c, v[it]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[it][i];
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * 256 + lane * 8;
      if (c < C) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[it][i] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0) stats[m] = make_float2(mean, rstd);
  }
}

// Streaming variant (bf16 rows of a multiple of 64 elements): 8 lanes per row, 4 rows per warp, every lane streams its 16-byte chunks with
// four loads in flight and accumulates sums SHIFTED by the row's first element (one pass, no E[x^2] - mean^2 cancellation: the shifted
// values are O(std)); 3 shuffles per reduction instead of 5 and no register-resident copy of the row.  The per-warp-row kernel above ran at
// 2.6 TB/s on the [40 962, 512] residual stream (16 us x 39 launches = 7 % of a cfg2 step); this one is bound by the read.
__global__ void __launch_bounds__(256) row_stats_stream_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, float2* __restrict__ stats, int64_t M,
                                                               int C, float eps, int reverse) {
  pdl_wait();  // PDL (common.cuh): x comes from the GEMM just before
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, sub = lane & 7, rw = lane >> 3;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int chunks = C >> 3;  // 16-byte chunks per row; lane `sub` takes chunks sub, sub + 8, ...
  const int64_t n_groups = (M + 3) / 4;
  for (int64_t g0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g0 < n_groups; g0 += groups_total) {
    const int64_t g = reverse ? n_groups - 1 - g0 : g0;  // reverse: from the rows the producer wrote last (still in L2) to the first
    const int64_t m = g * 4 + rw;
    const bool ok = m < M;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (ok ? m : M - 1) * ldx);
    const float x0 = __uint_as_float(__ldg(reinterpret_cast<const uint32_t*>(xr)) << 16);  // first element of the row: the shift
    float2 s2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
    const float2 sh = make_float2(-x0, -x0);
#pragma unroll 4
    for (int c = sub; c < chunks; c += 8) {
      const uint4 u = __ldg(xr + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 v = __fadd2_rn(make_float2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u)), sh);
        s2 = __fadd2_rn(s2, v);
        q2 = __ffma2_rn(v, v, q2);
      }
    }
    float s = s2.x + s2.y, q = q2.x + q2.y;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o), q += __shfl_xor_sync(0xffffffffu, q, o);
    const float invC = 1.0f / (float)C, ms = s * invC;
    if (ok && sub == 0) stats[m] = make_float2(x0 + ms, rsqrtf(fmaxf(q * invC - ms * ms, 0.f) + eps));
  }
}

// Partial row statistics of a stored matrix, the layout a GEMM epilogue writes into EpiParams::stats_out: per row and 64-column block
// (mean, M2).  Fallback for GEMM paths without the fused version; one warp per row, two columns per lane and block.
__global__ void __launch_bounds__(256) partial_row_stats_kernel(const void* __restrict__ x, int64_t ldx, int dtype, int64_t M, int64_t N, int parts,
                                                                float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps_total) {
    for (int b = 0; b < parts; ++b) {
      float v[2], s = 0.f, q = 0.f;
      const float nb = (float)min((int64_t)kStatsBlock, N - (int64_t)b * kStatsBlock);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t c = (int64_t)b * kStatsBlock + lane * 2 + j;
        v[j] = c < N ? load_as_f32(x, m * ldx + c, dtype) : 0.f;
        s += v[j];
      }
      const float mean = warp_sum(s) / nb;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t c = (int64_t)b * kStatsBlock + lane * 2 + j;
        if (c < N) q += (v[j] - mean) * (v[j] - mean);
      }
      q = warp_sum(q);
      if (lane == 0) stats[m * parts + b] = make_float2(mean, q);  // (block mean, block M2): two-pass, like the GEMM epilogue's shifted sums
    }
  }
}

int launch_partial_row_stats(const void* out, int64_t ldo, int o_dtype, int64_t M, int64_t N, float* stats_out, cudaStream_t s) {
  if (M == 0) return 0;
  const int parts = (int)((N + kStatsBlock - 1) / kStatsBlock);
  int64_t blocks = (M + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  partial_row_stats_kernel<<<(unsigned)blocks, 256, 0, s>>>(out, ldo, o_dtype, M, N, parts, reinterpret_cast<float2*>(stats_out));
  return launch_status("partial_row_stats_kernel");
}

// Conditional LayerNorm (layers/normalization.py:34-94): y = LN(x) * (1 + cond . Ws^T + bs) + (cond . Wb^T + bb), no learnable LayerNorm
// affine.  One warp per row: two-pass statistics, then per channel the two Dc-long dot products against the conditioning row (held one
// element per lane and broadcast by shuffles; Dc <= 32); the 2 x [C, Dc] weights stay in L1.  The per-row scale / bias tensors
// ([M, 2C] fp32 in the reference) are never materialised.
__global__ void __launch_bounds__(256) cond_layer_norm_kernel(const void* __restrict__ x, int64_t ldx, int x_dtype, const float* __restrict__ cond,
                                                               int64_t ldc, const float* __restrict__ ws, const float* __restrict__ bs,
                                                               const float* __restrict__ wb, const float* __restrict__ bb, void* __restrict__ y,
                                                               int64_t ldy, int y_dtype, int64_t M, int C, int Dc, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps_total) {
    const int64_t xo = m * ldx;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += load_as_f32(x, xo + c, x_dtype);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = load_as_f32(x, xo + c, x_dtype) - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    const float cd = lane < Dc ? cond[m * ldc + lane] : 0.f;
    for (int c0 = 0; c0 < C; c0 += 32) {  // uniform trip count: the shuffles below need the whole warp
      const int c = c0 + lane;
      const bool ok = c < C;
      float sc = ok ? bs[c] : 0.f, bi = ok ? bb[c] : 0.f;
      for (int d = 0; d < Dc; ++d) {
        const float cv = __shfl_sync(0xffffffffu, cd, d);
        if (ok) {
          sc = fmaf(cv, ws[(int64_t)c * Dc + d], sc);
          bi = fmaf(cv, wb[(int64_t)c * Dc + d], bi);
        }
      }
      if (ok) store_from_f32(y, m * ldy + c, y_dtype, (load_as_f32(x, xo + c, x_dtype) - mean) * rstd * (1.0f + sc) + bi);
    }
  }
}

// Generic path: any C / alignment.  Lane strides over the row; re-reads hit L1.
__global__ void __launch_bounds__(256) layer_norm_generic_kernel(const void* __restrict__ x, int64_t ldx, int x_dtype,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const void* __restrict__ res, int64_t ldr, int r_dtype, void* __restrict__ y,
                                                                 int64_t ldy, int y_dtype, int64_t M, int64_t groups, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < M * groups; w += warps_total) {
    const int64_t m = w / groups, g = w - m * groups;
    const int64_t xo = m * ldx + g * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += load_as_f32(x, xo + c, x_dtype);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = load_as_f32(x, xo + c, x_dtype) - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    for (int c = lane; c < C; c += 32) {
      float o = (load_as_f32(x, xo + c, x_dtype) - mean) * rstd;
      if (gamma) o *= gamma[c];
      if (beta) o += beta[c];
      if (res) o += load_as_f32(res, m * ldr + g * C + c, r_dtype);
      store_from_f32(y, m * ldy + g * C + c, y_dtype, o);
    }
  }
}

template <typename TI, typename TR, typename TO>
static int launch_vec(const void* x, int64_t ldx, const float* gamma, const float* beta, const void* res, int64_t ldr, void* y, int64_t ldy,
                      int64_t M, int64_t groups, int C, float eps, cudaStream_t s) {
  const int64_t rows = M * groups;
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
#define LN_LAUNCH(IT)                                                                                                             \
  layer_norm_vec_kernel<TI, TR, TO, IT><<<(unsigned)blocks, 256, 0, s>>>((const TI*)x, ldx, gamma, beta, (const TR*)res, ldr, (TO*)y, ldy, M, \
                                                                         groups, C, eps)
  if (C <= 256)
    LN_LAUNCH(1);
  else if (C <= 512)
    LN_LAUNCH(2);
  else if (C <= 1024)
    LN_LAUNCH(4);
  else
    LN_LAUNCH(8);
#undef LN_LAUNCH
  return launch_status("layer_norm_vec_kernel");
}

}  // namespace anemoi

using namespace anemoi;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int anemoi_b200_layer_norm(const void* x, int64_t ldx, int x_dtype, const float* gamma, const float* beta, const void* residual,
                                      int64_t ldr, int r_dtype, void* y, int64_t ldy, int y_dtype, int64_t M, int64_t groups, int64_t C,
                                      float eps, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && groups >= 1 && C >= 1 && C < (1 << 30), "layer_norm: bad shape");
  ANEMOI_CHECK_ARG(ldx >= groups * C && ldy >= groups * C, "layer_norm: leading dimension too small");
  ANEMOI_CHECK_ARG((x_dtype | 1) == 1 && (y_dtype | 1) == 1 && (r_dtype | 1) == 1, "layer_norm: bad dtype");
  if (M == 0) return 0;
  ANEMOI_CHECK_ARG(x && y, "layer_norm: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int xs = x_dtype == ANEMOI_BF16 ? 2 : 4, ys = y_dtype == ANEMOI_BF16 ? 2 : 4, rs = r_dtype == ANEMOI_BF16 ? 2 : 4;
  const bool vec = C % 8 == 0 && C <= 2048 && aligned16(x) && aligned16(y) && (ldx * xs) % 16 == 0 && (ldy * ys) % 16 == 0 &&
                   (C * xs) % 16 == 0 && (C * ys) % 16 == 0 && (!gamma || aligned16(gamma)) && (!beta || aligned16(beta)) &&
                   (!residual || (aligned16(residual) && (ldr * rs) % 16 == 0 && (C * rs) % 16 == 0));
  if (vec) {
    const int key = x_dtype * 4 + (residual ? r_dtype : x_dtype) * 2 + y_dtype;
    switch (key) {
#define LN_CASE(K, TI, TR, TO) \
  case K:                      \
    return launch_vec<TI, TR, TO>(x, ldx, gamma, beta, residual, ldr, y, ldy, M, groups, (int)C, eps, s)
      LN_CASE(0, float, float, float);
      LN_CASE(1, float, float, __nv_bfloat16);
      LN_CASE(2, float, __nv_bfloat16, float);
      LN_CASE(3, float, __nv_bfloat16, __nv_bfloat16);
      LN_CASE(4, __nv_bfloat16, float, float);
      LN_CASE(5, __nv_bfloat16, float, __nv_bfloat16);
      LN_CASE(6, __nv_bfloat16, __nv_bfloat16, float);
      LN_CASE(7, __nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
#undef LN_CASE
    }
  }
  const int64_t rows = M * groups;
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  layer_norm_generic_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, ldx, x_dtype, gamma, beta, residual, ldr, r_dtype, y, ldy, y_dtype, M, groups,
                                                             (int)C, eps);
  return launch_status("layer_norm_generic_kernel");
}

extern "C" int anemoi_b200_cond_layer_norm(const void* x, int64_t ldx, int x_dtype, const float* cond, int64_t ldc, const float* w_scale,
                                           const float* b_scale, const float* w_bias, const float* b_bias, void* y, int64_t ldy, int y_dtype,
                                           int64_t M, int64_t C, int64_t Dc, float eps, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && C >= 1 && C < (1 << 30) && Dc >= 1 && Dc <= 32, "cond_layer_norm: bad shape (condition width 1..32)");
  ANEMOI_CHECK_ARG(ldx >= C && ldy >= C && ldc >= Dc, "cond_layer_norm: leading dimension too small");
  ANEMOI_CHECK_ARG((x_dtype | 1) == 1 && (y_dtype | 1) == 1, "cond_layer_norm: bad dtype");
  if (M == 0) return 0;
  ANEMOI_CHECK_ARG(x && y && cond && w_scale && b_scale && w_bias && b_bias, "cond_layer_norm: null pointer");
  int64_t blocks = (M + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  cond_layer_norm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, x_dtype, cond, ldc, w_scale, b_scale, w_bias, b_bias, y, ldy,
                                                                            y_dtype, M, (int)C, (int)Dc, eps);
  return launch_status("cond_layer_norm_kernel");
}

extern "C" int anemoi_b200_row_stats(const void* x, int64_t ldx, int x_dtype, float* stats, int64_t M, int64_t C, float eps, int flags, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && C >= 8 && C % 8 == 0 && C <= 2048 && ldx >= C, "row_stats: need C % 8 == 0, 8 <= C <= 2048");
  ANEMOI_CHECK_ARG(x_dtype == ANEMOI_F32 || x_dtype == ANEMOI_BF16, "row_stats: bad dtype");
  if (M == 0) return 0;
  ANEMOI_CHECK_ARG(x && stats, "row_stats: null pointer");
  const int es = x_dtype == ANEMOI_BF16 ? 2 : 4;
  ANEMOI_CHECK_ARG(aligned16(x) && (ldx * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(stats) & 7) == 0, "row_stats: rows must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = (M + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (x_dtype == ANEMOI_BF16 && C % 64 == 0) {
    int64_t b2 = (M + 31) / 32;  // 4 rows per warp, 8 warps per block
    if (b2 > cap) b2 = cap;
    cudaError_t le = launch_pdl(row_stats_stream_kernel, dim3((unsigned)b2), dim3(256), 0, s, (const __nv_bfloat16*)x, ldx, reinterpret_cast<float2*>(stats),
                                M, (int)C, eps, (flags & ANEMOI_EPI_REVERSE) ? 1 : 0);
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(row_stats_stream_kernel)");
    return launch_status("row_stats_stream_kernel");
  }
#define RS_LAUNCH(T, IT) row_stats_kernel<T, IT><<<(unsigned)blocks, 256, 0, s>>>((const T*)x, ldx, reinterpret_cast<float2*>(stats), M, (int)C, eps)
#define RS_PICK(T)        \
  if (C <= 256)           \
    RS_LAUNCH(T, 1);      \
  else if (C <= 512)      \
    RS_LAUNCH(T, 2);      \
  else if (C <= 1024)     \
    RS_LAUNCH(T, 4);      \
  else                    \
    RS_LAUNCH(T, 8)
  if (x_dtype == ANEMOI_BF16) {
    RS_PICK(__nv_bfloat16);
  } else {
    RS_PICK(float);
  }
#undef RS_PICK
#undef RS_LAUNCH
  return launch_status("row_stats_kernel");
}
