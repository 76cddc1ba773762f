// graphconv.cu — tail of GraphConv (layers/conv.py:73-81): e' = LayerNorm(h) + e, out[d] = sum of e' over the
// edges into d.  Edges are dst-sorted, so a destination's edges are one contiguous run [colptr[d], colptr[d+1]):
// one warp walks the run, normalises each row in registers (fp32 statistics), writes e' once and keeps the
// running sum in registers — a deterministic segmented reduction with no atomics and no re-read of e'.
// HBM traffic per edge: read h, read e, write e' (3*C*b) + out once per node.
#include "common.cuh"

namespace anemoi {

template <typename T, int ITERS>
__global__ void __launch_bounds__(256)
    graphconv_ln_aggregate_kernel(const T* __restrict__ h, int64_t ldh, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const T* __restrict__ e, int64_t lde, T* __restrict__ e_new, int64_t ldn, const int32_t* __restrict__ colptr,
                                  T* __restrict__ out, int64_t ldo, int64_t n_dst, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  float gm[ITERS][8], bt[ITERS][8];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = it * 256 + lane * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) gm[it][i] = 1.f, bt[it][i] = 0.f;
    if (c < C) {
      if (gamma) load_vec_f32<float, 8>(gamma + c, gm[it]);
      if (beta) load_vec_f32<float, 8>(beta + c, bt[it]);
    }
  }
  for (int64_t d = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < n_dst; d += warps_total) {
    const int e0 = colptr[d], e1 = colptr[d + 1];
    float acc[ITERS][8];
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[it][i] = 0.f;
    for (int ei = e0; ei < e1; ++ei) {
      float v[ITERS][8];
      float s = 0.f;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c = it * 256 + lane * 8;
        if (c < C) {
          load_vec_f32<T, 8>(h + (int64_t)ei * ldh + c, v[it]);
#pragma unroll
          for (int i = 0; i < 8; ++i) s += v[it][i];
        }
      }
      const float mean = warp_sum(s) / (float)C;
      float qq = 0.f;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c = it * 256 + lane * 8;
        if (c < C) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float dlt = v[it][i] - mean;
            qq += dlt * dlt;
          }
        }
      }
      const float rstd = rsqrtf(warp_sum(qq) / (float)C + eps);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c = it * 256 + lane * 8;
        if (c < C) {
          float r[8], o[8];
          load_vec_f32<T, 8>(e + (int64_t)ei * lde + c, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            o[i] = (v[it][i] - mean) * rstd * gm[it][i] + bt[it][i] + r[i];
            // the reference sums the ROUNDED e' (scatter of the stored tensor, conv.py:79): round before accumulating
            o[i] = to_f32<T>(from_f32<T>(o[i]));
            acc[it][i] += o[i];
          }
          store_vec_f32<T, 8>(e_new + (int64_t)ei * ldn + c, o);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * 256 + lane * 8;
      if (c < C) store_vec_f32<T, 8>(out + d * ldo + c, acc[it]);
    }
  }
}

// any C / alignment
template <typename T>
__global__ void __launch_bounds__(256)
    graphconv_ln_aggregate_generic_kernel(const T* __restrict__ h, int64_t ldh, const float* __restrict__ gamma, const float* __restrict__ beta,
                                          const T* __restrict__ e, int64_t lde, T* __restrict__ e_new, int64_t ldn,
                                          const int32_t* __restrict__ colptr, T* __restrict__ out, int64_t ldo, int64_t n_dst, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t d = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < n_dst; d += warps_total) {
    const int e0 = colptr[d], e1 = colptr[d + 1];
    for (int c = lane; c < C; c += 32) out[d * ldo + c] = from_f32<T>(0.f);
    for (int ei = e0; ei < e1; ++ei) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s += to_f32<T>(h[(int64_t)ei * ldh + c]);
      const float mean = warp_sum(s) / (float)C;
      float qq = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float dlt = to_f32<T>(h[(int64_t)ei * ldh + c]) - mean;
        qq += dlt * dlt;
      }
      const float rstd = rsqrtf(warp_sum(qq) / (float)C + eps);
      for (int c = lane; c < C; c += 32) {
        float o = (to_f32<T>(h[(int64_t)ei * ldh + c]) - mean) * rstd * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f) +
                  to_f32<T>(e[(int64_t)ei * lde + c]);
        const T ot = from_f32<T>(o);
        e_new[(int64_t)ei * ldn + c] = ot;
        // fp32 running sum kept in the output row only for T = float; for bf16 this generic path accumulates in a register per lane-column
        out[d * ldo + c] = from_f32<T>(to_f32<T>(out[d * ldo + c]) + to_f32<T>(ot));
      }
    }
  }
}

template <typename T>
static int launch_gc(const void* h, int64_t ldh, const float* gamma, const float* beta, const void* e, int64_t lde, void* e_new, int64_t ldn,
                     const int32_t* colptr, void* out, int64_t ldo, int64_t n_dst, int C, float eps, bool vec, cudaStream_t s) {
  int64_t blocks = (n_dst + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
#define GC_ARGS (const T*)h, ldh, gamma, beta, (const T*)e, lde, (T*)e_new, ldn, colptr, (T*)out, ldo, n_dst, C, eps
  if (vec && C <= 256)
    graphconv_ln_aggregate_kernel<T, 1><<<(unsigned)blocks, 256, 0, s>>>(GC_ARGS);
  else if (vec && C <= 512)
    graphconv_ln_aggregate_kernel<T, 2><<<(unsigned)blocks, 256, 0, s>>>(GC_ARGS);
  else if (vec && C <= 1024)
    graphconv_ln_aggregate_kernel<T, 4><<<(unsigned)blocks, 256, 0, s>>>(GC_ARGS);
  else
    graphconv_ln_aggregate_generic_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(GC_ARGS);
#undef GC_ARGS
  return launch_status("graphconv_ln_aggregate_kernel");
}

}  // namespace anemoi

using namespace anemoi;

extern "C" int anemoi_b200_graphconv_ln_aggregate(const void* h, int64_t ldh, const float* gamma, const float* beta, const void* e, int64_t lde,
                                                  void* e_new, int64_t ldn, const int32_t* colptr32, void* out, int64_t ldo, int64_t n_dst,
                                                  int64_t C, float eps, int dtype, void* stream) {
  ANEMOI_CHECK_ARG(n_dst >= 0 && C >= 1 && C < (1 << 30), "graphconv_ln_aggregate: bad shape");
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "graphconv_ln_aggregate: bad dtype %d", dtype);
  ANEMOI_CHECK_ARG(ldh >= C && lde >= C && ldn >= C && ldo >= C, "graphconv_ln_aggregate: leading dimension too small");
  if (n_dst == 0) return 0;
  ANEMOI_CHECK_ARG(colptr32 && out, "graphconv_ln_aggregate: null pointer");
  const int es = dtype == ANEMOI_BF16 ? 2 : 4;
  auto al = [&](const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * es) % 16 == 0; };
  const bool vec = C % 8 == 0 && C <= 1024 && al(h, ldh) && al(e, lde) && al(e_new, ldn) && al(out, ldo) &&
                   (!gamma || (reinterpret_cast<uintptr_t>(gamma) & 15) == 0) && (!beta || (reinterpret_cast<uintptr_t>(beta) & 15) == 0);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ANEMOI_BF16)
    return launch_gc<__nv_bfloat16>(h, ldh, gamma, beta, e, lde, e_new, ldn, colptr32, out, ldo, n_dst, (int)C, eps, vec, s);
  return launch_gc<float>(h, ldh, gamma, beta, e, lde, e_new, ldn, colptr32, out, ldo, n_dst, (int)C, eps, vec, s);
}
