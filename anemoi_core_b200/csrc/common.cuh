// common.cuh — shared helpers for libanemoi_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../include/anemoi_b200.h"

namespace anemoi {

// ---- error reporting across the C ABI (thread-local message, errno-style codes) ----------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define ANEMOI_CHECK_ARG(cond, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      ::anemoi::set_error(__VA_ARGS__);  \
      return -1;                         \
    }                                    \
  } while (0)

#define ANEMOI_CUDA(call)                                         \
  do {                                                            \
    cudaError_t _e = (call);                                      \
    if (_e != cudaSuccess) return ::anemoi::cuda_fail(_e, #call); \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cuda_fail(e, what);
  }
  return 0;
}

// Per-device caches: function attributes (cudaFuncSetAttribute), occupancy results and the SM count belong to the CURRENT device of the
// calling thread; a process that drives several GPUs must not reuse what it learned on the first one.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0) n[dev] = 148;  // B200
  }
  return n[dev];
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start (become resident, run its prologue: barrier init,
// TMEM allocation, descriptor prefetch, index arithmetic) while its predecessor in the stream is still draining; griddepcontrol.wait then
// blocks until the predecessor has COMPLETED and its writes are visible, so everything that touches global memory sits after it.  Each
// kernel triggers its own dependents right after its wait (at most one kernel runs ahead).  Both instructions are no-ops for a kernel
// launched without the attribute, and the attribute is captured into CUDA graphs as a programmatic edge.  ANEMOI_B200_PDL=0 turns the
// attribute off (A/B).  What it hides: ~2-3 us of launch latency + prologue per launch, 109 launches per cfg2 step, ~180 per 8-GPU step.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ANEMOI_B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
// <<<grid, block, smem, stream>>> with the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- dtype helpers -------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// runtime-typed scalar access (epilogues: residual / out may be f32 or bf16)
__device__ __forceinline__ float load_as_f32(const void* p, int64_t i, int dtype) {
  return dtype == ANEMOI_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void store_from_f32(void* p, int64_t i, int dtype, float v) {
  if (dtype == ANEMOI_BF16)
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(p)[i] = v;
}

// exact (erf) GELU, torch.nn.GELU() default — layers/utils.py:107-110
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Same function to 5e-7 absolute with ONE MUFU per element.  gelu(x) = max(x, 0) - |x| * Phi(-|x|), and log2 Phi(-t) is smooth on
// t >= 0, so Phi(-t) = exp2(P5(t)) with a degree-5 polynomial (weighted minimax fit of log2 Phi(-t), weight t * Phi(-t) = the
// sensitivity of gelu; |gelu error| <= 4.7e-7 evaluated in fp32, see tests/test_host_logic.py::test_gelu_exp2_polynomial).
// The classical erfc forms (A&S 7.1.26) need rcp + ex2 = two MUFU per element, which at 16 MUFU/clk/SM costs exactly the tile's
// MMA time for K = 512 and made the mlp1 epilogue the critical path; this form is 5 FMA + ex2.  t is clamped to 10 (Phi(-10) ~ 1e-23).
#define ANEMOI_GELU_P5 -1.00003763f, -1.15078777f, -0.459992649f, -0.0518271656f, 0.00708446018f, -0.000473293894f
__device__ __forceinline__ float gelu_erf_fast(float x) {
  constexpr float c[6] = {ANEMOI_GELU_P5};
  const float t = fminf(fabsf(x), 10.f);
  float p = fmaf(c[5], t, c[4]);
  p = fmaf(p, t, c[3]);
  p = fmaf(p, t, c[2]);
  p = fmaf(p, t, c[1]);
  p = fmaf(p, t, c[0]);
  float ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(p));
  return fmaf(-t, ex, fmaxf(x, 0.f));
}

// Two GELUs per call on the packed fp32x2 pipe (FFMA2: one issue slot for two lanes of work).
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
  constexpr float c[6] = {ANEMOI_GELU_P5};
  const float2 t = make_float2(fminf(fabsf(x.x), 10.f), fminf(fabsf(x.y), 10.f));
  float2 p = __ffma2_rn(make_float2(c[5], c[5]), t, make_float2(c[4], c[4]));
  p = __ffma2_rn(p, t, make_float2(c[3], c[3]));
  p = __ffma2_rn(p, t, make_float2(c[2], c[2]));
  p = __ffma2_rn(p, t, make_float2(c[1], c[1]));
  p = __ffma2_rn(p, t, make_float2(c[0], c[0]));
  float2 ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(p.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(p.y));
  return __ffma2_rn(make_float2(-t.x, -t.y), ex, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

// load VEC contiguous elements (VEC*sizeof(T) bytes, 16B vectors where possible) as fp32
template <typename T, int VEC>
__device__ __forceinline__ void load_vec_f32(const T* __restrict__ p, float (&out)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 4; ++i) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        out[4 * i] = t.x, out[4 * i + 1] = t.y, out[4 * i + 2] = t.z, out[4 * i + 3] = t.w;
      }
    } else if constexpr (VEC % 2 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p) + i);
        out[2 * i] = t.x, out[2 * i + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) out[i] = __ldg(reinterpret_cast<const float*>(p) + i);
    }
  } else {
    if constexpr (VEC % 8 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 8; ++i) {
        uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + i);
        float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
        out[8 * i] = a.x, out[8 * i + 1] = a.y, out[8 * i + 2] = b.x, out[8 * i + 3] = b.y;
        out[8 * i + 4] = c.x, out[8 * i + 5] = c.y, out[8 * i + 6] = d.x, out[8 * i + 7] = d.y;
      }
    } else if constexpr (VEC % 4 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 4; ++i) {
        uint2 t = __ldg(reinterpret_cast<const uint2*>(p) + i);
        float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
        out[4 * i] = a.x, out[4 * i + 1] = a.y, out[4 * i + 2] = b.x, out[4 * i + 3] = b.y;
      }
    } else if constexpr (VEC % 2 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) {
        float2 a = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(p) + i));
        out[2 * i] = a.x, out[2 * i + 1] = a.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) out[i] = __bfloat162float(p[i]);
    }
  }
}

template <typename T, int VEC>
__device__ __forceinline__ void store_vec_f32(T* __restrict__ p, const float (&v)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if constexpr (VEC % 2 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) reinterpret_cast<float*>(p)[i] = v[i];
    }
  } else {
    if constexpr (VEC % 8 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 8; ++i)
        reinterpret_cast<uint4*>(p)[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                                    pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
    } else if constexpr (VEC % 4 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 4; ++i)
        reinterpret_cast<uint2*>(p)[i] = make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
    } else if constexpr (VEC % 2 == 0) {
#pragma unroll
      for (int i = 0; i < VEC / 2; ++i) reinterpret_cast<uint32_t*>(p)[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) p[i] = __float2bfloat16_rn(v[i]);
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- epilogue shared by the two GEMM kernels --------------------------------------------------------------
struct EpiParams {
  const float* bias;    // [N] or null
  const float* g1;      // gather-add table 1 [*, ldg] fp32 or null
  const int32_t* idx1;  // [M]
  const float* g2;
  const int32_t* idx2;
  int64_t ldg;
  const void* residual;  // [M, ldr] or null
  int64_t ldr;
  int r_dtype;
  void* out;  // [M, ldo]
  int64_t ldo;
  int o_dtype;
  int64_t M, N;
  int flags;
  // folded LayerNorm of the A operand: out = rstd_m * (acc - mean_m * colsum_n) + bias_n, applied before GELU
  const float* ln_stats;   // [M, 2] = (mean, rstd) per row (ln_parts == 0) or [M, ln_parts, 2] partial (block mean, block M2), or null
  const float* ln_colsum;  // [N] = sum_k W'[n, k] of the gamma-scaled weight
  // Row statistics produced by one GEMM's epilogue for the LayerNorm folded into the next: stats_out [M, ceil(N / 64), 2] receives, per
  // 64-column block, (mean, sum of squared deviations from that mean) of the output row AS STORED (after rounding to o_dtype); the
  // consumer passes the buffer as ln_stats with ln_parts = ceil(K / 64), ln_dim = K (the normalised width) and ln_eps.
  float* stats_out;
  int ln_parts, ln_dim;
  float ln_eps;
};

constexpr int kStatsBlock = 64;  // columns per partial of stats_out

// (mean, rstd) of row m for the folded LayerNorm, from either form of ln_stats.  Partial form: per 64-column block (block mean, M2 = sum of
// squared deviations from the block mean), merged with Chan's formula - no E[x^2] - mean^2 cancellation, whatever the row's mean.
__device__ __forceinline__ float2 ln_row_mean_rstd(const EpiParams& ep, int64_t m) {
  if (ep.ln_parts <= 0) return __ldg(reinterpret_cast<const float2*>(ep.ln_stats) + m);
  const float2* p = reinterpret_cast<const float2*>(ep.ln_stats) + m * ep.ln_parts;
  const int last_n = ep.ln_dim - (ep.ln_parts - 1) * kStatsBlock;  // columns in the (possibly ragged) last block
  float s = 0.f;
  for (int i = 0; i < ep.ln_parts; ++i) s += __ldg(p + i).x * (float)(i == ep.ln_parts - 1 ? last_n : kStatsBlock);
  const float inv = 1.0f / (float)ep.ln_dim;
  const float mean = s * inv;
  float m2 = 0.f;
  for (int i = 0; i < ep.ln_parts; ++i) {
    const float2 t = __ldg(p + i);
    const float d = t.x - mean;
    m2 += t.y + (float)(i == ep.ln_parts - 1 ? last_n : kStatsBlock) * d * d;
  }
  return make_float2(mean, rsqrtf(m2 * inv + ep.ln_eps));
}

}  // namespace anemoi
