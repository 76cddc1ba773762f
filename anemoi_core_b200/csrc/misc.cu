// misc.cu — error plumbing, CSR build (integer path), cast/pad/gather copies.
#include "common.cuh"

namespace anemoi {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return -2;
}

// ---- CSR build ---------------------------------------------------------------------------------------------
// colptr[i] = #{ e : dst[e] < i }  for dst non-decreasing == first edge position whose dst >= i.
// One thread per edge boundary: edge e (0..E) fills colptr[(dst[e-1], dst[e]]] = e  (dst[-1] = -1, dst[E] = n_dst).
// Each colptr entry is written by exactly one thread => no atomics, bit-exact with index2ptr on sorted input
// (triton/utils.py:61: index2ptr(col, num_nodes[1])).  Also narrows src to int32 and validates.
__global__ void csr_build_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E, int64_t n_src, int64_t n_dst,
                                 int64_t* __restrict__ colptr64, int32_t* __restrict__ colptr32, int32_t* __restrict__ src32,
                                 int32_t* __restrict__ dst32, int32_t* __restrict__ status) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > E) return;
  int64_t hi = (e == E) ? n_dst : dst[e];
  int64_t lo = (e == 0) ? -1 : dst[e - 1];
  int bad = 0;
  if (e < E) {
    int64_t s = src[e];
    if (s < 0 || s >= n_src || hi < 0 || hi >= n_dst) bad |= 2;
    src32[e] = (int32_t)s;
    if (dst32) dst32[e] = (int32_t)hi;
  }
  if (hi < lo) bad |= 1;
  if (bad) {
    atomicOr(status, bad);
    return;
  }
  if (lo < -1) lo = -1;
  if (hi > n_dst) hi = n_dst;
  for (int64_t i = lo + 1; i <= hi; ++i) {
    if (colptr64) colptr64[i] = e;
    colptr32[i] = (int32_t)e;
  }
}

// ---- cast / pad / gather -----------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void cast_pad_kernel(const TI* __restrict__ in, int64_t ldi, const int32_t* __restrict__ idx, TO* __restrict__ out, int64_t ldo,
                                int64_t M, int64_t K, int64_t Kpad) {
  int64_t total = M * Kpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = i / Kpad, c = i - m * Kpad;
    float v = 0.f;
    if (c < K) {
      int64_t r = idx ? (int64_t)idx[m] : m;
      v = to_f32<TI>(in[r * ldi + c]);
    }
    out[m * ldo + c] = from_f32<TO>(v);
  }
}

// 4 columns per thread: 16-byte (fp32) / 8-byte (bf16) accesses; the scalar kernel above reaches ~1.2 TB/s, this one is HBM-bound
template <typename TI, typename TO>
__global__ void cast_pad_vec4_kernel(const TI* __restrict__ in, int64_t ldi, TO* __restrict__ out, int64_t ldo, int64_t M, int64_t K, int64_t Kpad) {
  const int64_t q = Kpad >> 2, total = M * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / q, c = (i - m * q) << 2;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < K) load_vec_f32<TI, 4>(in + m * ldi + c, v);  // K % 4 == 0: a group is entirely data or entirely padding
    if constexpr (sizeof(TO) == 4) {
      *reinterpret_cast<float4*>(out + m * ldo + c) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      *reinterpret_cast<uint2*>(out + m * ldo + c) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
  }
}

// bf16 + bf16 -> bf16, 8 columns per thread (16-byte accesses)
__global__ void add_bf16_vec8_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, const __nv_bfloat16* __restrict__ b, int64_t ldb,
                                     __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t M, int64_t C) {
  const int64_t q = C >> 3, total = M * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / q, c = (i - m * q) << 3;
    float x[8], y[8];
    load_vec_f32<__nv_bfloat16, 8>(a + m * lda + c, x);
    load_vec_f32<__nv_bfloat16, 8>(b + m * ldb + c, y);
    *reinterpret_cast<uint4*>(out + m * ldo + c) = make_uint4(pack_bf16x2(x[0] + y[0], x[1] + y[1]), pack_bf16x2(x[2] + y[2], x[3] + y[3]),
                                                              pack_bf16x2(x[4] + y[4], x[5] + y[5]), pack_bf16x2(x[6] + y[6], x[7] + y[7]));
  }
}

__global__ void add_kernel(const void* __restrict__ a, int64_t lda, int adt, const void* __restrict__ b, int64_t ldb, int bdt, void* __restrict__ out,
                           int64_t ldo, int odt, int64_t M, int64_t C) {
  const int64_t total = M * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / C, c = i - m * C;
    store_from_f32(out, m * ldo + c, odt, load_as_f32(a, m * lda + c, adt) + load_as_f32(b, m * ldb + c, bdt));
  }
}

// ---- gated feed-forward layers (layers/mlp.py:38-53 GatedMLPLayer): out = act(gate) * value, gate | value = the two halves of ONE GEMM
// output row (gate_proj and value_proj run as one GEMM on concatenated weights).  act: 0 sigmoid (glu), 1 silu (swiglu), 2 gelu-erf (geglu),
// 3 relu (reglu).  fp32 math, one rounding; HBM-bound: reads 2H, writes H per row.
template <typename T, int VEC>
__global__ void glu_combine_kernel(const T* __restrict__ in, int64_t ldi, T* __restrict__ out, int64_t ldo, int64_t M, int64_t H, int act) {
  const int64_t q = H / VEC, total = M * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / q, c = (i - m * q) * VEC;
    float g[VEC], v[VEC], y[VEC];
    load_vec_f32<T, VEC>(in + m * ldi + c, g);
    load_vec_f32<T, VEC>(in + m * ldi + H + c, v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float a;
      if (act == 0) a = 1.0f / (1.0f + __expf(-g[j]));
      else if (act == 1) a = g[j] / (1.0f + __expf(-g[j]));
      else if (act == 2) a = gelu_erf(g[j]);
      else a = fmaxf(g[j], 0.f);
      y[j] = a * v[j];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[m * ldo + c + j] = from_f32<T>(y[j]);
  }
}

// Backward of glu_combine: for y = act(g) * v with the cotangent dy,  din[:, :H] = dy * v * act'(g)  and  din[:, H:] = dy * act(g)
// (the gradient of the [gate | value] GEMM output; PyTorch autograd of layers/mlp.py:38-53 in the reference).  fp32 math, one rounding.
template <typename T>
__global__ void glu_combine_bwd_kernel(const T* __restrict__ in, int64_t ldi, const T* __restrict__ dy, int64_t lddy, T* __restrict__ din, int64_t ldd,
                                       int64_t M, int64_t H, int act) {
  const int64_t total = M * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / H, c = i - m * H;
    const float g = to_f32<T>(in[m * ldi + c]), v = to_f32<T>(in[m * ldi + H + c]), d = to_f32<T>(dy[m * lddy + c]);
    float a, da;
    if (act == 0) {
      a = 1.0f / (1.0f + __expf(-g));
      da = a * (1.0f - a);
    } else if (act == 1) {
      const float sg = 1.0f / (1.0f + __expf(-g));
      a = g * sg;
      da = sg * (1.0f + g * (1.0f - sg));
    } else if (act == 2) {
      const float cdf = 0.5f * (1.0f + erff(g * 0.70710678118654752440f));
      a = g * cdf;
      da = cdf + g * 0.3989422804014327f * __expf(-0.5f * g * g);
    } else {
      a = fmaxf(g, 0.f);
      da = g > 0.f ? 1.0f : 0.f;
    }
    din[m * ldd + c] = from_f32<T>(d * v * da);
    din[m * ldd + H + c] = from_f32<T>(d * a);
  }
}

// ---- fp32 GEMMs on the bf16 tensor cores: exact three-way split --------------------------------------------------------------
// x = x1 + x2 + x3 with x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2): 24 mantissa bits, i.e. the fp32 value (up to 2^-24 |x|).
// A product of two bf16 numbers is exact in fp32, so  a.w = sum_{i+j<=4} a_i w_j  (6 terms; the dropped ones are below 2^-24 |a||w|)
// evaluated with fp32 accumulation is an fp32-grade dot product.  The six partial GEMMs are ONE tcgen05 GEMM over a 6K-long inner
// dimension: A' = [a3|a2|a1|a2|a1|a1], W' = [w1|w2|w3|w1|w2|w1].  This kernel writes the [M, 6K] bf16 operand for either side.
// (The fp32 parity mode used to run its GEMMs on the FFMA kernel: 201 ms per cfg2 step against 131 ms for the reference's cuBLAS fp32.)
__global__ void split_bf16x3_kernel(const float* __restrict__ in, int64_t ldi, __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t M, int64_t K,
                                    int weight_side) {
  const int64_t q = K / 4, total = M * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / q, k = (i - m * q) * 4;
    const float4 v = *reinterpret_cast<const float4*>(in + m * ldi + k);
    const float x[4] = {v.x, v.y, v.z, v.w};
    uint32_t p[3][2];  // three parts x two packed pairs
    float r[4], h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = x[j];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[j] = __bfloat162float(__float2bfloat16_rn(r[j]));
        r[j] -= h[j];
      }
      p[t][0] = pack_bf16x2(h[0], h[1]), p[t][1] = pack_bf16x2(h[2], h[3]);
    }
    // pairs in the order (3,1) (2,2) (1,3) (2,1) (1,2) (1,1): SMALLEST terms first.  The tensor core's fp32 accumulation truncates; while
    // the corrections are summed the accumulator is still 2^-8 of its final size, so only the K/16 MMAs of the (1,1) block can lose bits
    // at full scale (measured: 2e-5 -> 3e-6 of the output scale against the largest-first order).
    const int pa[6] = {2, 1, 0, 1, 0, 0}, pw[6] = {0, 1, 2, 0, 1, 0};
    __nv_bfloat16* o = out + m * ldo + k;
#pragma unroll
    for (int s6 = 0; s6 < 6; ++s6) {
      const int t = weight_side ? pw[s6] : pa[s6];
      *reinterpret_cast<uint2*>(o + s6 * K) = make_uint2(p[t][0], p[t][1]);
    }
  }
}

// ---- model glue either side of the encoder / decoder ------------------------------------------------------
// assemble_input: out[(b e g), t * V + v] = x[b, t, e, g, v];  out[(b e g), T * V + a] = attrs[row % attr_rows, a];  zero pad up to Kpad.
// Replaces einops.rearrange + torch.cat (+ the autocast cast of the embedding Linear) of `_assemble_input`
// (models/encoder_processor_decoder.py:98-127) with one pass that writes the embedding GEMM's A operand directly.
// One warp per output row: T contiguous V-float reads, one contiguous A-float read, one contiguous row write.
template <typename TO>
__global__ void assemble_input_kernel(const float* __restrict__ x, int64_t B, int64_t T, int64_t E, int64_t G, int64_t V,
                                      const float* __restrict__ attrs, int64_t A, int64_t attr_rows, TO* __restrict__ out, int64_t ldo,
                                      int64_t Kpad) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = B * E * G;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t g = row % G, be = row / G, e = be % E, b = be / E;
    TO* o = out + row * ldo;
    for (int64_t t = 0; t < T; ++t) {
      const float* xr = x + (((b * T + t) * E + e) * G + g) * V;
      for (int64_t v = lane; v < V; v += 32) o[t * V + v] = from_f32<TO>(__ldg(xr + v));
    }
    const float* ar = attrs ? attrs + (row % attr_rows) * A : nullptr;
    for (int64_t a = lane; a < Kpad - T * V; a += 32) o[T * V + a] = from_f32<TO>(a < A ? __ldg(ar + a) : 0.f);
  }
}

// assemble_output: y[b, t, e, g, v] = bound_v( dec[(b e g), t * V + v] + (skip_src[v] >= 0 ? x[b, step, e, g, skip_src[v]] : 0) )
// with bound_v = identity / relu / leaky_relu(0.01) per output variable.  Replaces the rearrange, dtype cast, clone, indexed residual
// add and the index_put of every bounding layer of `_assemble_output` (models/encoder_processor_decoder.py:129-163,
// layers/residual.py:60-81 SkipConnection, layers/bounding.py:81-94).
template <typename TI>
__global__ void assemble_output_kernel(const TI* __restrict__ dec, int64_t ldd, const float* __restrict__ x, int64_t B, int64_t T_in, int64_t E,
                                       int64_t G, int64_t V_in, int64_t step, const int32_t* __restrict__ skip_src,
                                       const int32_t* __restrict__ bound, float* __restrict__ y, int64_t T_out, int64_t V_out) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = B * E * G;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t g = row % G, be = row / G, e = be % E, b = be / E;
    const float* xr = x ? x + (((b * T_in + step) * E + e) * G + g) * V_in : nullptr;
    for (int64_t v = lane; v < V_out; v += 32) {
      const int32_t src = skip_src ? __ldg(skip_src + v) : -1;
      const float sk = (src >= 0 && xr) ? __ldg(xr + src) : 0.f;
      const int32_t bd = bound ? __ldg(bound + v) : 0;
      for (int64_t t = 0; t < T_out; ++t) {
        float val = to_f32<TI>(dec[row * ldd + t * V_out + v]) + sk;
        if (bd == 1) val = fmaxf(val, 0.f);
        else if (bd == 2) val = val > 0.f ? val : 0.01f * val;
        y[(((b * T_out + t) * E + e) * G + g) * V_out + v] = val;
      }
    }
  }
}

}  // namespace anemoi

using namespace anemoi;

extern "C" int anemoi_b200_glu_combine(const void* in, int64_t ldi, void* out, int64_t ldo, int64_t M, int64_t H, int act, int dtype, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && H >= 1 && ldi >= 2 * H && ldo >= H, "glu_combine: bad shape");
  ANEMOI_CHECK_ARG(act >= 0 && act <= 3, "glu_combine: activation code %d (0 sigmoid, 1 silu, 2 gelu, 3 relu)", act);
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "glu_combine: bad dtype %d", dtype);
  if (M == 0) return 0;
  ANEMOI_CHECK_ARG(in && out, "glu_combine: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int es = dtype == ANEMOI_BF16 ? 2 : 4;
  const bool vec = H % 4 == 0 && ldi % 4 == 0 && (reinterpret_cast<uintptr_t>(in) % (4 * es)) == 0;
  int64_t blocks = (M * (vec ? H / 4 : H) + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  if (dtype == ANEMOI_BF16) {
    if (vec) glu_combine_kernel<__nv_bfloat16, 4><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)in, ldi, (__nv_bfloat16*)out, ldo, M, H, act);
    else glu_combine_kernel<__nv_bfloat16, 1><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)in, ldi, (__nv_bfloat16*)out, ldo, M, H, act);
  } else {
    if (vec) glu_combine_kernel<float, 4><<<(unsigned)blocks, 256, 0, s>>>((const float*)in, ldi, (float*)out, ldo, M, H, act);
    else glu_combine_kernel<float, 1><<<(unsigned)blocks, 256, 0, s>>>((const float*)in, ldi, (float*)out, ldo, M, H, act);
  }
  return launch_status("glu_combine_kernel");
}

extern "C" int anemoi_b200_glu_combine_bwd(const void* in, int64_t ldi, const void* dy, int64_t lddy, void* din, int64_t ldd, int64_t M, int64_t H,
                                          int act, int dtype, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && H >= 1 && ldi >= 2 * H && lddy >= H && ldd >= 2 * H, "glu_combine_bwd: bad shape");
  ANEMOI_CHECK_ARG(act >= 0 && act <= 3, "glu_combine_bwd: activation code %d (0 sigmoid, 1 silu, 2 gelu, 3 relu)", act);
  ANEMOI_CHECK_ARG(dtype == ANEMOI_F32 || dtype == ANEMOI_BF16, "glu_combine_bwd: bad dtype %d", dtype);
  if (M == 0) return 0;
  ANEMOI_CHECK_ARG(in && dy && din, "glu_combine_bwd: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t blocks = (M * H + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  if (dtype == ANEMOI_BF16)
    glu_combine_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)in, ldi, (const __nv_bfloat16*)dy, lddy,
                                                                            (__nv_bfloat16*)din, ldd, M, H, act);
  else
    glu_combine_bwd_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)in, ldi, (const float*)dy, lddy, (float*)din, ldd, M, H, act);
  return launch_status("glu_combine_bwd_kernel");
}

extern "C" int anemoi_b200_split_bf16x3(const float* in, int64_t ldi, void* out, int64_t ldo, int64_t M, int64_t K, int weight_side, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && K >= 4 && K % 4 == 0 && ldi >= K && ldo >= 6 * K, "split_bf16x3: need K %% 4 == 0, ldi >= K, ldo >= 6 K");
  if (M == 0) return 0;
  ANEMOI_CHECK_ARG(in && out, "split_bf16x3: null pointer");
  ANEMOI_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && ldi % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0 && ldo % 4 == 0,
                   "split_bf16x3: rows must be 16-byte (input) / 8-byte (output) aligned");
  int64_t blocks = (M * (K / 4) + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  split_bf16x3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, ldi, (__nv_bfloat16*)out, ldo, M, K, weight_side ? 1 : 0);
  return launch_status("split_bf16x3_kernel");
}

extern "C" int anemoi_b200_assemble_input(const float* x, int64_t B, int64_t T, int64_t E, int64_t G, int64_t V, const float* attrs, int64_t A,
                                          int64_t attr_rows, void* out, int64_t ldo, int64_t Kpad, int o_dtype, void* stream) {
  ANEMOI_CHECK_ARG(B >= 0 && T >= 0 && E >= 0 && G >= 0 && V >= 0 && A >= 0, "assemble_input: negative size");
  ANEMOI_CHECK_ARG(Kpad >= T * V + A && ldo >= Kpad, "assemble_input: Kpad / ldo too small");
  const int64_t rows = B * E * G;
  if (rows == 0 || Kpad == 0) return 0;
  ANEMOI_CHECK_ARG(A == 0 || (attrs && attr_rows > 0), "assemble_input: attributes without rows");
  ANEMOI_CHECK_ARG(out && (x || T * V == 0), "assemble_input: null pointer");
  int64_t blocks = (rows + 7) / 8;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t s = (cudaStream_t)stream;
  if (o_dtype == ANEMOI_F32)
    assemble_input_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(x, B, T, E, G, V, attrs, A, attr_rows, (float*)out, ldo, Kpad);
  else if (o_dtype == ANEMOI_BF16)
    assemble_input_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>(x, B, T, E, G, V, attrs, A, attr_rows, (__nv_bfloat16*)out, ldo, Kpad);
  else {
    set_error("assemble_input: bad dtype %d", o_dtype);
    return -1;
  }
  return launch_status("assemble_input_kernel");
}

extern "C" int anemoi_b200_assemble_output(const void* dec, int64_t ldd, int d_dtype, const float* x, int64_t B, int64_t T_in, int64_t E, int64_t G,
                                           int64_t V_in, int64_t step, const int32_t* skip_src, const int32_t* bound, float* y, int64_t T_out,
                                           int64_t V_out, void* stream) {
  ANEMOI_CHECK_ARG(B >= 0 && E >= 0 && G >= 0 && T_out >= 0 && V_out >= 0 && ldd >= T_out * V_out, "assemble_output: bad shape");
  ANEMOI_CHECK_ARG(!skip_src || !x || (step >= 0 && step < T_in && V_in > 0), "assemble_output: residual step out of range");
  const int64_t rows = B * E * G;
  if (rows == 0 || T_out * V_out == 0) return 0;
  ANEMOI_CHECK_ARG(dec && y, "assemble_output: null pointer");
  int64_t blocks = (rows + 7) / 8;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t s = (cudaStream_t)stream;
  if (d_dtype == ANEMOI_F32)
    assemble_output_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)dec, ldd, x, B, T_in, E, G, V_in, step, skip_src, bound, y, T_out, V_out);
  else if (d_dtype == ANEMOI_BF16)
    assemble_output_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)dec, ldd, x, B, T_in, E, G, V_in, step, skip_src, bound,
                                                                          y, T_out, V_out);
  else {
    set_error("assemble_output: bad dtype %d", d_dtype);
    return -1;
  }
  return launch_status("assemble_output_kernel");
}

extern "C" int anemoi_b200_abi_version(void) { return 6; }
extern "C" const char* anemoi_b200_last_error(void) { return g_err; }

extern "C" int anemoi_b200_csr_build(const int64_t* edge_index, int64_t n_edges, int64_t n_src, int64_t n_dst, int64_t* colptr64,
                                     int32_t* colptr32, int32_t* src32, int32_t* dst32, int32_t* status, void* stream) {
  ANEMOI_CHECK_ARG(n_edges >= 0 && n_src >= 0 && n_dst >= 0, "csr_build: negative size");
  ANEMOI_CHECK_ARG(n_edges < (1ll << 31) && n_src < (1ll << 31) && n_dst < (1ll << 31) - 1, "csr_build: sizes must fit int32");
  ANEMOI_CHECK_ARG(colptr32 && status && (n_edges == 0 || (edge_index && src32)), "csr_build: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  ANEMOI_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
  int threads = 256;
  int64_t blocks = (n_edges + 1 + threads - 1) / threads;
  csr_build_kernel<<<(unsigned)blocks, threads, 0, s>>>(edge_index, edge_index ? edge_index + n_edges : nullptr, n_edges, n_src, n_dst,
                                                        colptr64, colptr32, src32, dst32, status);
  return launch_status("csr_build_kernel");
}

extern "C" int anemoi_b200_cast_pad(const void* in, int64_t ldi, int i_dtype, const int32_t* idx, void* out, int64_t ldo, int o_dtype,
                                    int64_t M, int64_t K, int64_t Kpad, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && K >= 0 && Kpad >= K && ldo >= Kpad && ldi >= K, "cast_pad: bad shape");
  if (M == 0 || Kpad == 0) return 0;
  ANEMOI_CHECK_ARG(in && out, "cast_pad: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int threads = 256;
  int64_t blocks = (M * Kpad + threads - 1) / threads;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  const int is = i_dtype == ANEMOI_BF16 ? 2 : 4, os = o_dtype == ANEMOI_BF16 ? 2 : 4;
  const bool vec4 = !idx && K % 4 == 0 && Kpad % 4 == 0 && ldi % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(in) % (4 * is)) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) % (4 * os)) == 0 && (i_dtype | 1) == 1 && (o_dtype | 1) == 1;
  if (vec4) {
    blocks = (M * (Kpad / 4) + threads - 1) / threads;
    if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  }
#define LAUNCH(TI, TO)                                                                                                               \
  do {                                                                                                                               \
    if (vec4)                                                                                                                        \
      cast_pad_vec4_kernel<TI, TO><<<(unsigned)blocks, threads, 0, s>>>((const TI*)in, ldi, (TO*)out, ldo, M, K, Kpad);              \
    else                                                                                                                             \
      cast_pad_kernel<TI, TO><<<(unsigned)blocks, threads, 0, s>>>((const TI*)in, ldi, idx, (TO*)out, ldo, M, K, Kpad);              \
  } while (0)
  if (i_dtype == ANEMOI_F32 && o_dtype == ANEMOI_F32)
    LAUNCH(float, float);
  else if (i_dtype == ANEMOI_F32 && o_dtype == ANEMOI_BF16)
    LAUNCH(float, __nv_bfloat16);
  else if (i_dtype == ANEMOI_BF16 && o_dtype == ANEMOI_F32)
    LAUNCH(__nv_bfloat16, float);
  else if (i_dtype == ANEMOI_BF16 && o_dtype == ANEMOI_BF16)
    LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else {
    set_error("cast_pad: bad dtype %d -> %d", i_dtype, o_dtype);
    return -1;
  }
#undef LAUNCH
  return launch_status("cast_pad_kernel");
}

extern "C" int anemoi_b200_add(const void* a, int64_t lda, int a_dtype, const void* b, int64_t ldb, int b_dtype, void* out, int64_t ldo,
                               int o_dtype, int64_t M, int64_t C, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && C >= 0 && lda >= C && ldb >= C && ldo >= C, "add: bad shape");
  ANEMOI_CHECK_ARG((a_dtype | 1) == 1 && (b_dtype | 1) == 1 && (o_dtype | 1) == 1, "add: bad dtype");
  if (M == 0 || C == 0) return 0;
  ANEMOI_CHECK_ARG(a && b && out, "add: null pointer");
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (a_dtype == ANEMOI_BF16 && b_dtype == ANEMOI_BF16 && o_dtype == ANEMOI_BF16 && C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0 &&
      a16(a) && a16(b) && a16(out)) {
    int64_t vblocks = (M * (C / 8) + 255) / 256;
    if (vblocks > (int64_t)num_sms() * 16) vblocks = (int64_t)num_sms() * 16;
    add_bf16_vec8_kernel<<<(unsigned)vblocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b, ldb,
                                                                              (__nv_bfloat16*)out, ldo, M, C);
    return launch_status("add_bf16_vec8_kernel");
  }
  int64_t blocks = (M * C + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  add_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, lda, a_dtype, b, ldb, b_dtype, out, ldo, o_dtype, M, C);
  return launch_status("add_kernel");
}
