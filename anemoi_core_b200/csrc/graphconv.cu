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

// ---- pipelined version (cp.async ring) -----------------------------------------------------------------------------
// Streaming kernel: per edge it reads one h row and one e row and writes one e' row (3*C*b bytes), per node it writes one out
// row.  Each warp owns a contiguous range of dst nodes holding ~E/#warps edges (edge-balanced, boundaries by a 32-ary search of
// colptr) and streams the h | e rows of its contiguous edge run through a private shared-memory ring, kGcSlots edges ahead of the
// LayerNorm arithmetic, so the loads never wait on the reduction's dependency chain.  Lane l owns the 16-byte chunks (j*32 + l).
#ifndef GC_SLOTS
#define GC_SLOTS 4
#endif
constexpr int kGcSlots = GC_SLOTS;

__device__ __forceinline__ void gc_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void gc_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gc_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 gc_lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int gc_lower_bound(const int32_t* __restrict__ colptr, int n_dst, int x, int lane) {
  int lo = 0, hi = n_dst;
  while (hi - lo > 0) {
    const int step = (hi - lo + 31) / 32;
    const int pos = min(lo + lane * step, hi);
    const unsigned m = __ballot_sync(0xffffffffu, __ldg(colptr + pos) >= x);
    if (m == 0) {
      lo = min(lo + 31 * step, hi) + 1;
      if (lo > hi) return hi;
    } else {
      const int f = __ffs(m) - 1;
      const int nh = min(lo + f * step, hi);
      lo = f == 0 ? nh : min(lo + (f - 1) * step, hi) + 1;
      hi = nh;
      if (lo > hi) lo = hi;
    }
  }
  return lo;
}

template <typename T>
struct GcChunk {
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
  // rounds to T and returns the packed 16 bytes; `f` is overwritten with the ROUNDED values (the reference sums the stored tensor)
  __device__ static __forceinline__ uint4 pack_round(float (&f)[EPC]) {
    if constexpr (sizeof(T) == 4) {
      return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    } else {
      const uint4 u = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      unpack(u, f);
      return u;
    }
  }
};

// NCH = 16-byte chunks per lane: C = NCH * 32 * EPC exactly (C = 256/512/1024 for bf16 with NCH = 1/2/4; 128/256/512 for fp32)
template <typename T, int NCH>
__global__ void __launch_bounds__(128) graphconv_ln_aggregate_pipe_kernel(const T* __restrict__ h, int64_t ldh, const float* __restrict__ gamma,
                                                                          const float* __restrict__ beta, const T* __restrict__ e, int64_t lde,
                                                                          T* __restrict__ e_new, int64_t ldn, const int32_t* __restrict__ colptr,
                                                                          T* __restrict__ out, int64_t ldo, int n_dst, float eps) {
  using CT = GcChunk<T>;
  constexpr int EPC = CT::EPC;
  constexpr int C = NCH * 32 * EPC;
  constexpr int kRow = NCH * 512;        // bytes of one row
  constexpr int kSlotBytes = 2 * kRow;   // h | e
  extern __shared__ __align__(16) uint8_t gc_ring[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(gc_ring) + (uint32_t)(wib * kGcSlots * kSlotBytes);
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int w = blockIdx.x * (blockDim.x >> 5) + wib;
  const int n_edges = __ldg(colptr + n_dst);
  const int lo_e = (int)((int64_t)n_edges * w / warps_total), hi_e = (int)((int64_t)n_edges * (w + 1) / warps_total);
  const int n_lo = w == 0 ? 0 : gc_lower_bound(colptr, n_dst, lo_e, lane);
  const int n_hi = w == warps_total - 1 ? n_dst : gc_lower_bound(colptr, n_dst, hi_e, lane);
  if (n_lo >= n_hi) return;
  const int E0 = __ldg(colptr + n_lo), E1 = __ldg(colptr + n_hi);
  const char* hp = reinterpret_cast<const char*>(h) + lane * 16;
  const char* epp = reinterpret_cast<const char*>(e) + lane * 16;
  const int64_t ldh_b = ldh * (int64_t)sizeof(T), lde_b = lde * (int64_t)sizeof(T), ldn_b = ldn * (int64_t)sizeof(T);

  float gm[NCH][EPC], bt[NCH][EPC];
#pragma unroll
  for (int j = 0; j < NCH; ++j)
#pragma unroll
    for (int i = 0; i < EPC; ++i) {
      const int c = (j * 32 + lane) * EPC + i;
      gm[j][i] = gamma ? __ldg(gamma + c) : 1.f;
      bt[j][i] = beta ? __ldg(beta + c) : 0.f;
    }

  int pe = E0;
  auto issue = [&]() {
    if (pe < E1) {
      const uint32_t slot = ring + (uint32_t)(((pe - E0) % kGcSlots) * kSlotBytes);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        gc_cp_async16(slot + j * 512 + lane * 16, hp + (int64_t)pe * ldh_b + j * 512);
        gc_cp_async16(slot + kRow + j * 512 + lane * 16, epp + (int64_t)pe * lde_b + j * 512);
      }
      ++pe;
    }
    gc_commit();
  };
#pragma unroll 1
  for (int i = 0; i < kGcSlots; ++i) issue();

  int cbase = n_lo;
  int cp = __ldg(colptr + min(cbase + lane, n_hi));
  for (int d = n_lo; d < n_hi; ++d) {
    if (d - cbase >= 31) {
      cbase = d;
      cp = __ldg(colptr + min(cbase + lane, n_hi));
    }
    const int e0 = __shfl_sync(0xffffffffu, cp, d - cbase), e1 = __shfl_sync(0xffffffffu, cp, d - cbase + 1);
    float acc[NCH][EPC];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int i = 0; i < EPC; ++i) acc[j][i] = 0.f;
    for (int ei = e0; ei < e1; ++ei) {
      gc_wait<kGcSlots - 1>();
      __syncwarp();
      const uint32_t slot = ring + (uint32_t)(((ei - E0) % kGcSlots) * kSlotBytes);
      float v[NCH][EPC];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        CT::unpack(gc_lds16(slot + j * 512 + lane * 16), v[j]);
#pragma unroll
        for (int i = 0; i < EPC; ++i) s += v[j][i];
      }
      const float mean = warp_sum(s) * (1.0f / (float)C);
      float qq = 0.f;
#pragma unroll
      for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int i = 0; i < EPC; ++i) {
          const float dl = v[j][i] - mean;
          qq += dl * dl;
        }
      const float rstd = rsqrtf(warp_sum(qq) * (1.0f / (float)C) + eps);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float r[EPC], o[EPC];
        CT::unpack(gc_lds16(slot + kRow + j * 512 + lane * 16), r);
#pragma unroll
        for (int i = 0; i < EPC; ++i) o[i] = (v[j][i] - mean) * rstd * gm[j][i] + bt[j][i] + r[i];
        const uint4 packed = CT::pack_round(o);
        *reinterpret_cast<uint4*>(reinterpret_cast<char*>(e_new) + (int64_t)ei * ldn_b + j * 512 + lane * 16) = packed;
#pragma unroll
        for (int i = 0; i < EPC; ++i) acc[j][i] += o[i];
      }
      __syncwarp();
      issue();
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + ((int64_t)d * ldo) * (int64_t)sizeof(T) + j * 512 + lane * 16) = CT::pack_round(acc[j]);
    }
  }
  gc_wait<0>();
}

template <typename T, int NCH>
static int launch_gc_pipe(const void* h, int64_t ldh, const float* gamma, const float* beta, const void* e, int64_t lde, void* e_new, int64_t ldn,
                          const int32_t* colptr, void* out, int64_t ldo, int64_t n_dst, float eps, cudaStream_t s) {
  constexpr int smem = 4 * kGcSlots * 2 * NCH * 512;
  static int blocks_per_sm_dev[kMaxDevices] = {};
  int& blocks_per_sm = blocks_per_sm_dev[current_device()];
  if (blocks_per_sm == 0) {
    cudaError_t err = cudaFuncSetAttribute(graphconv_ln_aggregate_pipe_kernel<T, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(graphconv)");
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, graphconv_ln_aggregate_pipe_kernel<T, NCH>, 128, smem);
    if (err != cudaSuccess || blocks_per_sm < 1) blocks_per_sm = 1;
  }
  int64_t blocks = (int64_t)num_sms() * blocks_per_sm;
  const int64_t max_useful = (n_dst + 3) / 4;
  if (blocks > max_useful) blocks = max_useful;
  graphconv_ln_aggregate_pipe_kernel<T, NCH><<<(unsigned)blocks, 128, smem, s>>>((const T*)h, ldh, gamma, beta, (const T*)e, lde, (T*)e_new, ldn,
                                                                                 colptr, (T*)out, ldo, (int)n_dst, eps);
  return launch_status("graphconv_ln_aggregate_pipe_kernel");
}

// any alignment, C <= 1024: lane-strided columns, fp32 register accumulators
template <typename T>
__global__ void __launch_bounds__(256)
    graphconv_ln_aggregate_generic_kernel(const T* __restrict__ h, int64_t ldh, const float* __restrict__ gamma, const float* __restrict__ beta,
                                          const T* __restrict__ e, int64_t lde, T* __restrict__ e_new, int64_t ldn,
                                          const int32_t* __restrict__ colptr, T* __restrict__ out, int64_t ldo, int64_t n_dst, int C, float eps) {
  constexpr int MAXV = 32;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t d = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < n_dst; d += warps_total) {
    const int e0 = colptr[d], e1 = colptr[d + 1];
    float acc[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) acc[i] = 0.f;
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
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
          const float o = (to_f32<T>(h[(int64_t)ei * ldh + c]) - mean) * rstd * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f) +
                          to_f32<T>(e[(int64_t)ei * lde + c]);
          const T ot = from_f32<T>(o);
          e_new[(int64_t)ei * ldn + c] = ot;
          acc[i] += to_f32<T>(ot);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < C) out[d * ldo + c] = from_f32<T>(acc[i]);
    }
  }
}

template <typename T>
static int launch_gc(const void* h, int64_t ldh, const float* gamma, const float* beta, const void* e, int64_t lde, void* e_new, int64_t ldn,
                     const int32_t* colptr, void* out, int64_t ldo, int64_t n_dst, int C, float eps, bool vec, cudaStream_t s) {
  if (vec) {  // full-warp chunked rows: the pipelined streaming kernel
    constexpr int EPC = 16 / (int)sizeof(T);
    if (C == 32 * EPC) return launch_gc_pipe<T, 1>(h, ldh, gamma, beta, e, lde, e_new, ldn, colptr, out, ldo, n_dst, eps, s);
    if (C == 64 * EPC) return launch_gc_pipe<T, 2>(h, ldh, gamma, beta, e, lde, e_new, ldn, colptr, out, ldo, n_dst, eps, s);
    if (C == 128 * EPC) return launch_gc_pipe<T, 4>(h, ldh, gamma, beta, e, lde, e_new, ldn, colptr, out, ldo, n_dst, eps, s);
  }
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
  ANEMOI_CHECK_ARG(n_dst >= 0 && C >= 1 && C <= 1024, "graphconv_ln_aggregate: need 1 <= C <= 1024");
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
