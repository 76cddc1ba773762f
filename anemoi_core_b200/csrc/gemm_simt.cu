// gemm_simt.cu — FFMA GEMM with the shared fused epilogue.  Used for (a) the fp32 parity mode (exact fp32
// products, fp32 accumulation: no TF32 rounding, so the 1e-4 reference tolerance holds through 16 layers) and
// (b) bf16 operands whose shape cannot feed TMA/UMMA (K < 16 or unaligned rows: lin_edge / emb_edges with
// edge_dim = 3..11, in_channels_dst = 12), which are HBM-bound on the [M, N] output anyway.
// 64x64 output tile, BK = 16, 256 threads, 4x4 register tile per thread, smem operands stored k-major.
#include "common.cuh"
#include "gemm.h"

namespace anemoi {

__device__ __forceinline__ void epilogue_store(const EpiParams& ep, int64_t m, int64_t n, float acc) {
  if (ep.ln_stats) {
    const float2 st = ln_row_mean_rstd(ep, m);
    acc = st.y * (acc - st.x * ep.ln_colsum[n]);
  }
  if (ep.bias) acc += ep.bias[n];
  if (ep.g1) acc += load_as_f32(ep.g1, (int64_t)ep.idx1[m] * ep.ldg + n, (ep.flags & ANEMOI_EPI_G1_BF16) ? ANEMOI_BF16 : ANEMOI_F32);
  if (ep.g2) acc += load_as_f32(ep.g2, (int64_t)ep.idx2[m] * ep.ldg + n, (ep.flags & ANEMOI_EPI_G2_BF16) ? ANEMOI_BF16 : ANEMOI_F32);
  if (ep.flags & ANEMOI_EPI_GELU) acc = gelu_erf(acc);
  if (ep.residual) acc += load_as_f32(ep.residual, m * ep.ldr + n, ep.r_dtype);
  store_from_f32(ep.out, m * ep.ldo + n, ep.o_dtype, acc);
}

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ W, int64_t ldw, int64_t K,
                                                        const EpiParams ep) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float sA[BK][BM + 4];
  __shared__ float sW[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each a 4x4 micro-tile
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < K; k0 += BK) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread; consecutive threads read consecutive k (coalesced rows)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int64_t gm = m0 + r, gn = n0 + r, gk = k0 + kk;
      sA[kk][r] = (gm < ep.M && gk < K) ? to_f32<T>(A[gm * lda + gk]) : 0.f;
      sW[kk][r] = (gn < ep.N && gk < K) ? to_f32<T>(W[gn * ldw + gk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sW[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= ep.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n < ep.N) epilogue_store(ep, m, n, acc[i][j]);
    }
  }
}

int linear_simt(const void* A, int64_t lda, const void* W, int64_t ldw, int a_dtype, int64_t K, const EpiParams& ep, cudaStream_t s) {
  dim3 grid((unsigned)((ep.M + 63) / 64), (unsigned)((ep.N + 63) / 64));
  if (grid.y > 65535) {
    set_error("linear(simt): N too large for grid.y");
    return -3;
  }
  if (a_dtype == ANEMOI_BF16)
    gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)A, lda, (const __nv_bfloat16*)W, ldw, K, ep);
  else
    gemm_simt_kernel<float><<<grid, 256, 0, s>>>((const float*)A, lda, (const float*)W, ldw, K, ep);
  int rc = launch_status("gemm_simt_kernel");
  if (rc == 0 && ep.stats_out) rc = launch_partial_row_stats(ep.out, ep.ldo, ep.o_dtype, ep.M, ep.N, ep.stats_out, s);
  return rc;
}

}  // namespace anemoi
