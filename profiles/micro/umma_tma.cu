// umma_tma.cu — does a concurrent TMA stream into shared memory slow tcgen05.mma down?  One CTA per SM:
//   thread 0   : 128x256x16 UMMAs back to back on RESIDENT operands (same loop as umma_rate.cu)
//   thread 32  : (optional) TMA producer: 2-D bulk tensor loads of 16 KB + 32 KB boxes (the GEMM's per-k-block traffic) from an
//                L2-resident global buffer into a separate 4-slot ring, re-issued as fast as they complete.
// Prints cycles per MMA and the TMA bytes per cycle per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tma umma_tma.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int iters, int with_tma,
                                           int rows_a, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done, bar_ld[4];
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  constexpr int kStage = 49152;
  for (int i = threadIdx.x; i < kStage / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
    for (int s = 0; s < 4; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_ld[s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = make_sw128_desc(base), bd = make_sw128_desc(base + 16384);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma(tmem, ad + 2 * kk, bd + 2 * kk, idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
    mbar_wait(smem_u32(&bar_done), 0);
    long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (threadIdx.x == 32 && with_tma) {
    long long n = 0;
    uint32_t phase[4] = {0, 0, 0, 0};
    int kb = 0;
    const int row0 = (blockIdx.x * 128) % rows_a;
    long long t0 = clock64();
    for (int s = 0; s < 4; ++s) {  // prime 4 slots
      const uint32_t dst = base + kStage * (1 + (s % 3)) , b = smem_u32(&bar_ld[s]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(kStage) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)&tmA), "r"(b), "r"(kb * 64), "r"(row0) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst + 16384), "l"((uint64_t)&tmB), "r"(b), "r"(kb * 64), "r"(0) : "memory");
      kb = (kb + 1) & 7;
    }
    int s = 0;
    while (!stop) {
      const uint32_t dst = base + kStage * (1 + (s % 3)), b = smem_u32(&bar_ld[s]);
      mbar_wait(b, phase[s]);
      phase[s] ^= 1u;
      ++n;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(kStage) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)&tmA), "r"(b), "r"(kb * 64), "r"(row0) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst + 16384), "l"((uint64_t)&tmB), "r"(b), "r"(kb * 64), "r"(0) : "memory");
      kb = (kb + 1) & 7;
      s = (s + 1) & 3;
    }
    long long t1 = clock64();
    for (int q = 0; q < 4; ++q) mbar_wait(smem_u32(&bar_ld[q]), phase[q]);  // drain
    if (blockIdx.x == 0) { out[1] = n * kStage; out[2] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)fn;
  for (int rows_a : {128, 18944}) {  // 128: every CTA re-reads the same A rows (pure L2 hits); 18944: each CTA its own rows (A = 19.4 MB, still L2-resident)
    const int K = 512;
    void *dA, *dB;
    cudaMalloc(&dA, (size_t)rows_a * K * 2);
    cudaMalloc(&dB, (size_t)256 * K * 2);
    cudaMemset(dA, 0, (size_t)rows_a * K * 2);
    cudaMemset(dB, 0, (size_t)256 * K * 2);
    CUtensorMap tmA, tmB;
    cuuint64_t gA[2] = {(cuuint64_t)K, (cuuint64_t)rows_a}, gB[2] = {(cuuint64_t)K, 256}, st[1] = {(cuuint64_t)K * 2};
    cuuint32_t bA[2] = {64, 128}, bB[2] = {64, 256}, es[2] = {1, 1};
    enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gA, st, bA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gB, st, bB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    long long* d;
    cudaMalloc(&d, 24);
    const int smem = 5 * 49152 + 1024 > 232448 ? 232448 : 5 * 49152 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152 + 1024 + 0);
    for (int with_tma : {0, 1}) {
      cudaMemset(d, 0, 24);
      // ring of 3 load slots + 1 resident operand stage = 4 x 48 KB = 192 KB
      k<<<148, 128, 4 * 49152 + 1024>>>(tmA, tmB, 4000, with_tma, rows_a, d);
      long long h[3] = {0, 0, 0};
      cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
      printf("rows_a=%5d tma=%d : %.1f cycles per MMA", rows_a, with_tma, (double)h[0] / 16000.0);
      if (with_tma && h[2]) printf(" | TMA %.1f B/cycle/SM (%.2f TB/s chip at 1.965 GHz)", (double)h[1] / h[2], (double)h[1] / h[2] * 148 * 1.965e9 / 1e12);
      printf("  [%s]\n", cudaGetErrorString(e));
    }
    (void)smem;
    cudaFree(dA), cudaFree(dB), cudaFree(d);
  }
  return 0;
}
