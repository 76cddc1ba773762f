// linear.cu — C-ABI entry for the fused Linear: picks the tcgen05 kernel for bf16 operands that can feed TMA,
// the FFMA kernel otherwise (fp32 parity mode, tiny / unaligned K).  Shape-based choice inside one sm_100a
// library — not a backend switch; there is no CPU or library (cuBLAS) fallback.
#include "common.cuh"
#include "gemm.h"

using namespace anemoi;

extern "C" int anemoi_b200_linear(const void* A, int64_t lda, const void* W, int64_t ldw, int a_dtype, const float* bias, const float* g1,
                                  const int32_t* idx1, const float* g2, const int32_t* idx2, int64_t ldg, const void* residual, int64_t ldr,
                                  int r_dtype, void* out, int64_t ldo, int o_dtype, int64_t M, int64_t N, int64_t K, int flags, const float* ln_stats,
                                  const float* ln_colsum, int64_t ln_parts, int64_t ln_dim, float ln_eps, float* stats_out, void* stream) {
  ANEMOI_CHECK_ARG(M >= 0 && N >= 1 && K >= 1, "linear: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  ANEMOI_CHECK_ARG(a_dtype == ANEMOI_F32 || a_dtype == ANEMOI_BF16, "linear: bad operand dtype %d", a_dtype);
  ANEMOI_CHECK_ARG(o_dtype == ANEMOI_F32 || o_dtype == ANEMOI_BF16, "linear: bad output dtype %d", o_dtype);
  ANEMOI_CHECK_ARG(!residual || r_dtype == ANEMOI_F32 || r_dtype == ANEMOI_BF16, "linear: bad residual dtype %d", r_dtype);
  if (M == 0) return 0;  // an empty edge / node set: nothing to do (and empty tensors carry null pointers)
  ANEMOI_CHECK_ARG(lda >= K && ldw >= K && ldo >= N && (!residual || ldr >= N), "linear: leading dimension too small");
  ANEMOI_CHECK_ARG((g1 == nullptr) == (idx1 == nullptr) && (g2 == nullptr) == (idx2 == nullptr), "linear: gather table without indices");
  ANEMOI_CHECK_ARG((!g1 && !g2) || ldg >= N, "linear: gather leading dimension too small");
  ANEMOI_CHECK_ARG((ln_stats == nullptr) == (ln_colsum == nullptr), "linear: ln_stats and ln_colsum go together");
  ANEMOI_CHECK_ARG(ln_parts >= 0 && ln_parts < 4096 && (ln_parts == 0 || (ln_stats && ln_dim > 0 && ln_eps >= 0.f)),
                   "linear: partial ln_stats need ln_parts > 0, ln_dim > 0, ln_eps >= 0");
  ANEMOI_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "linear: sizes must fit int32");
  if (M == 0) return 0;
  ANEMOI_CHECK_ARG(A && W && out, "linear: null pointer");
  EpiParams ep;
  ep.bias = bias, ep.g1 = g1, ep.idx1 = idx1, ep.g2 = g2, ep.idx2 = idx2, ep.ldg = ldg;
  ep.residual = residual, ep.ldr = ldr, ep.r_dtype = r_dtype;
  ep.out = out, ep.ldo = ldo, ep.o_dtype = o_dtype, ep.M = M, ep.N = N, ep.flags = flags;
  ep.ln_stats = ln_stats, ep.ln_colsum = ln_colsum;
  ep.ln_parts = (int)ln_parts, ep.ln_dim = (int)ln_dim, ep.ln_eps = ln_eps, ep.stats_out = stats_out;
  cudaStream_t s = (cudaStream_t)stream;
  const bool tma_ok = a_dtype == ANEMOI_BF16 && K >= 64 && lda % 8 == 0 && ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  if (tma_ok) return linear_tcgen05(A, lda, W, ldw, K, ep, s);
  return linear_simt(A, lda, W, ldw, a_dtype, K, ep, s);
}
