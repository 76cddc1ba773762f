// gemm.h — internal interface between linear.cu and the two GEMM implementations.
#pragma once
#include "common.cuh"

namespace anemoi {
int linear_simt(const void* A, int64_t lda, const void* W, int64_t ldw, int a_dtype, int64_t K, const EpiParams& ep, cudaStream_t s);
int linear_tcgen05(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t K, const EpiParams& ep, cudaStream_t s);
// stats_out for GEMM paths whose epilogue does not produce it: one pass over the stored output (layernorm.cu)
int launch_partial_row_stats(const void* out, int64_t ldo, int o_dtype, int64_t M, int64_t N, float* stats_out, cudaStream_t s);
}  // namespace anemoi
