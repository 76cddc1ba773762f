// gemm.h — internal interface between linear.cu and the two GEMM implementations.
#pragma once
#include "common.cuh"

namespace anemoi {
int linear_simt(const void* A, int64_t lda, const void* W, int64_t ldw, int a_dtype, int64_t K, const EpiParams& ep, cudaStream_t s);
int linear_tcgen05(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t K, const EpiParams& ep, cudaStream_t s);
}  // namespace anemoi
