// attention.h - parameter block shared by the attention kernels (attention.cu, attention_mma.cu).
#pragma once
#include "common.cuh"

namespace anemoi {

constexpr int kMaxEdgeDim = 16;

struct AttnParams {
  const void *q, *k, *v, *e, *add;
  void* out;
  int64_t ldq, ldk, ldv, lde_proj, ldadd, ldo;
  const float* edge_attr;  // [E, lde] fp32
  int64_t lde;
  int edge_dim;
  const float* w_edge;  // [H*Ch, ldw_e] fp32 (generic kernel only)
  int64_t ldw_e;
  const float* b_edge;  // [H*Ch] or null
  const void* qw;       // [n_dst, ldqw]: per-head W_e^T q (slab kernel, MODE 2)
  void* abar;           // [n_dst, ldabar]: per-head sum_e alpha_e a_e (slab kernel, MODE 2)
  int64_t ldqw, ldabar;
  int dp;               // per-head stride inside qw / abar rows
  const int32_t* src;
  const int32_t* colptr;
  int64_t n_dst;
  int heads, ch;
  float scale;
  float* lse;  // [n_dst, heads] natural-log softmax normaliser (nullable); pipe / generic kernels only
  int reverse;  // pipe kernel: warps take their dst ranges from the end of the node list (L2 scheduling hint)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// first n in [0, n_dst] with colptr[n] >= x (colptr non-decreasing, colptr[n_dst] = E); warp-wide 32-ary search
__device__ __forceinline__ int colptr_lower_bound(const int32_t* __restrict__ colptr, int n_dst, int x, int lane) {
  int lo = 0, hi = n_dst;  // answer in [lo, hi]
  while (hi - lo > 0) {
    const int span = hi - lo;
    const int step = (span + 31) / 32;
    const int pos = min(lo + lane * step, hi);
    const bool ge = __ldg(colptr + pos) >= x;
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    if (m == 0) {  // all probed positions < x: answer beyond the last probe
      lo = min(lo + 31 * step, hi) + 1;
      if (lo > hi) return hi;  // cannot happen when colptr[hi] >= x
    } else {
      const int f = __ffs(m) - 1;  // first probe with colptr >= x
      hi = min(lo + f * step, hi);
      lo = f == 0 ? hi : min(lo + (f - 1) * step, hi) + 1;
      if (lo > hi) lo = hi;
    }
  }
  return lo;
}

// attention_mma.cu: mma.sync formulation of the folded (MODE 2) bf16 path; returns 1 when the shape is not one it handles
int launch_gt_attention_mma(const AttnParams& p, cudaStream_t s);

}  // namespace anemoi
