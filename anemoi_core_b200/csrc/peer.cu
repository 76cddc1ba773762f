// peer.cu — the exchange step of the dst-range sharded forward over NVLink peer memory (one process per GPU, one 8 x B200 NVSwitch box).
//
// Replaces, for CUDA tensors, the NCCL all-to-all of the reference's halo exchange (distributed/graph.py:466-484 `halo_exchange`,
// layers/block.py:1159-1169) and the all-gather of source rows (`sync_tensor`, graph.py:227-240): every rank WRITES the rows its peers
// need straight into the peers' tables with plain stores on peer-mapped pointers (CUDA IPC), so the whole exchange is three small kernels
// on the compute stream — capturable in the same CUDA graph as the GEMMs and the attention, no host round trip, no NCCL launch:
//     peer_rendezvous : announce "I have started exchange c" to every peer and wait until every peer has (=> each peer has finished
//                       reading what the previous exchange put into the table, so it may be overwritten)
//     halo_push       : one warp per row: rows[send_idx[j]] -> peer table; the last block publishes "my rows of exchange c have arrived"
//     halo_wait       : wait until every peer's rows of exchange c have arrived; the consumer kernel follows in stream order
// All flags are exchange NUMBERS in a per-channel control block (rank-indexed slots, written by exactly one peer each, compared
// wrap-safe), the number itself lives on the device so that a replayed CUDA graph keeps counting.  A peer that never shows up traps
// the waiting kernel after ~10 s instead of hanging the GPU.
//
// The symmetric buffers are cudaMalloc'ed by anemoi_b200_ipc_alloc and mapped into the peers with cudaIpcOpenMemHandle: the ONLY device
// memory this library allocates, through explicit create / destroy entry points (include/anemoi_b200.h).
#include "common.cuh"

namespace anemoi {
namespace {

constexpr int kMaxPeers = 16;
constexpr long long kSpinLimit = 100000000ll;  // x ~100 ns of __nanosleep: ~10 s

struct PeerCtl {  // one per channel and rank, in symmetric memory (256 bytes)
  int started[kMaxPeers];  // [r] = last exchange number rank r has started (written by r)
  int arrived[kMaxPeers];  // [r] = last exchange number whose rows from rank r have landed here (written by r)
  int counter;             // this rank's exchange number (local)
  int done_blocks;         // block ticket of halo_push (local)
};

struct PeerPtrs {
  uint64_t ctl[kMaxPeers];   // peer-mapped address of every rank's control block (own rank: the local one)
  uint64_t dst[kMaxPeers];   // where this rank's first row for peer p goes (peer-mapped), 0 for itself
};

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// All three kernels are PDL launches (common.cuh): they become resident while their predecessor drains and block in griddepcontrol.wait until
// it has completed - stream order, which the protocol relies on ("my consumers of exchange c-1 are done"), is unchanged.
__global__ void peer_rendezvous_kernel(PeerCtl* mine, const PeerPtrs pp, int world, int rank) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x;
  const int c = mine->counter + 1;
  if (lane < world && lane != rank) st_release_sys(&reinterpret_cast<PeerCtl*>(pp.ctl[lane])->started[rank], c);
  if (lane < world && lane != rank) {
    long long spins = 0;
    while (ld_acquire_sys(&mine->started[lane]) - c < 0) {
      __nanosleep(100);
      if (++spins > kSpinLimit) __trap();  // a peer never started this exchange: fail loudly
    }
  }
  __syncwarp();
  if (lane == 0) mine->counter = c;
}

// rows [*, ld_bytes] -> peers.  send_off[world + 1]: rows send_off[p] .. send_off[p+1] of send_idx go to rank p, stored contiguously from pp.dst[p].
__global__ void __launch_bounds__(256) halo_push_kernel(const char* __restrict__ rows, int64_t ld_bytes, int row_bytes, const int32_t* __restrict__ send_idx,
                                                        const int32_t* __restrict__ send_off, int64_t dst_ld_bytes, PeerCtl* mine, const PeerPtrs pp,
                                                        int world, int rank) {
  __shared__ int s_off[kMaxPeers + 1];
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x <= world) s_off[threadIdx.x] = send_off[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n_send = s_off[world];
  const int chunks = row_bytes >> 4;
  for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n_send; j += gridDim.x * wpb) {
    int p = 0;
    while (j >= s_off[p + 1]) ++p;
    const char* src = rows + (int64_t)__ldg(send_idx + j) * ld_bytes;
    char* dst = reinterpret_cast<char*>(pp.dst[p]) + (int64_t)(j - s_off[p]) * dst_ld_bytes;
    for (int c = lane; c < chunks; c += 32) reinterpret_cast<uint4*>(dst)[c] = __ldg(reinterpret_cast<const uint4*>(src) + c);
  }
  __threadfence_system();  // this thread's peer stores are visible system-wide before the block's ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(&mine->done_blocks), 1u);
    if (ticket == gridDim.x - 1) {  // last block: every block's rows are out
      mine->done_blocks = 0;
      __threadfence_system();
      const int c = mine->counter;
      for (int p = 0; p < world; ++p)
        if (p != rank) st_release_sys(&reinterpret_cast<PeerCtl*>(pp.ctl[p])->arrived[rank], c);
    }
  }
}

__global__ void halo_wait_kernel(PeerCtl* mine, int world, int rank) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x;
  const int c = mine->counter;
  if (lane < world && lane != rank) {
    long long spins = 0;
    while (ld_acquire_sys(&mine->arrived[lane]) - c < 0) {
      __nanosleep(100);
      if (++spins > kSpinLimit) __trap();
    }
  }
}

}  // namespace
}  // namespace anemoi

using namespace anemoi;

extern "C" int anemoi_b200_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle64) {
  ANEMOI_CHECK_ARG(bytes > 0 && dev_ptr && handle64, "ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void* p = nullptr;
  ANEMOI_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "ipc_alloc (cudaMemset / cudaIpcGetMemHandle)");
  }
  *dev_ptr = p;
  return 0;
}

extern "C" int anemoi_b200_ipc_open(const void* handle64, void** dev_ptr) {
  ANEMOI_CHECK_ARG(handle64 && dev_ptr, "ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  ANEMOI_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int anemoi_b200_ipc_close(void* dev_ptr) {
  if (dev_ptr) ANEMOI_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

extern "C" int anemoi_b200_ipc_free(void* dev_ptr) {
  if (dev_ptr) ANEMOI_CUDA(cudaFree(dev_ptr));
  return 0;
}

static int fill_ptrs(PeerPtrs& pp, const uint64_t* ctl_ptrs, const uint64_t* dst_ptrs, int64_t world) {
  for (int i = 0; i < kMaxPeers; ++i) pp.ctl[i] = 0, pp.dst[i] = 0;
  for (int i = 0; i < world; ++i) {
    pp.ctl[i] = ctl_ptrs[i];
    if (dst_ptrs) pp.dst[i] = dst_ptrs[i];
  }
  return 0;
}

// ctl_ptrs / dst_ptrs: HOST arrays [world] of peer-mapped device addresses (own rank: local control block; dst may be 0).
extern "C" int anemoi_b200_peer_rendezvous(const uint64_t* ctl_ptrs, int64_t world, int64_t rank, void* stream) {
  ANEMOI_CHECK_ARG(ctl_ptrs && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "peer_rendezvous: bad argument");
  PeerPtrs pp;
  fill_ptrs(pp, ctl_ptrs, nullptr, world);
  cudaError_t le = launch_pdl(peer_rendezvous_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, reinterpret_cast<PeerCtl*>(ctl_ptrs[rank]), pp, (int)world,
                              (int)rank);
  if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(peer_rendezvous_kernel)");
  return launch_status("peer_rendezvous_kernel");
}

extern "C" int anemoi_b200_halo_push(const void* rows, int64_t ld_bytes, int64_t row_bytes, const int32_t* send_idx, const int32_t* send_off,
                                     int64_t n_send, const uint64_t* dst_ptrs, int64_t dst_ld_bytes, const uint64_t* ctl_ptrs, int64_t world,
                                     int64_t rank, void* stream) {
  ANEMOI_CHECK_ARG(ctl_ptrs && dst_ptrs && send_off && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "halo_push: bad argument");
  ANEMOI_CHECK_ARG(n_send == 0 || (rows && send_idx), "halo_push: null rows");
  ANEMOI_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0 && ld_bytes % 16 == 0 && dst_ld_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0,
                   "halo_push: rows must be 16-byte aligned multiples of 16 bytes");
  PeerPtrs pp;
  fill_ptrs(pp, ctl_ptrs, dst_ptrs, world);
  int64_t blocks = (n_send + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;  // the signalling block runs even when nothing is sent
  cudaError_t le = launch_pdl(halo_push_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const char*>(rows), ld_bytes,
                              (int)row_bytes, send_idx, send_off, dst_ld_bytes, reinterpret_cast<PeerCtl*>(ctl_ptrs[rank]), pp, (int)world, (int)rank);
  if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(halo_push_kernel)");
  return launch_status("halo_push_kernel");
}

extern "C" int anemoi_b200_halo_wait(const uint64_t* ctl_ptrs, int64_t world, int64_t rank, void* stream) {
  ANEMOI_CHECK_ARG(ctl_ptrs && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "halo_wait: bad argument");
  cudaError_t le = launch_pdl(halo_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, reinterpret_cast<PeerCtl*>(ctl_ptrs[rank]), (int)world, (int)rank);
  if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(halo_wait_kernel)");
  return launch_status("halo_wait_kernel");
}
