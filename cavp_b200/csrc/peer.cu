// SyncBatchNorm statistics over NVLink peer memory.
//
// Reference: main_vpo_mono.py:130 converts every BatchNorm to torch.nn.SyncBatchNorm, whose forward all-gathers
// [mean, invstd, count] and whose backward all-reduces [sum_dy, sum_dy_xmu] - two small NCCL collectives per layer
// and step, ~120 dependent ones per train step.  Each is a few KB ([2C+1] doubles forward, [2C] floats backward), so
// the cost is pure latency.  Here every rank owns one cudaMalloc'ed exchange buffer that its peers map through CUDA
// IPC; ONE single-CTA kernel per collective PUSHES the local vector into slot [rank] of every peer's buffer (NVLink
// stores), releases a per-peer flag, waits for the flags of the other ranks and sums the slots of its OWN buffer in
// rank order - every rank adds the same numbers in the same order, so the results are bit-identical across ranks.
//
// Buffer of one rank: data[2][P2P_MAX_RANKS][slot_bytes] | flags[2][P2P_MAX_RANKS] (uint32), double-buffered on the
// parity of the call's sequence number.  Slot [parity] is rewritten at call seq + 2; a peer can only get there after it
// has seen this rank's flag of call seq + 1, which this rank releases after it has finished reading call seq.
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "../../include/cavp_b200.h"
#include "common.cuh"

namespace cavp {

constexpr int P2P_MAX_RANKS = 8;
struct PeerBases {
  char* base[P2P_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(512, 1)
    peer_allreduce_kernel(T* __restrict__ io, int n, PeerBases peers, int rank, int world, long long slot_bytes,
                          unsigned seq) {
  const int par = static_cast<int>(seq & 1u);
  const long long flags_off = 2LL * P2P_MAX_RANKS * slot_bytes;
  const long long my_slot = (static_cast<long long>(par) * P2P_MAX_RANKS + rank) * slot_bytes;
  __shared__ int failed;
  if (threadIdx.x == 0) failed = 0;
  // push: the local vector into slot [par][rank] of every rank's buffer (this rank's own included)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const T v = io[i];
    for (int r = 0; r < world; ++r) reinterpret_cast<T*>(peers.base[r] + my_slot)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    const int r = threadIdx.x;
    st_release_sys(reinterpret_cast<unsigned*>(peers.base[r] + flags_off) + par * P2P_MAX_RANKS + rank, seq);
    // wait for rank r's vector of this call to land in this rank's buffer
    const unsigned* f = reinterpret_cast<const unsigned*>(peers.base[rank] + flags_off) + par * P2P_MAX_RANKS + r;
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys(f) - seq) < 0) {
      if (clock64() - t0 > 60000000000LL) {  // ~30 s: a peer died; poison the result instead of hanging the GPU
        failed = 1;
        break;
      }
      __nanosleep(20);
    }
  }
  __syncthreads();
  const char* mine = peers.base[rank] + static_cast<long long>(par) * P2P_MAX_RANKS * slot_bytes;
  const bool bad = failed != 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    T s = 0;
    for (int r = 0; r < world; ++r)
      s += *reinterpret_cast<const volatile T*>(mine + static_cast<long long>(r) * slot_bytes + sizeof(T) * i);
    io[i] = bad ? static_cast<T>(NAN) : s;
  }
}

}  // namespace cavp

using namespace cavp;

static long long peer_buffer_bytes(long long slot_bytes) {
  return 2LL * P2P_MAX_RANKS * slot_bytes + 2LL * P2P_MAX_RANKS * sizeof(unsigned);
}
extern "C" int cavp_peer_alloc(long long slot_bytes, void** buf, unsigned char* handle64) {
  if (!buf || !handle64) return CAVP_ERR_NULL;
  if (slot_bytes <= 0 || (slot_bytes & 15)) return CAVP_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  const long long bytes = peer_buffer_bytes(slot_bytes);
  cudaError_t e = cudaMalloc(buf, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemset(*buf, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *buf);
  if (e != cudaSuccess) {
    cudaFree(*buf);
    *buf = nullptr;
    return static_cast<int>(e);
  }
  std::memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int cavp_peer_open(const unsigned char* handle64, void** buf) {
  if (!buf || !handle64) return CAVP_ERR_NULL;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  return static_cast<int>(cudaIpcOpenMemHandle(buf, h, cudaIpcMemLazyEnablePeerAccess));
}
extern "C" int cavp_peer_close(void* buf, int own) {
  if (!buf) return 0;
  return static_cast<int>(own ? cudaFree(buf) : cudaIpcCloseMemHandle(buf));
}
extern "C" int cavp_peer_allreduce(void* io, int n, int is_f64, void* const* bases, int rank, int world,
                                   long long slot_bytes, int seq_i, void* stream) {
  if (!io || !bases) return CAVP_ERR_NULL;
  if (world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world || n < 0 ||
      static_cast<long long>(n) * (is_f64 ? 8 : 4) > slot_bytes)
    return CAVP_ERR_ARG;
  const unsigned seq = static_cast<unsigned>(seq_i);
  PeerBases pb;
  for (int r = 0; r < P2P_MAX_RANKS; ++r) pb.base[r] = r < world ? static_cast<char*>(bases[r]) : nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f64)
    peer_allreduce_kernel<double><<<1, 512, 0, st>>>(static_cast<double*>(io), n, pb, rank, world, slot_bytes, seq);
  else
    peer_allreduce_kernel<float><<<1, 512, 0, st>>>(static_cast<float*>(io), n, pb, rank, world, slot_bytes, seq);
  CAVP_LAUNCH_CHECK();
}
