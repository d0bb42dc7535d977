// Shared helpers for the HBM-bound kernels (vectorised access, warp/block reductions, grid sizing).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cavp {

constexpr int NUM_SMS = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in every thread; `sh` must hold 32 floats
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? sh[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? sh[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : v * slope;
  return v;
}
// derivative expressed through the OUTPUT z of relu / leaky-relu (slope > 0 keeps the sign)
__device__ __forceinline__ float act_bwd_from_out(float z, int act, float slope) {
  if (act == 1) return z > 0.f ? 1.f : 0.f;
  if (act == 2) return z > 0.f ? 1.f : slope;
  return 1.f;
}

inline int grid_for(long long work_items, int per_block, int max_waves = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(NUM_SMS) * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

#define CAVP_LAUNCH_CHECK() return static_cast<int>(cudaGetLastError())

}  // namespace cavp
