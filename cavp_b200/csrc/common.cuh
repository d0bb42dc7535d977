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

// Source index exactly as ATen's area_pixel_compute_source_index (fp32 arithmetic).
struct Lerp { int i0, i1; float w0, w1; };
__device__ __forceinline__ Lerp lerp_index(int dst, int in_size, float scale, int align_corners) {
  float src;
  if (align_corners) {
    src = scale * dst;
  } else {
    src = scale * (dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  Lerp l;
  l.i0 = static_cast<int>(src);
  if (l.i0 > in_size - 1) l.i0 = in_size - 1;
  l.i1 = l.i0 + (l.i0 < in_size - 1 ? 1 : 0);
  l.w1 = src - l.i0;
  l.w0 = 1.f - l.w1;
  return l;
}
__host__ __device__ inline float lerp_scale(int in_size, int out_size, int align_corners) {
  if (align_corners) return out_size > 1 ? static_cast<float>(in_size - 1) / (out_size - 1) : 0.f;
  return static_cast<float>(in_size) / out_size;
}

// One fixed evaluation order (explicit fma / mul), shared by the full-resolution upsample and by the fused
// upsample + argmax of the eval epilogue, so that both see bit-identical logits.
__device__ __forceinline__ float bilerp(const Lerp& ly, const Lerp& lx, float a, float b, float c, float d) {
  const float top = __fmaf_rn(lx.w0, a, __fmul_rn(lx.w1, b));
  const float bot = __fmaf_rn(lx.w0, c, __fmul_rn(lx.w1, d));
  return __fmaf_rn(ly.w0, top, __fmul_rn(ly.w1, bot));
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
