// HBM-bound kernels around the tensor-core tiles: layout changes, BatchNorm (train/eval, fwd/bwd), pooling,
// bilinear resampling.  All tensors are NHWC fp32 with an explicit pixel stride (ld) so channel slices work in place.
// Each extern "C" entry cites the reference op it stands in for (paths relative to the reference repo).
#include <cuda_bf16.h>
#include "common.cuh"
#include "../../include/cavp_b200.h"

namespace cavp {

// ------------------------------------------------------------------------------------------------ layout
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int c, int hw,
                                    int cpad) {
  const long long total = static_cast<long long>(n) * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long img = i / hw;
    const long long pix = i - img * hw;
    const float* s = src + img * c * hw + pix;
    float* d = dst + i * cpad;
    for (int ch = 0; ch < cpad; ++ch) d[ch] = ch < c ? s[static_cast<long long>(ch) * hw] : 0.f;
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int c, int hw,
                                    int ld) {
  const long long total = static_cast<long long>(n) * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long img = i / hw;
    const long long pix = i - img * hw;
    const float* s = src + i * ld;
    float* d = dst + img * c * hw + pix;
    for (int ch = 0; ch < c; ++ch) d[static_cast<long long>(ch) * hw] = s[ch];
  }
}
// dst[b][cidx][r] = src[b][r][cidx] with leading dimensions; 32x32 tiles through shared memory
// dst[b][j][i] = src[b][i][j]; with `lo` the transposed matrix is written as its TF32 split (hi -> dst, lo -> lo): the
// data-gradient GEMM's weight operand in one pass instead of transpose + split
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ lo,
                                 int rows, int cols, long long src_ld, long long dst_ld, long long src_bs,
                                 long long dst_bs) {
  __shared__ float tile[32][33];
  const float* s = src + blockIdx.z * src_bs;
  float* d = dst + blockIdx.z * dst_bs;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, cc = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && cc < cols) ? s[r * src_ld + cc] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int cc = c0 + j, r = r0 + threadIdx.x;
    if (cc < cols && r < rows) {
      const float v = tile[threadIdx.x][j];
      if (lo) {
        const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
        d[cc * dst_ld + r] = h;
        lo[blockIdx.z * dst_bs + cc * dst_ld + r] = __uint_as_float((__float_as_uint(v - h) + 0x1000u) & 0xFFFFE000u);
      } else {
        d[cc * dst_ld + r] = v;
      }
    }
  }
}
// The same transpose + TF32 split for EVERY dgrad weight operand of a model in one launch (the per-layer launches were
// launch-bound: ~140 kernels of a few microseconds per step).  Table rows mirror the int64[9] the host packs
// (cavp_b200/engine.py:WeightSplitCache); a work item is one 32 x 32 tile of one batch (= filter tap) of one tensor.
struct TsTensor {
  const float* src;
  float* hi;
  float* lo;
  int rows, cols;
  long long src_ld, dst_ld, src_bs, dst_bs;
  int tiles_c, tiles_r;
};
static_assert(sizeof(TsTensor) == 72, "table row layout");

template <bool BF16>
__global__ void __launch_bounds__(256)
transpose_split_multi_kernel(const TsTensor* __restrict__ table, const int2* __restrict__ work, int nwork) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int2 w = work[wi];
    const TsTensor t = table[w.x];
    const int per = t.tiles_c * t.tiles_r;
    const int z = w.y / per, rem = w.y - z * per;
    const int c0 = (rem % t.tiles_c) * 32, r0 = (rem / t.tiles_c) * 32;
    const float* s = t.src + z * t.src_bs;
    __syncthreads();  // the previous item's reads of `tile` are done
    for (int j = ty; j < 32; j += 8) {
      const int r = r0 + j, cc = c0 + tx;
      tile[j][tx] = (r < t.rows && cc < t.cols) ? __ldg(s + r * t.src_ld + cc) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int cc = c0 + j, r = r0 + tx;
      if (cc < t.cols && r < t.rows) {
        const float v = tile[tx][j];
        const long long o = z * t.dst_bs + cc * t.dst_ld + r;
        if (BF16) {  // `hi` is a bf16 destination (the bf16 dgrad operand of the configs[2] path)
          reinterpret_cast<__nv_bfloat16*>(t.hi)[o] = __float2bfloat16_rn(v);
        } else {
          const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
          t.hi[o] = h;
          t.lo[o] = __uint_as_float((__float_as_uint(v - h) + 0x1000u) & 0xFFFFE000u);
        }
      }
    }
  }
}
__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n4, float alpha) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += alpha * b.x; a.y += alpha * b.y; a.z += alpha * b.z; a.w += alpha * b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}
__global__ void fill_strided_kernel(float* __restrict__ dst, int ld, long long rows, int C, float value) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C;
    dst[r * ld + (i - r * C)] = value;
  }
}
// rows gathered: dst[i][:] = src[idx[i]][:]   (forward_audio feature shuffle, models/cavp_model.py:171-173)
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx,
                                   float* __restrict__ dst, int nrows, int c, int accumulate_scatter) {
  const int r = blockIdx.x;
  if (r >= nrows) return;
  const long long sidx = idx[r];
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    if (accumulate_scatter)
      atomicAdd(dst + sidx * c + ch, src[static_cast<long long>(r) * c + ch]);
    else
      dst[static_cast<long long>(r) * c + ch] = src[sidx * c + ch];
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm
// Reduce the per-(m_tile, quarter) partial sums written by the igemm epilogue (fp64 accumulate) and finish the batch
// statistics.  block = (32 channels, 32 part-lanes): reads are coalesced across channels.
__device__ __forceinline__ float tf32_rn_dev(float x) {  // = ptx.cuh:tf32_rn (round to nearest, ties away)
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// one channel: statistics -> mean / invstd / scale / shift, running-stat update (SyncBatchNorm modes: see the header)
__device__ __forceinline__ void bn_finish_channel(int ch, int C, double s1, double s2, double count,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  float* __restrict__ running_mean, float* __restrict__ running_var,
                                                  float momentum, float eps, float* __restrict__ mean_out,
                                                  float* __restrict__ invstd_out, float* __restrict__ scale_out,
                                                  float* __restrict__ shift_out, double* __restrict__ sums_io,
                                                  int sums_mode, long long* __restrict__ num_batches_tracked) {
  if (sums_mode == 1) {  // only export local sums + the local count (SyncBatchNorm: all-reduced, then mode 2)
    sums_io[ch] = s1;
    sums_io[C + ch] = s2;
    if (ch == 0) sums_io[2 * C] = count;
    return;
  }
  if (sums_mode == 2) {  // global sums and the GLOBAL count come from the all-reduced buffer: no host round trip
    s1 = sums_io[ch];
    s2 = sums_io[C + ch];
    count = sums_io[2 * C];
  }
  if (ch == 0 && num_batches_tracked) num_batches_tracked[0] += 1;  // nn.BatchNorm2d bookkeeping, same launch
  const double mean = s1 / count;
  double var = s2 / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = 1.0f / sqrtf(static_cast<float>(var) + eps);
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  mean_out[ch] = static_cast<float>(mean);
  invstd_out[ch] = invstd;
  scale_out[ch] = g * invstd;
  shift_out[ch] = b - static_cast<float>(mean) * g * invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * static_cast<float>(mean);
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * static_cast<float>(unbiased);
  }
}

template <typename PT>
__global__ void bn_finalize_kernel(const PT* __restrict__ partials, int nparts, int ldstat, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   float* __restrict__ scale_out, float* __restrict__ shift_out,
                                   double* __restrict__ sums_io, int sums_mode,
                                   long long* __restrict__ num_batches_tracked) {
  __shared__ double sh1[32][33], sh2[32][33];
  const int ch = blockIdx.x * 32 + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (sums_mode != 2 && ch < C) {
    // four independent chains: the loop is latency-bound (thousands of partial rows per channel at 112x112), so keep
    // eight loads in flight per thread instead of two
    double t1[4] = {0.0, 0.0, 0.0, 0.0}, t2[4] = {0.0, 0.0, 0.0, 0.0};
    int i = threadIdx.y;
    for (; i + 96 < nparts; i += 128) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const PT* pp = partials + static_cast<size_t>(i + 32 * u) * 2 * ldstat;
        t1[u] += pp[ch];
        t2[u] += pp[ldstat + ch];
      }
    }
    for (; i < nparts; i += 32) {
      const PT* pp = partials + static_cast<size_t>(i) * 2 * ldstat;
      t1[0] += pp[ch];
      t2[0] += pp[ldstat + ch];
    }
    s1 = (t1[0] + t1[1]) + (t1[2] + t1[3]);
    s2 = (t2[0] + t2[1]) + (t2[2] + t2[3]);
  }
  sh1[threadIdx.y][threadIdx.x] = s1;
  sh2[threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && ch < C) {
    for (int j = 1; j < 32; ++j) {
      s1 += sh1[j][threadIdx.x];
      s2 += sh2[j][threadIdx.x];
    }
    bn_finish_channel(ch, C, s1, s2, count, gamma, beta, running_mean, running_var, momentum, eps, mean_out, invstd_out,
                      scale_out, shift_out, sums_io, sums_mode, num_batches_tracked);
  }
}

// Two-stage finalize in ONE launch: block (x = 32 channels, y = slice) reduces its slice of the partial rows to
// ws[slice][2][C]; the LAST slice block of a channel group (atomic ticket, counter reset for the next launch) adds the S
// slice rows in a fixed order and finishes the channels.  Same arithmetic and order as partial_reduce + finalize<double>.
__global__ void bn_reduce_finalize_kernel(const float* __restrict__ partials, int nparts, int ldstat, int C, double count,
                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                          float* __restrict__ running_mean, float* __restrict__ running_var,
                                          float momentum, float eps, float* __restrict__ mean_out,
                                          float* __restrict__ invstd_out, float* __restrict__ scale_out,
                                          float* __restrict__ shift_out, double* __restrict__ sums_io, int sums_mode,
                                          long long* __restrict__ num_batches_tracked, double* __restrict__ ws,
                                          int* __restrict__ counters) {
  __shared__ double sh1[32][33], sh2[32][33];
  __shared__ int is_last;
  const int ch = blockIdx.x * 32 + threadIdx.x;
  const int S = gridDim.y, sl = blockIdx.y;
  const int per = (nparts + S - 1) / S;
  const int beg = sl * per, end = min(nparts, beg + per);
  double t1[4] = {0.0, 0.0, 0.0, 0.0}, t2[4] = {0.0, 0.0, 0.0, 0.0};
  if (ch < C) {
    int i = beg + threadIdx.y;
    for (; i + 96 < end; i += 128) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* pp = partials + static_cast<size_t>(i + 32 * u) * 2 * ldstat;
        t1[u] += pp[ch];
        t2[u] += pp[ldstat + ch];
      }
    }
    for (; i < end; i += 32) {
      const float* pp = partials + static_cast<size_t>(i) * 2 * ldstat;
      t1[0] += pp[ch];
      t2[0] += pp[ldstat + ch];
    }
  }
  sh1[threadIdx.y][threadIdx.x] = (t1[0] + t1[1]) + (t1[2] + t1[3]);
  sh2[threadIdx.y][threadIdx.x] = (t2[0] + t2[1]) + (t2[2] + t2[3]);
  __syncthreads();
  if (threadIdx.y == 0 && ch < C) {
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < 32; ++j) {
      s1 += sh1[j][threadIdx.x];
      s2 += sh2[j][threadIdx.x];
    }
    ws[(static_cast<size_t>(sl) * 2 + 0) * C + ch] = s1;
    ws[(static_cast<size_t>(sl) * 2 + 1) * C + ch] = s2;
    __threadfence();  // this slice row is visible device-wide before the ticket is taken
  }
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int ticket = atomicAdd(&counters[blockIdx.x], 1);
    is_last = ticket == S - 1;
    if (is_last) counters[blockIdx.x] = 0;  // ready for the next launch (launches of a stream are ordered)
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.y == 0 && ch < C) {
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < S; ++j) {  // fixed order, whichever block arrives last
      s1 += __ldcg(ws + (static_cast<size_t>(j) * 2 + 0) * C + ch);
      s2 += __ldcg(ws + (static_cast<size_t>(j) * 2 + 1) * C + ch);
    }
    bn_finish_channel(ch, C, s1, s2, count, gamma, beta, running_mean, running_var, momentum, eps, mean_out, invstd_out,
                      scale_out, shift_out, sums_io, sums_mode, num_batches_tracked);
  }
}
// eval mode: scale/shift from running statistics
__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  const float invstd = 1.0f / sqrtf(rv[ch] + eps);
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  scale[ch] = g * invstd;
  shift[ch] = b - rm[ch] * g * invstd;
}

// (flat grid-stride layout: measured faster here than the column-stationary one used by bn_bwd_apply - only two of the
// four loads are per-channel constants)
__global__ void bn_apply_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ scale,
                                const float* __restrict__ shift, const float* __restrict__ res, int ldr,
                                float* __restrict__ out, int ldo, long long rows, int C4, int act, float slope) {
  const long long total = rows * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C4;
    const int c4 = static_cast<int>(i - r * C4);
    const float4 v = *reinterpret_cast<const float4*>(y + r * ldy + c4 * 4);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c4 * 4);
    const float4 sh = *reinterpret_cast<const float4*>(shift + c4 * 4);
    float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    if (res) {
      const float4 rr = *reinterpret_cast<const float4*>(res + r * ldr + c4 * 4);
      o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
    }
    o.x = act_fwd(o.x, act, slope); o.y = act_fwd(o.y, act, slope);
    o.z = act_fwd(o.z, act, slope); o.w = act_fwd(o.w, act, slope);
    *reinterpret_cast<float4*>(out + r * ldo + c4 * 4) = o;
  }
}

// Column reductions of g = dz * act'(z) and g * xhat over a block of rows.  Also used as "activation backward +
// bias gradient" (y == nullptr) and as a plain column sum (z == nullptr).  grid = (ceil(C4/64), nblk), block = 256.
__global__ void colreduce_kernel(const float* __restrict__ dz, int lddz, const float* __restrict__ z, int ldz,
                                 const float* __restrict__ y, int ldy, const float* __restrict__ mean,
                                 const float* __restrict__ invstd, long long rows, int C4, int act, float slope,
                                 float* __restrict__ gout, int ldg, float* __restrict__ partials, int ldp,
                                 const float* __restrict__ zscale, const float* __restrict__ zshift, int LC) {
  // LC (16 / 32 / 64) float4 column lanes per block, 256 / LC row lanes: with a fixed 64 x 4 layout the C = 64 layers
  // (the largest maps of the network) kept 3 of 4 threads idle
  __shared__ float4 sh[2][256];
  const int cl = threadIdx.x & (LC - 1), rl = threadIdx.x / LC;
  const int nrl = 256 / LC;
  const int c4 = blockIdx.x * LC + cl;
  const long long rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const long long rbeg = blockIdx.y * rows_per;
  const long long rend = rbeg + rows_per < rows ? rbeg + rows_per : rows;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (c4 < C4) {
    float4 mu = s1, is = s1, zs = s1, zb = s1;
    if (y) {
      mu = *reinterpret_cast<const float4*>(mean + c4 * 4);
      is = *reinterpret_cast<const float4*>(invstd + c4 * 4);
    }
    // zscale given (and z == NULL): the activation input is recomputed as fma(y, scale, shift) - the expression of
    // bn_apply_kernel - instead of reading the activation output z back (one tensor read less)
    const bool remask = !z && zscale && y && act != 0;
    if (remask) {
      zs = *reinterpret_cast<const float4*>(zscale + c4 * 4);
      zb = *reinterpret_cast<const float4*>(zshift + c4 * 4);
    }
#pragma unroll 2
    for (long long r = rbeg + rl; r < rend; r += nrl) {
      float4 g = *reinterpret_cast<const float4*>(dz + r * lddz + c4 * 4);
      float4 yy = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y) yy = *reinterpret_cast<const float4*>(y + r * ldy + c4 * 4);
      if (z) {
        const float4 zz = *reinterpret_cast<const float4*>(z + r * ldz + c4 * 4);
        g.x *= act_bwd_from_out(zz.x, act, slope); g.y *= act_bwd_from_out(zz.y, act, slope);
        g.z *= act_bwd_from_out(zz.z, act, slope); g.w *= act_bwd_from_out(zz.w, act, slope);
      } else if (remask) {
        g.x *= act_bwd_from_out(fmaf(yy.x, zs.x, zb.x), act, slope); g.y *= act_bwd_from_out(fmaf(yy.y, zs.y, zb.y), act, slope);
        g.z *= act_bwd_from_out(fmaf(yy.z, zs.z, zb.z), act, slope); g.w *= act_bwd_from_out(fmaf(yy.w, zs.w, zb.w), act, slope);
      }
      if (gout) *reinterpret_cast<float4*>(gout + r * ldg + c4 * 4) = g;
      s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
      if (y) {
        s2.x += g.x * (yy.x - mu.x) * is.x; s2.y += g.y * (yy.y - mu.y) * is.y;
        s2.z += g.z * (yy.z - mu.z) * is.z; s2.w += g.w * (yy.w - mu.w) * is.w;
      }
    }
  }
  sh[0][rl * LC + cl] = s1;
  sh[1][rl * LC + cl] = s2;
  __syncthreads();
  if (rl == 0 && c4 < C4) {
    for (int k = 1; k < nrl; ++k) {
      const float4 a = sh[0][k * LC + cl], b = sh[1][k * LC + cl];
      s1.x += a.x; s1.y += a.y; s1.z += a.z; s1.w += a.w;
      s2.x += b.x; s2.y += b.y; s2.z += b.z; s2.w += b.w;
    }
    float* pp = partials + static_cast<size_t>(blockIdx.y) * 2 * ldp;
    *reinterpret_cast<float4*>(pp + c4 * 4) = s1;
    *reinterpret_cast<float4*>(pp + ldp + c4 * 4) = s2;
  }
}
// out[k][c] = sum_i partials[i][k][c]  (k = 0..nk-1), fp64 accumulate.  block = (32 channels, 32 part-lanes)
__global__ void partials_sum_kernel(const float* __restrict__ partials, int nparts, int ldp, int C, int nk,
                                    float* __restrict__ out) {
  __shared__ double sh[32][33];
  const int i = blockIdx.x * 32 + threadIdx.x;  // flat (k, channel)
  double s = 0.0;
  if (i < nk * C) {
    const int k = i / C, ch = i - k * C;
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    int pidx = threadIdx.y;
    for (; pidx + 96 < nparts; pidx += 128) {
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] += partials[(static_cast<size_t>(pidx + 32 * u) * nk + k) * ldp + ch];
    }
    for (; pidx < nparts; pidx += 32) t[0] += partials[(static_cast<size_t>(pidx) * nk + k) * ldp + ch];
    s = (t[0] + t[1]) + (t[2] + t[3]);
  }
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < nk * C) {
    for (int j = 1; j < 32; ++j) s += sh[j][threadIdx.x];
    out[i] = static_cast<float>(s);
  }
}

// Column-stationary layout (as colreduce): a thread keeps its four channels' seven coefficient vectors in registers and
// walks rows, instead of re-loading them for every float4 of a flat grid-stride loop (7 of 10 loads were constants).
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dz, int lddz, const float* __restrict__ z, int ldz,
                    const float* __restrict__ y, int ldy, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ sums, float inv_count, long long rows, int C4, int act,
                    float slope, float* __restrict__ dy, int lddy, float* __restrict__ dres,
                    int lddres, const float* __restrict__ zscale, const float* __restrict__ zshift, int LC,
                    const double* __restrict__ count_dev, float* __restrict__ dy_hi, float* __restrict__ dy_lo,
                    int split_bf16) {
  if (count_dev) inv_count = static_cast<float>(1.0 / count_dev[0]);  // SyncBatchNorm: the all-reduced global count
  const int cl = threadIdx.x & (LC - 1), rl = threadIdx.x / LC;
  const int nrl = 256 / LC;
  const int c4 = blockIdx.x * LC + cl;
  if (c4 >= C4) return;
  const int C = C4 * 4;
  const int c = c4 * 4;
  const bool remask = !z && zscale && act != 0;
  const long long rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const long long rbeg = blockIdx.y * rows_per;
  const long long rend = rbeg + rows_per < rows ? rbeg + rows_per : rows;
  const float4 mu = *reinterpret_cast<const float4*>(mean + c);
  const float4 is = *reinterpret_cast<const float4*>(invstd + c);
  const float4 ga = gamma ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 sg = *reinterpret_cast<const float4*>(sums + c);
  const float4 sgx = *reinterpret_cast<const float4*>(sums + C + c);
  float4 zs = make_float4(0.f, 0.f, 0.f, 0.f), zb = zs;
  if (remask) {
    zs = *reinterpret_cast<const float4*>(zscale + c);
    zb = *reinterpret_cast<const float4*>(zshift + c);
  }
#pragma unroll 2
  for (long long r = rbeg + rl; r < rend; r += nrl) {
    float4 g = *reinterpret_cast<const float4*>(dz + r * lddz + c);
    const float4 yy = *reinterpret_cast<const float4*>(y + r * ldy + c);
    if (z) {
      const float4 zz = *reinterpret_cast<const float4*>(z + r * ldz + c);
      g.x *= act_bwd_from_out(zz.x, act, slope); g.y *= act_bwd_from_out(zz.y, act, slope);
      g.z *= act_bwd_from_out(zz.z, act, slope); g.w *= act_bwd_from_out(zz.w, act, slope);
    } else if (remask) {
      g.x *= act_bwd_from_out(fmaf(yy.x, zs.x, zb.x), act, slope); g.y *= act_bwd_from_out(fmaf(yy.y, zs.y, zb.y), act, slope);
      g.z *= act_bwd_from_out(fmaf(yy.z, zs.z, zb.z), act, slope); g.w *= act_bwd_from_out(fmaf(yy.w, zs.w, zb.w), act, slope);
    }
    if (dres) *reinterpret_cast<float4*>(dres + r * lddres + c) = g;
    float4 o;
    o.x = ga.x * is.x * (g.x - sg.x * inv_count - (yy.x - mu.x) * is.x * sgx.x * inv_count);
    o.y = ga.y * is.y * (g.y - sg.y * inv_count - (yy.y - mu.y) * is.y * sgx.y * inv_count);
    o.z = ga.z * is.z * (g.z - sg.z * inv_count - (yy.z - mu.z) * is.z * sgx.z * inv_count);
    o.w = ga.w * is.w * (g.w - sg.w * inv_count - (yy.w - mu.w) * is.w * sgx.w * inv_count);
    *reinterpret_cast<float4*>(dy + r * lddy + c) = o;
    if (dy_hi && split_bf16) {  // bf16 row: the dense bf16 copy of dy the bf16 weight-gradient kernel fetches by TMA
      __nv_bfloat162 a = __floats2bfloat162_rn(o.x, o.y), b = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dy_hi) + r * C + c) = pk;
    } else if (dy_hi) {
      // the TF32 hi | lo split of dy, dense [rows][C]: the TMA-fed weight-gradient kernels read it directly, so the
      // separate split pass (one more read of dy) disappears
      const float h0 = tf32_rn_dev(o.x), h1 = tf32_rn_dev(o.y), h2 = tf32_rn_dev(o.z), h3 = tf32_rn_dev(o.w);
      *reinterpret_cast<float4*>(dy_hi + r * C + c) = make_float4(h0, h1, h2, h3);
      if (dy_lo)
        *reinterpret_cast<float4*>(dy_lo + r * C + c) =
            make_float4(tf32_rn_dev(o.x - h0), tf32_rn_dev(o.y - h1), tf32_rn_dev(o.z - h2), tf32_rn_dev(o.w - h3));
    }
  }
}

// ------------------------------------------------------------------------------------------------ pooling
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy,
                                   int* __restrict__ idx, int n, int h, int w, int C4, int k, int stride, int pad,
                                   int ho, int wo) {
  const long long total = static_cast<long long>(n) * ho * wo * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / C4;
    const int c = static_cast<int>(i - pix * C4) * 4;
    const int ox = static_cast<int>(pix % wo);
    const int oy = static_cast<int>((pix / wo) % ho);
    const int img = static_cast<int>(pix / (static_cast<long long>(wo) * ho));
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int4 bi = make_int4(-1, -1, -1, -1);
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= w) continue;
        const int ip = (img * h + iy) * w + ix;
        const float4 v = *reinterpret_cast<const float4*>(x + static_cast<long long>(ip) * ldx + c);
        if (v.x > best.x || bi.x < 0) { best.x = v.x; bi.x = ip; }
        if (v.y > best.y || bi.y < 0) { best.y = v.y; bi.y = ip; }
        if (v.z > best.z || bi.z < 0) { best.z = v.z; bi.z = ip; }
        if (v.w > best.w || bi.w < 0) { best.w = v.w; bi.w = ip; }
      }
    }
    *reinterpret_cast<float4*>(y + pix * ldy + c) = best;
    if (idx) *reinterpret_cast<int4*>(idx + pix * (C4 * 4) + c) = bi;
  }
}
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, int lddy, const int* __restrict__ idx,
                                   float* __restrict__ dx, int lddx, long long opix, int C) {
  const long long total = opix * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / C;
    const int c = static_cast<int>(i - pix * C);
    const int ip = idx[i];
    if (ip >= 0) atomicAdd(dx + static_cast<long long>(ip) * lddx + c, dy[pix * lddy + c]);
  }
}
// out[img][c] = scale * sum_p x[img][p][c];  grid = (ceil(C/128), n), block = (128 channels? no: 32 x 8)
__global__ void pixel_sum_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int hw, int C,
                                 float scale) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int img = blockIdx.y;
  float s = 0.f;
  if (c < C)
    for (int pidx = threadIdx.y; pidx < hw; pidx += 8) s += x[(static_cast<long long>(img) * hw + pidx) * ldx + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int k = 1; k < 8; ++k) s += sh[k][threadIdx.x];
    out[static_cast<long long>(img) * C + c] = s * scale;
  }
}
// global max over pixels with argmax (AdaptiveMaxPool2d((1,1)), models/audio/audio_network.py:24)
__global__ void pixel_max_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int* __restrict__ arg,
                                 int hw, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int img = blockIdx.y;
  if (c >= C) return;
  float best = -INFINITY;
  int bi = 0;
  for (int pidx = 0; pidx < hw; ++pidx) {
    const float v = x[(static_cast<long long>(img) * hw + pidx) * ldx + c];
    if (v > best || pidx == 0) { best = v; bi = pidx; }
  }
  out[static_cast<long long>(img) * C + c] = best;
  arg[static_cast<long long>(img) * C + c] = bi;
}
__global__ void pixel_max_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg,
                                     float* __restrict__ dx, int lddx, int hw, int C, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * C) return;
  const int img = i / C, c = i - img * C;
  dx[(static_cast<long long>(img) * hw + arg[i]) * lddx + c] = dout[i];
}
// dx[img][p][c] (+)= scale * dout[img][c]
__global__ void pixel_bcast_kernel(const float* __restrict__ dout, float* __restrict__ dx, int lddx, int hw, int C4,
                                   long long rows, float scale, int accumulate) {
  const long long total = rows * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C4;
    const int c = static_cast<int>(i - r * C4) * 4;
    const long long img = r / hw;
    const float4 g = *reinterpret_cast<const float4*>(dout + img * (C4 * 4) + c);
    float4* d = reinterpret_cast<float4*>(dx + r * lddx + c);
    float4 o = make_float4(g.x * scale, g.y * scale, g.z * scale, g.w * scale);
    if (accumulate) {
      const float4 a = *d;
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    *d = o;
  }
}

// ------------------------------------------------------------------------------------------------ bilinear
__global__ void bilinear_fwd_nhwc_kernel(const float* __restrict__ x, int ldx, int hin, int win,
                                         float* __restrict__ y, int ldy, int hout, int wout, int n, int C4,
                                         int align, float sh, float sw) {
  const long long total = static_cast<long long>(n) * hout * wout * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i / C4;
    const int c = static_cast<int>(i - pix * C4) * 4;
    const int ox = static_cast<int>(pix % wout);
    const int oy = static_cast<int>((pix / wout) % hout);
    const long long img = pix / (static_cast<long long>(wout) * hout);
    const Lerp ly = lerp_index(oy, hin, sh, align), lx = lerp_index(ox, win, sw, align);
    const float* base = x + img * hin * win * ldx + c;
    const float4 a = *reinterpret_cast<const float4*>(base + (static_cast<long long>(ly.i0) * win + lx.i0) * ldx);
    const float4 b = *reinterpret_cast<const float4*>(base + (static_cast<long long>(ly.i0) * win + lx.i1) * ldx);
    const float4 cc = *reinterpret_cast<const float4*>(base + (static_cast<long long>(ly.i1) * win + lx.i0) * ldx);
    const float4 d = *reinterpret_cast<const float4*>(base + (static_cast<long long>(ly.i1) * win + lx.i1) * ldx);
    float4 o;
    o.x = ly.w0 * (lx.w0 * a.x + lx.w1 * b.x) + ly.w1 * (lx.w0 * cc.x + lx.w1 * d.x);
    o.y = ly.w0 * (lx.w0 * a.y + lx.w1 * b.y) + ly.w1 * (lx.w0 * cc.y + lx.w1 * d.y);
    o.z = ly.w0 * (lx.w0 * a.z + lx.w1 * b.z) + ly.w1 * (lx.w0 * cc.z + lx.w1 * d.z);
    o.w = ly.w0 * (lx.w0 * a.w + lx.w1 * b.w) + ly.w1 * (lx.w0 * cc.w + lx.w1 * d.w);
    *reinterpret_cast<float4*>(y + pix * ldy + c) = o;
  }
}
// NHWC (low-res logits) -> NCHW full-resolution prediction; thread per output pixel, loop over classes
__global__ void bilinear_fwd_nchw_kernel(const float* __restrict__ x, int ldx, int hin, int win,
                                         float* __restrict__ y, int hout, int wout, int n, int C, int align, float sh,
                                         float sw) {
  const long long total = static_cast<long long>(n) * hout * wout;
  const long long plane = static_cast<long long>(hout) * wout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % wout);
    const int oy = static_cast<int>((i / wout) % hout);
    const long long img = i / plane;
    const Lerp ly = lerp_index(oy, hin, sh, align), lx = lerp_index(ox, win, sw, align);
    const float* base = x + img * hin * win * ldx;
    const float* pa = base + (static_cast<long long>(ly.i0) * win + lx.i0) * ldx;
    const float* pb = base + (static_cast<long long>(ly.i0) * win + lx.i1) * ldx;
    const float* pc = base + (static_cast<long long>(ly.i1) * win + lx.i0) * ldx;
    const float* pd = base + (static_cast<long long>(ly.i1) * win + lx.i1) * ldx;
    float* o = y + img * C * plane + static_cast<long long>(oy) * wout + ox;
    for (int ch = 0; ch < C; ++ch)
      o[ch * plane] = bilerp(ly, lx, pa[ch], pb[ch], pc[ch], pd[ch]);
  }
}
// Gather-form backward (deterministic, no atomics): every low-res element collects from the output pixels whose
// 2x2 footprint touches it.  dy is NHWC (ld) or NCHW (nchw=1).  Images >= n_valid have an all-zero gradient.
__global__ void bilinear_bwd_kernel(const float* __restrict__ dy, int lddy, int hout, int wout,
                                    float* __restrict__ dx, int lddx, int hin, int win, int n, int C, int align,
                                    float sh, float sw, int nchw, int n_valid) {
  const long long total = static_cast<long long>(n) * hin * win * C;
  const long long plane = static_cast<long long>(hout) * wout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int ch, ix, iy;
    long long img;
    if (nchw) {  // threads vary over x fastest for the NCHW source
      ix = static_cast<int>(i % win);
      iy = static_cast<int>((i / win) % hin);
      ch = static_cast<int>((i / (static_cast<long long>(win) * hin)) % C);
      img = i / (static_cast<long long>(win) * hin * C);
    } else {
      ch = static_cast<int>(i % C);
      ix = static_cast<int>((i / C) % win);
      iy = static_cast<int>((i / (static_cast<long long>(C) * win)) % hin);
      img = i / (static_cast<long long>(C) * win * hin);
    }
    float acc = 0.f;
    if (img < n_valid) {
      // candidate output rows / cols: source coordinate within (i-1, i+1)
      int y_lo, y_hi, x_lo, x_hi;
      if (align) {
        y_lo = sh > 0.f ? static_cast<int>(floorf((iy - 1) / sh)) : 0;
        y_hi = sh > 0.f ? static_cast<int>(ceilf((iy + 1) / sh)) : hout - 1;
        x_lo = sw > 0.f ? static_cast<int>(floorf((ix - 1) / sw)) : 0;
        x_hi = sw > 0.f ? static_cast<int>(ceilf((ix + 1) / sw)) : wout - 1;
      } else {
        y_lo = static_cast<int>(floorf((iy - 1 + 0.5f) / sh - 0.5f));
        y_hi = static_cast<int>(ceilf((iy + 1 + 0.5f) / sh - 0.5f));
        x_lo = static_cast<int>(floorf((ix - 1 + 0.5f) / sw - 0.5f));
        x_hi = static_cast<int>(ceilf((ix + 1 + 0.5f) / sw - 0.5f));
      }
      y_lo = max(y_lo - 1, 0); y_hi = min(y_hi + 1, hout - 1);
      x_lo = max(x_lo - 1, 0); x_hi = min(x_hi + 1, wout - 1);
      for (int oy = y_lo; oy <= y_hi; ++oy) {
        const Lerp ly = lerp_index(oy, hin, sh, align);
        const float wy = (ly.i0 == iy ? ly.w0 : 0.f) + (ly.i1 == iy ? ly.w1 : 0.f);
        if (wy == 0.f) continue;
        for (int ox = x_lo; ox <= x_hi; ++ox) {
          const Lerp lx = lerp_index(ox, win, sw, align);
          const float wx = (lx.i0 == ix ? lx.w0 : 0.f) + (lx.i1 == ix ? lx.w1 : 0.f);
          if (wx == 0.f) continue;
          const float g = nchw ? dy[(img * C + ch) * plane + static_cast<long long>(oy) * wout + ox]
                               : dy[(img * plane + static_cast<long long>(oy) * wout + ox) * lddy + ch];
          acc += wy * wx * g;
        }
      }
    }
    dx[((img * hin + iy) * win + ix) * static_cast<long long>(lddx) + ch] = acc;
  }
}

}  // namespace cavp

using namespace cavp;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int cavp_zero(void* ptr, long long bytes, void* stream) {
  return static_cast<int>(cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), ST(stream)));
}
extern "C" int cavp_fill_strided(float* dst, int ld, long long rows, int c, float value, void* stream) {
  fill_strided_kernel<<<grid_for(rows * c, 256), 256, 0, ST(stream)>>>(dst, ld, rows, c, value);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_nchw_to_nhwc(const float* src, float* dst, int n, int c, int hw, int cpad, void* stream) {
  nchw_to_nhwc_kernel<<<grid_for(static_cast<long long>(n) * hw, 256), 256, 0, ST(stream)>>>(src, dst, n, c, hw, cpad);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_nhwc_to_nchw(const float* src, float* dst, int n, int c, int hw, int ld, void* stream) {
  nhwc_to_nchw_kernel<<<grid_for(static_cast<long long>(n) * hw, 256), 256, 0, ST(stream)>>>(src, dst, n, c, hw, ld);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_transpose(const float* src, float* dst, int rows, int cols, long long src_ld, long long dst_ld,
                              int batch, long long src_bs, long long dst_bs, void* stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
  transpose_kernel<<<grid, block, 0, ST(stream)>>>(src, dst, nullptr, rows, cols, src_ld, dst_ld, src_bs, dst_bs);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_transpose_split(const float* src, float* hi, float* lo, int rows, int cols, long long src_ld,
                                    long long dst_ld, int batch, long long src_bs, long long dst_bs, void* stream) {
  if (!src || !hi || !lo) return CAVP_ERR_NULL;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
  transpose_kernel<<<grid, block, 0, ST(stream)>>>(src, hi, lo, rows, cols, src_ld, dst_ld, src_bs, dst_bs);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_transpose_split_multi(const void* table, const int* work, int nwork, int bf16, void* stream) {
  if (!table || !work) return CAVP_ERR_NULL;
  if (nwork <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(table) & 7) || (reinterpret_cast<uintptr_t>(work) & 7)) return CAVP_ERR_ALIGN;
  const int grid = nwork < NUM_SMS * 16 ? nwork : NUM_SMS * 16;
  if (bf16)
    transpose_split_multi_kernel<true><<<grid, 256, 0, ST(stream)>>>(static_cast<const TsTensor*>(table),
                                                                     reinterpret_cast<const int2*>(work), nwork);
  else
    transpose_split_multi_kernel<false><<<grid, 256, 0, ST(stream)>>>(static_cast<const TsTensor*>(table),
                                                                      reinterpret_cast<const int2*>(work), nwork);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_add_inplace(float* dst, const float* src, long long n, float alpha, void* stream) {
  if (n & 3) return CAVP_ERR_ALIGN;
  add_inplace_kernel<<<grid_for(n / 4, 256), 256, 0, ST(stream)>>>(dst, src, n / 4, alpha);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_gather_rows(const float* src, const long long* idx, float* dst, int nrows, int c,
                                int accumulate_scatter, void* stream) {
  gather_rows_kernel<<<nrows, 128, 0, ST(stream)>>>(src, idx, dst, nrows, c, accumulate_scatter);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_bn_finalize(const float* partials, int nparts, int ldstat, int C, double count, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                float* mean_out, float* invstd_out, float* scale_out, float* shift_out, double* sums_io,
                                int sums_mode, long long* num_batches_tracked, void* stream) {
  constexpr int S = 16, WS_C = 8192;
  if (sums_mode != 2 && nparts >= 512 && C <= WS_C) {
    // two-stage path; the fp64 workspace is per device and reused by every call (launches of one stream are ordered;
    // the engine issues all BatchNorm work on one stream)
    static double* ws_dev[64] = {nullptr};
    static int* cnt_dev[64] = {nullptr};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return CAVP_ERR_ARG;
    if (!ws_dev[dev]) {
      cudaError_t e = cudaMalloc(&ws_dev[dev], sizeof(double) * S * 2 * WS_C);
      if (e != cudaSuccess) return static_cast<int>(e);
      e = cudaMalloc(&cnt_dev[dev], sizeof(int) * (WS_C / 32));
      if (e != cudaSuccess) return static_cast<int>(e);
      e = cudaMemset(cnt_dev[dev], 0, sizeof(int) * (WS_C / 32));  // once; the kernel leaves the tickets at zero
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    bn_reduce_finalize_kernel<<<dim3((C + 31) / 32, S), dim3(32, 32), 0, ST(stream)>>>(
        partials, nparts, ldstat, C, count, gamma, beta, running_mean, running_var, momentum, eps, mean_out, invstd_out,
        scale_out, shift_out, sums_io, sums_mode, num_batches_tracked, ws_dev[dev], cnt_dev[dev]);
    CAVP_LAUNCH_CHECK();
  }
  bn_finalize_kernel<float><<<(C + 31) / 32, dim3(32, 32), 0, ST(stream)>>>(
      partials, nparts, ldstat, C, count, gamma, beta, running_mean, running_var, momentum, eps, mean_out, invstd_out,
      scale_out, shift_out, sums_io, sums_mode, num_batches_tracked);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_bn_eval_coeffs(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                   int C, float* scale, float* shift, void* stream) {
  bn_eval_coeffs_kernel<<<(C + 255) / 256, 256, 0, ST(stream)>>>(gamma, beta, rm, rv, eps, C, scale, shift);
  CAVP_LAUNCH_CHECK();
}
// column lanes per 256-thread block for the column-stationary kernels: the candidate that wastes the fewest lanes on
// the last block (ties: wider); row blocks so that the grid is ~8 blocks per SM
static int pick_lc(int C4) {
  int LC = 64, best = -1;
  for (int cand = 64; cand >= 16; cand >>= 1) {
    const int util = 1000 * C4 / (((C4 + cand - 1) / cand) * cand);
    if (util > best) {
      best = util;
      LC = cand;
    }
  }
  return LC;
}
static dim3 colgrid(long long rows, int C4, int LC) {
  const int gx = (C4 + LC - 1) / LC, nrl = 256 / LC;
  long long gy = (rows + 4LL * nrl - 1) / (4LL * nrl);
  const long long cap = (8LL * NUM_SMS + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3(gx, static_cast<unsigned>(gy));
}
extern "C" int cavp_bn_apply(const float* y, int ldy, const float* scale, const float* shift, const float* res, int ldr,
                             float* out, int ldo, long long rows, int C, int act, float slope, void* stream) {
  if ((C & 3) || (ldy & 3) || (ldo & 3) || (res && (ldr & 3))) return CAVP_ERR_ALIGN;
  bn_apply_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, ST(stream)>>>(y, ldy, scale, shift, res, ldr, out, ldo, rows,
                                                                         C / 4, act, slope);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_colreduce(const float* dz, int lddz, const float* z, int ldz, const float* y, int ldy,
                              const float* mean, const float* invstd, long long rows, int C, int act, float slope,
                              float* gout, int ldg, float* partials, int ldp, int nblk, const float* zscale,
                              const float* zshift, void* stream) {
  if ((C & 3) || (lddz & 3) || (z && (ldz & 3)) || (y && (ldy & 3)) || (gout && (ldg & 3)) || (ldp & 3))
    return CAVP_ERR_ALIGN;
  const int C4 = C / 4;
  const int LC = pick_lc(C4);
  dim3 grid((C4 + LC - 1) / LC, nblk);
  colreduce_kernel<<<grid, 256, 0, ST(stream)>>>(dz, lddz, z, ldz, y, ldy, mean, invstd, rows, C4, act, slope, gout,
                                                 ldg, partials, ldp, zscale, zshift, LC);
  CAVP_LAUNCH_CHECK();
}
// few, large slabs (deterministic split-K: out = slab_0 + slab_1 + ... in that order): one float4 per thread and step
__global__ void slab_sum_kernel(const float* __restrict__ slabs, int nslabs, long long slab_elems, long long n4,
                                float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
    for (int k = 0; k < nslabs; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(slabs + k * slab_elems) + i);
      a += v.x; b += v.y; c += v.z; d += v.w;
    }
    reinterpret_cast<float4*>(out)[i] =
        make_float4(static_cast<float>(a), static_cast<float>(b), static_cast<float>(c), static_cast<float>(d));
  }
}
extern "C" int cavp_partials_sum(const float* partials, int nparts, int ldp, int C, int nk, float* out, void* stream) {
  if (nk == 1 && nparts <= 64 && C >= 16384 && (C & 3) == 0 && (ldp & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const long long n4 = C / 4;
    slab_sum_kernel<<<grid_for(n4, 256, 16), 256, 0, ST(stream)>>>(partials, nparts, ldp, n4, out);
    CAVP_LAUNCH_CHECK();
  }
  partials_sum_kernel<<<(nk * C + 31) / 32, dim3(32, 32), 0, ST(stream)>>>(partials, nparts, ldp, C, nk, out);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_bn_bwd_apply(const float* dz, int lddz, const float* z, int ldz, const float* y, int ldy,
                                 const float* mean, const float* invstd, const float* gamma, const float* sums,
                                 float inv_count, long long rows, int C, int act, float slope, float* dy, int lddy,
                                 float* dres, int lddres, const float* zscale, const float* zshift,
                                 const double* count_dev, float* dy_hi, float* dy_lo, int split_bf16, void* stream) {
  if ((C & 3) || (lddz & 3) || (ldy & 3) || (lddy & 3)) return CAVP_ERR_ALIGN;
  const int LC = pick_lc(C / 4);
  bn_bwd_apply_kernel<<<colgrid(rows, C / 4, LC), 256, 0, ST(stream)>>>(
      dz, lddz, z, ldz, y, ldy, mean, invstd, gamma, sums, inv_count, rows, C / 4, act, slope, dy, lddy, dres, lddres,
      zscale, zshift, LC, count_dev, dy_hi, dy_lo, split_bf16);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_maxpool_fwd(const float* x, int ldx, float* y, int ldy, int* idx, int n, int h, int w, int c, int k,
                                int stride, int pad, int ho, int wo, void* stream) {
  if ((c & 3) || (ldx & 3) || (ldy & 3)) return CAVP_ERR_ALIGN;
  maxpool_fwd_kernel<<<grid_for(static_cast<long long>(n) * ho * wo * (c / 4), 256), 256, 0, ST(stream)>>>(
      x, ldx, y, ldy, idx, n, h, w, c / 4, k, stride, pad, ho, wo);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_maxpool_bwd(const float* dy, int lddy, const int* idx, float* dx, int lddx, long long opix, int c,
                                void* stream) {
  maxpool_bwd_kernel<<<grid_for(opix * c, 256), 256, 0, ST(stream)>>>(dy, lddy, idx, dx, lddx, opix, c);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_pixel_sum(const float* x, int ldx, float* out, int n, int hw, int c, float scale, void* stream) {
  dim3 grid((c + 31) / 32, n), block(32, 8);
  pixel_sum_kernel<<<grid, block, 0, ST(stream)>>>(x, ldx, out, hw, c, scale);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_pixel_max(const float* x, int ldx, float* out, int* arg, int n, int hw, int c, void* stream) {
  dim3 grid((c + 127) / 128, n);
  pixel_max_kernel<<<grid, 128, 0, ST(stream)>>>(x, ldx, out, arg, hw, c);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_pixel_max_bwd(const float* dout, const int* arg, float* dx, int lddx, int n, int hw, int c,
                                  void* stream) {
  pixel_max_bwd_kernel<<<(n * c + 255) / 256, 256, 0, ST(stream)>>>(dout, arg, dx, lddx, hw, c, n);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_pixel_bcast(const float* dout, float* dx, int lddx, int n, int hw, int c, float scale,
                                int accumulate, void* stream) {
  if ((c & 3) || (lddx & 3)) return CAVP_ERR_ALIGN;
  const long long rows = static_cast<long long>(n) * hw;
  pixel_bcast_kernel<<<grid_for(rows * (c / 4), 256), 256, 0, ST(stream)>>>(dout, dx, lddx, hw, c / 4, rows, scale,
                                                                            accumulate);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_bilinear_fwd(const float* x, int ldx, int hin, int win, float* y, int ldy, int hout, int wout,
                                 int n, int c, int align_corners, int nchw_out, void* stream) {
  const float sh = lerp_scale(hin, hout, align_corners), sw = lerp_scale(win, wout, align_corners);
  if (nchw_out) {
    bilinear_fwd_nchw_kernel<<<grid_for(static_cast<long long>(n) * hout * wout, 256, 16), 256, 0, ST(stream)>>>(
        x, ldx, hin, win, y, hout, wout, n, c, align_corners, sh, sw);
  } else {
    if ((c & 3) || (ldx & 3) || (ldy & 3)) return CAVP_ERR_ALIGN;
    bilinear_fwd_nhwc_kernel<<<grid_for(static_cast<long long>(n) * hout * wout * (c / 4), 256, 16), 256, 0,
                               ST(stream)>>>(x, ldx, hin, win, y, ldy, hout, wout, n, c / 4, align_corners, sh, sw);
  }
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_bilinear_bwd(const float* dy, int lddy, int hout, int wout, float* dx, int lddx, int hin, int win,
                                 int n, int c, int align_corners, int nchw_in, int n_valid, void* stream) {
  const float sh = lerp_scale(hin, hout, align_corners), sw = lerp_scale(win, wout, align_corners);
  bilinear_bwd_kernel<<<grid_for(static_cast<long long>(n) * hin * win * c, 256, 16), 256, 0, ST(stream)>>>(
      dy, lddy, hout, wout, dx, lddx, hin, win, n, c, align_corners, sh, sw, nchw_in, n_valid);
  CAVP_LAUNCH_CHECK();
}
