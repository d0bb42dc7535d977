// Loss kernels.  Reference: loss/losser.py:60-62 (CrossEntropyLoss(ignore_index=255), mean over valid pixels) and
// loss/contrastive_aud.py:17-74 (pixel InfoNCE).  Coalesced, vectorised reductions with warp shuffles; every
// cross-block reduction goes through a partials buffer and a single finishing block, so results are deterministic.
#include <cstdlib>
#include "common.cuh"
#include "../../include/cavp_b200.h"

namespace cavp {

// ------------------------------------------------------------------------------------------------ cross entropy
// logits NCHW [B][C][HW] (only the first B images of the buffer are read), labels int64 [B][HW].
__global__ void ce_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C,
                              long long HW, int ignore_index, float* __restrict__ partials) {
  __shared__ float sh[32];
  const long long total = static_cast<long long>(B) * HW;
  float loss = 0.f, cnt = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long lab = labels[i];
    if (lab == ignore_index || lab < 0 || lab >= C) continue;  // out-of-range labels never index the logits
    const long long img = i / HW, pix = i - img * HW;
    const float* lp = logits + img * C * HW + pix;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, lp[c * HW]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(lp[c * HW] - mx);
    loss += mx + logf(s) - lp[lab * HW];
    cnt += 1.f;
  }
  const float bl = block_sum(loss, sh);
  const float bc = block_sum(cnt, sh);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2] = bl;
    partials[blockIdx.x * 2 + 1] = bc;
  }
}
// out[0] = sum(loss)/count, out[1] = count
__global__ void ce_finish_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  __shared__ float sh[32];
  double l = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    l += partials[i * 2];
    c += partials[i * 2 + 1];
  }
  const float fl = block_sum(static_cast<float>(l), sh);
  const float fc = block_sum(static_cast<float>(c), sh);
  if (threadIdx.x == 0) {
    out[0] = fl / fc;
    out[1] = fc;
  }
}
// dlogits = (softmax - onehot) * gscale[0] / count for valid pixels, 0 for ignored ones
__global__ void ce_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C,
                              long long HW, int ignore_index, const float* __restrict__ loss_and_count,
                              const float* __restrict__ gscale, float* __restrict__ dlogits) {
  const long long total = static_cast<long long>(B) * HW;
  const float coef = (gscale ? gscale[0] : 1.f) / loss_and_count[1];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long lab = labels[i];
    const long long img = i / HW, pix = i - img * HW;
    const float* lp = logits + img * C * HW + pix;
    float* dp = dlogits + img * C * HW + pix;
    if (lab == ignore_index || lab < 0 || lab >= C) {
      for (int c = 0; c < C; ++c) dp[c * HW] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, lp[c * HW]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(lp[c * HW] - mx);
    const float inv = 1.f / s;
    for (int c = 0; c < C; ++c) {
      const float pr = expf(lp[c * HW] - mx) * inv;
      dp[c * HW] = (pr - (c == lab ? 1.f : 0.f)) * coef;
    }
  }
}

// ------------------------------------------------------------------------------------- fused bilinear upsample + CE
// Training never needs the full-resolution prediction of forward_cls (models/cavp_model.py:138-141) as a tensor: the
// trainer feeds it straight into CrossEntropyLoss (trainer_cavp_vpo_mono.py:171,187; loss/losser.py:60-62).  These two
// kernels read only the low-resolution NHWC logits [n][hin][win][ldx] (B*nc*h*w*4 bytes instead of writing and
// re-reading B*nc*H*W*4 twice): every output pixel takes the bilinear sample (align_corners=False) of every class with
// the arithmetic of cavp_bilinear_fwd (bilerp), then its log-sum-exp.  The forward keeps the log-sum-exp per output
// pixel (4 bytes) for the backward.
__global__ void __launch_bounds__(256)
upsample_ce_fwd_kernel(const float* __restrict__ x, int ldx, int hin, int win, int hout, int wout, int B, int C,
                       float sh_, float sw_, const long long* __restrict__ labels, int ignore_index,
                       float* __restrict__ lse, float* __restrict__ partials) {
  __shared__ float sh[32];
  const long long plane = static_cast<long long>(hout) * wout;
  const long long total = static_cast<long long>(B) * plane;
  float loss = 0.f, cnt = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % wout);
    const int oy = static_cast<int>((i / wout) % hout);
    const long long img = i / plane;
    const Lerp ly = lerp_index(oy, hin, sh_, 0), lx = lerp_index(ox, win, sw_, 0);
    const float* base = x + img * hin * win * ldx;
    const float* pa = base + (static_cast<long long>(ly.i0) * win + lx.i0) * ldx;
    const float* pb = base + (static_cast<long long>(ly.i0) * win + lx.i1) * ldx;
    const float* pc = base + (static_cast<long long>(ly.i1) * win + lx.i0) * ldx;
    const float* pd = base + (static_cast<long long>(ly.i1) * win + lx.i1) * ldx;
    const long long lab = labels[i];
    float mx = -INFINITY, picked = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = bilerp(ly, lx, __ldg(pa + c), __ldg(pb + c), __ldg(pc + c), __ldg(pd + c));
      mx = fmaxf(mx, v);
      if (c == lab) picked = v;
    }
    float s = 0.f;
    for (int c = 0; c < C; ++c)
      s += expf(bilerp(ly, lx, __ldg(pa + c), __ldg(pb + c), __ldg(pc + c), __ldg(pd + c)) - mx);
    const float l = mx + logf(s);
    lse[i] = l;
    if (lab == ignore_index || lab < 0 || lab >= C) continue;
    loss += l - picked;
    cnt += 1.f;
  }
  const float bl = block_sum(loss, sh);
  const float bc = block_sum(cnt, sh);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 2] = bl;
    partials[blockIdx.x * 2 + 1] = bc;
  }
}
// Gather-form backward (deterministic, no atomics): thread = one low-resolution element (img, iy, ix, ch).  It visits
// the output pixels whose 2x2 footprint touches (iy, ix), recomputes that pixel's class-ch logit, and accumulates
// w * (softmax - onehot) * gscale / count.  Pad channels (ch >= C) and images >= n_valid get exact zeros, so the
// low-resolution gradient buffer needs no memset.
__global__ void __launch_bounds__(256)
upsample_ce_bwd_kernel(const float* __restrict__ x, int ldx, int hin, int win, int hout, int wout, int n, int n_valid,
                       int C, int Cpad, float sh_, float sw_, const long long* __restrict__ labels, int ignore_index,
                       const float* __restrict__ lse, const float* __restrict__ loss_and_count,
                       const float* __restrict__ gscale, float* __restrict__ dx, int lddx) {
  const long long total = static_cast<long long>(n) * hin * win * Cpad;
  const long long plane = static_cast<long long>(hout) * wout;
  const float coef = (gscale ? gscale[0] : 1.f) / loss_and_count[1];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % Cpad);
    const int ix = static_cast<int>((i / Cpad) % win);
    const int iy = static_cast<int>((i / (static_cast<long long>(Cpad) * win)) % hin);
    const long long img = i / (static_cast<long long>(Cpad) * win * hin);
    float acc = 0.f;
    if (img < n_valid && ch < C) {
      int y_lo = static_cast<int>(floorf((iy - 1 + 0.5f) / sh_ - 0.5f));
      int y_hi = static_cast<int>(ceilf((iy + 1 + 0.5f) / sh_ - 0.5f));
      int x_lo = static_cast<int>(floorf((ix - 1 + 0.5f) / sw_ - 0.5f));
      int x_hi = static_cast<int>(ceilf((ix + 1 + 0.5f) / sw_ - 0.5f));
      y_lo = max(y_lo - 1, 0); y_hi = min(y_hi + 1, hout - 1);
      x_lo = max(x_lo - 1, 0); x_hi = min(x_hi + 1, wout - 1);
      const float* base = x + img * hin * win * ldx + ch;
      for (int oy = y_lo; oy <= y_hi; ++oy) {
        const Lerp ly = lerp_index(oy, hin, sh_, 0);
        const float wy = (ly.i0 == iy ? ly.w0 : 0.f) + (ly.i1 == iy ? ly.w1 : 0.f);
        if (wy == 0.f) continue;
        const float* r0 = base + static_cast<long long>(ly.i0) * win * ldx;
        const float* r1 = base + static_cast<long long>(ly.i1) * win * ldx;
        for (int ox = x_lo; ox <= x_hi; ++ox) {
          const Lerp lx = lerp_index(ox, win, sw_, 0);
          const float wx = (lx.i0 == ix ? lx.w0 : 0.f) + (lx.i1 == ix ? lx.w1 : 0.f);
          if (wx == 0.f) continue;
          const long long o = img * plane + static_cast<long long>(oy) * wout + ox;
          const long long lab = labels[o];
          if (lab == ignore_index || lab < 0 || lab >= C) continue;
          const float v = bilerp(ly, lx, __ldg(r0 + static_cast<long long>(lx.i0) * ldx),
                                 __ldg(r0 + static_cast<long long>(lx.i1) * ldx),
                                 __ldg(r1 + static_cast<long long>(lx.i0) * ldx),
                                 __ldg(r1 + static_cast<long long>(lx.i1) * ldx));
          const float pr = expf(v - lse[o]);
          acc += wy * wx * (pr - (ch == lab ? 1.f : 0.f));
        }
      }
      acc *= coef;
    }
    dx[((img * hin + iy) * win + ix) * static_cast<long long>(lddx) + ch] = acc;
  }
}

// Exact 4x upsampling (the model's case: logits at stride 4 of an input whose sides are multiples of 4): the same
// gradient, tiled through shared memory.  A block owns an 8 x 8 tile of low-resolution pixels of one image.  The
// 36 x 36 full-resolution pixels whose footprint touches the tile are visited ONCE per 8-channel chunk: phase B
// evaluates g = softmax - onehot for (pixel, channel) from the low-resolution tile held in shared memory, phase C
// gathers the <= 8 x 8 weighted g values of every low-resolution element (weights from lerp_index, so the clamped
// borders match the forward exactly).  Deterministic (no atomics); ~6 shared-memory operations per (output pixel,
// channel) instead of 64 x (4 global loads + exp) per low-resolution element.
constexpr int UCE_T = 8;                  // low-resolution tile side
constexpr int UCE_R = 4 * UCE_T + 4;      // full-resolution region side (36)
constexpr int UCE_H = UCE_T + 2;          // low-resolution tile + halo (10)
constexpr int UCE_CC = 8;                 // channels per chunk
constexpr int UCE_GS = UCE_CC + 1;        // padded row of the g tile (bank spread)
constexpr int UCE_SMEM = UCE_R * UCE_R * (4 + 2) + UCE_H * UCE_H * UCE_CC * 4 + UCE_R * UCE_R * UCE_GS * 4;

__global__ void __launch_bounds__(256)
upsample_ce_bwd_x4_kernel(const float* __restrict__ x, int ldx, int hin, int win, int n_valid, int C, int Cpad,
                          const long long* __restrict__ labels, int ignore_index, const float* __restrict__ lse,
                          const float* __restrict__ loss_and_count, const float* __restrict__ gscale,
                          float* __restrict__ dx, int lddx, int tiles_x) {
  extern __shared__ __align__(16) unsigned char uce_smem[];
  float* lse_s = reinterpret_cast<float*>(uce_smem);
  short* lab_s = reinterpret_cast<short*>(lse_s + UCE_R * UCE_R);
  float* xs = reinterpret_cast<float*>(lab_s + UCE_R * UCE_R);
  float* gs = xs + UCE_H * UCE_H * UCE_CC;
  const int hout = 4 * hin, wout = 4 * win;
  const int img = blockIdx.y;
  const int y0 = (blockIdx.x / tiles_x) * UCE_T, x0 = (blockIdx.x % tiles_x) * UCE_T;
  const int oy0 = 4 * y0 - 2, ox0 = 4 * x0 - 2;
  const int tid = threadIdx.x;
  const float sh_ = 0.25f, sw_ = 0.25f;
  if (img >= n_valid) {  // zero-weighted half of the batch (trainer_cavp_vpo_mono.py:171): exact zeros
    for (int i = tid; i < UCE_T * UCE_T * Cpad; i += 256) {
      const int ch = i % Cpad, pp = i / Cpad;
      const int iy = y0 + pp / UCE_T, ix = x0 + pp % UCE_T;
      if (iy < hin && ix < win) dx[((static_cast<long long>(img) * hin + iy) * win + ix) * lddx + ch] = 0.f;
    }
    return;
  }
  const float coef = (gscale ? gscale[0] : 1.f) / loss_and_count[1];
  const long long plane = static_cast<long long>(hout) * wout;
  for (int i = tid; i < UCE_R * UCE_R; i += 256) {
    const int oy = oy0 + i / UCE_R, ox = ox0 + i % UCE_R;
    float l = 0.f;
    short lb = -1;
    if (oy >= 0 && oy < hout && ox >= 0 && ox < wout) {
      const long long o = img * plane + static_cast<long long>(oy) * wout + ox;
      const long long lab = labels[o];
      l = lse[o];
      lb = (lab == ignore_index || lab < 0 || lab >= C) ? static_cast<short>(-1) : static_cast<short>(lab);
    }
    lse_s[i] = l;
    lab_s[i] = lb;
  }
  // per-thread gather weights of phase C: items (pp, cc) = (tid >> 3) + 32 k, tid & 7
  const int cc = tid & 7;
  for (int c0 = 0; c0 < Cpad; c0 += UCE_CC) {
    __syncthreads();  // previous chunk's phase C is done with xs / gs (first pass: lse_s / lab_s are written)
    for (int i = tid; i < UCE_H * UCE_H * UCE_CC; i += 256) {
      const int ch = i % UCE_CC, pp = i / UCE_CC;
      int iy = y0 - 1 + pp / UCE_H, ix = x0 - 1 + pp % UCE_H;
      iy = min(max(iy, 0), hin - 1);
      ix = min(max(ix, 0), win - 1);
      xs[i] = (c0 + ch < C) ? __ldg(x + ((static_cast<long long>(img) * hin + iy) * win + ix) * ldx + c0 + ch) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < UCE_R * UCE_R * UCE_CC; i += 256) {
      const int ch = i % UCE_CC, o = i / UCE_CC;
      const int lb = lab_s[o];
      float g = 0.f;
      if (lb >= 0 && c0 + ch < C) {
        const int oy = oy0 + o / UCE_R, ox = ox0 + o % UCE_R;
        const Lerp ly = lerp_index(oy, hin, sh_, 0), lx = lerp_index(ox, win, sw_, 0);
        const int r0 = (ly.i0 - (y0 - 1)) * UCE_H, r1 = (ly.i1 - (y0 - 1)) * UCE_H;
        const int q0 = lx.i0 - (x0 - 1), q1 = lx.i1 - (x0 - 1);
        const float v = bilerp(ly, lx, xs[(r0 + q0) * UCE_CC + ch], xs[(r0 + q1) * UCE_CC + ch],
                               xs[(r1 + q0) * UCE_CC + ch], xs[(r1 + q1) * UCE_CC + ch]);
        g = expf(v - lse_s[o]) - (c0 + ch == lb ? 1.f : 0.f);
      }
      gs[o * UCE_GS + ch] = g;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int pp = (tid >> 3) + 32 * k;
      const int py = pp / UCE_T, px = pp % UCE_T;
      const int iy = y0 + py, ix = x0 + px;
      if (iy >= hin || ix >= win) continue;
      float wy[8], wx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int oy = 4 * iy - 2 + j, ox = 4 * ix - 2 + j;
        wy[j] = 0.f;
        wx[j] = 0.f;
        if (oy >= 0 && oy < hout) {
          const Lerp l = lerp_index(oy, hin, sh_, 0);
          wy[j] = (l.i0 == iy ? l.w0 : 0.f) + (l.i1 == iy ? l.w1 : 0.f);
        }
        if (ox >= 0 && ox < wout) {
          const Lerp l = lerp_index(ox, win, sw_, 0);
          wx[j] = (l.i0 == ix ? l.w0 : 0.f) + (l.i1 == ix ? l.w1 : 0.f);
        }
      }
      float acc = 0.f;
#pragma unroll
      for (int jy = 0; jy < 8; ++jy) {
        const float* row = gs + ((4 * py + jy) * UCE_R + 4 * px) * UCE_GS + cc;
        float t = 0.f;
#pragma unroll
        for (int jx = 0; jx < 8; ++jx) t = fmaf(wx[jx], row[jx * UCE_GS], t);
        acc = fmaf(wy[jy], t, acc);
      }
      if (c0 + cc < Cpad)
        dx[((static_cast<long long>(img) * hin + iy) * win + ix) * lddx + c0 + cc] = (c0 + cc < C) ? acc * coef : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------ contrastive
// anchors[i][:] = f[pix[i]][:] / max(||f[pix[i]]||, 1e-12)   (F.normalize over channels, then the gather of
// contrastive_aud.py:100-135).  One warp per anchor.  pix indexes pixels of the whole [rows*h*w] NHWC buffer.
__global__ void l2norm_gather_kernel(const float* __restrict__ f, int ld, const long long* __restrict__ pix, int A,
                                     int C4, float* __restrict__ anchors, int lda, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= A) return;
  const float4* src = reinterpret_cast<const float4*>(f + pix[i] * ld);
  float ss = 0.f;
  for (int k = lane; k < C4; k += 32) {
    const float4 v = src[k];
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  float4* dst = reinterpret_cast<float4*>(anchors + static_cast<long long>(i) * lda);
  for (int k = lane; k < C4; k += 32) {
    const float4 v = src[k];
    dst[k] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
  if (lane == 0) inv_norm[i] = inv;
}
// df[pix[i]] += (dA_i - a_i <a_i, dA_i>) * inv_norm_i
__global__ void l2norm_scatter_bwd_kernel(const float* __restrict__ danchors, const float* __restrict__ anchors,
                                          int lda, const float* __restrict__ inv_norm,
                                          const long long* __restrict__ pix, int A, int C4, float* __restrict__ df,
                                          int ld) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= A) return;
  const float4* da = reinterpret_cast<const float4*>(danchors + static_cast<long long>(i) * lda);
  const float4* a = reinterpret_cast<const float4*>(anchors + static_cast<long long>(i) * lda);
  float dot = 0.f;
  for (int k = lane; k < C4; k += 32) {
    const float4 u = da[k], w = a[k];
    dot += u.x * w.x + u.y * w.y + u.z * w.z + u.w * w.w;
  }
  dot = warp_sum(dot);
  const float inv = inv_norm[i];
  float* dst = df + pix[i] * ld;
  for (int k = lane; k < C4; k += 32) {
    const float4 u = da[k], w = a[k];
    atomicAdd(dst + k * 4 + 0, (u.x - w.x * dot) * inv);
    atomicAdd(dst + k * 4 + 1, (u.y - w.y * dot) * inv);
    atomicAdd(dst + k * 4 + 2, (u.z - w.z * dot) * inv);
    atomicAdd(dst + k * 4 + 3, (u.w - w.w * dot) * inv);
  }
}

// InfoNCE rows (contrastive_aud.py:41-74).  S = anchors anchors^T (not yet divided by the temperature), one block per
// row i:  z_ij = S_ij/T - max_j S_ij/T;  neg_i = sum_{lab_j != lab_i} exp z_ij;
// m_i = sum_{j != i, lab_j == lab_i} (z_ij - log(exp z_ij + neg_i)) / (P_i + 1e-12).
__global__ void infonce_fwd_kernel(const float* __restrict__ S, int lds, const long long* __restrict__ labels, int A,
                                   float inv_temp, float* __restrict__ rowmax, float* __restrict__ rowneg,
                                   float* __restrict__ rowmean) {
  __shared__ float sh[32];
  const int i = blockIdx.x;
  const float* row = S + static_cast<long long>(i) * lds;
  const long long li = labels[i];
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < A; j += blockDim.x) mx = fmaxf(mx, row[j] * inv_temp);
  mx = block_max(mx, sh);
  float neg = 0.f;
  for (int j = threadIdx.x; j < A; j += blockDim.x)
    if (labels[j] != li) neg += expf(row[j] * inv_temp - mx);
  neg = block_sum(neg, sh);
  float s = 0.f, cnt = 0.f;
  for (int j = threadIdx.x; j < A; j += blockDim.x)
    if (j != i && labels[j] == li) {
      const float z = row[j] * inv_temp - mx;
      s += z - logf(expf(z) + neg);
      cnt += 1.f;
    }
  s = block_sum(s, sh);
  cnt = block_sum(cnt, sh);
  if (threadIdx.x == 0) {
    rowmax[i] = mx;
    rowneg[i] = neg;
    rowmean[i] = s / (cnt + 1e-12f);
  }
}
// out[0] = scale * sum(x) (fp64 accumulate, single block)
__global__ void vec_sum_kernel(const float* __restrict__ x, int n, float scale, float* __restrict__ out) {
  __shared__ float sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  const float t = block_sum(static_cast<float>(s), sh);
  if (threadIdx.x == 0) out[0] = t * scale;
}
// G_ij = dL/dS_ij for loss = -mean_i m_i (times gscale[0]); written in place of S is NOT allowed (S is re-read).
__global__ void infonce_bwd_kernel(const float* __restrict__ S, int lds, const long long* __restrict__ labels, int A,
                                   float inv_temp, const float* __restrict__ rowmax, const float* __restrict__ rowneg,
                                   const float* __restrict__ gscale, float* __restrict__ G, int ldg) {
  __shared__ float sh[32];
  const int i = blockIdx.x;
  const float* row = S + static_cast<long long>(i) * lds;
  float* grow = G + static_cast<long long>(i) * ldg;
  const long long li = labels[i];
  const float mx = rowmax[i], neg = rowneg[i];
  float T = 0.f, cnt = 0.f;  // T_i = sum_{pos} 1/(exp z + neg)
  for (int j = threadIdx.x; j < A; j += blockDim.x)
    if (j != i && labels[j] == li) {
      T += 1.f / (expf(row[j] * inv_temp - mx) + neg);
      cnt += 1.f;
    }
  T = block_sum(T, sh);
  cnt = block_sum(cnt, sh);
  // dL/dm_i = -g/A ;  c_i = dL/dm_i / (P_i + eps)
  const float ci = -(gscale ? gscale[0] : 1.f) / static_cast<float>(A) / (cnt + 1e-12f) * inv_temp;
  for (int j = threadIdx.x; j < ldg; j += blockDim.x) {
    float g = 0.f;
    if (j < A && cnt > 0.f) {
      const float e = expf(row[j] * inv_temp - mx);
      if (labels[j] != li) {
        g = -ci * e * T;
      } else if (j != i) {
        g = ci * (1.f - e / (e + neg));
      }
    }
    grow[j] = g;
  }
}

}  // namespace cavp

using namespace cavp;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int cavp_ce_nblocks(int B, long long HW) { return grid_for(static_cast<long long>(B) * HW, 256, 8); }
extern "C" int cavp_ce_fwd(const float* logits, const long long* labels, int B, int C, long long HW, int ignore_index,
                           float* partials, float* loss_and_count, void* stream) {
  const int nb = cavp_ce_nblocks(B, HW);
  ce_fwd_kernel<<<nb, 256, 0, ST(stream)>>>(logits, labels, B, C, HW, ignore_index, partials);
  ce_finish_kernel<<<1, 256, 0, ST(stream)>>>(partials, nb, loss_and_count);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_ce_bwd(const float* logits, const long long* labels, int B, int C, long long HW, int ignore_index,
                           const float* loss_and_count, const float* gscale, float* dlogits, void* stream) {
  ce_bwd_kernel<<<grid_for(static_cast<long long>(B) * HW, 256, 16), 256, 0, ST(stream)>>>(
      logits, labels, B, C, HW, ignore_index, loss_and_count, gscale, dlogits);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_upsample_ce_fwd(const float* x, int ldx, int hin, int win, int hout, int wout, int B, int C,
                                    const long long* labels, int ignore_index, float* lse, float* partials,
                                    float* loss_and_count, void* stream) {
  if (!x || !labels || !lse || !partials || !loss_and_count) return CAVP_ERR_NULL;
  if (B <= 0 || C <= 0 || hin <= 0 || win <= 0 || hout <= 0 || wout <= 0 || ldx < C) return CAVP_ERR_ARG;
  const int nb = cavp_ce_nblocks(B, static_cast<long long>(hout) * wout);
  upsample_ce_fwd_kernel<<<nb, 256, 0, ST(stream)>>>(x, ldx, hin, win, hout, wout, B, C, lerp_scale(hin, hout, 0),
                                                     lerp_scale(win, wout, 0), labels, ignore_index, lse, partials);
  ce_finish_kernel<<<1, 256, 0, ST(stream)>>>(partials, nb, loss_and_count);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_upsample_ce_bwd(const float* x, int ldx, int hin, int win, int hout, int wout, int n, int n_valid,
                                    int C, int Cpad, const long long* labels, int ignore_index, const float* lse,
                                    const float* loss_and_count, const float* gscale, float* dx, int lddx,
                                    void* stream) {
  if (!x || !labels || !lse || !loss_and_count || !dx) return CAVP_ERR_NULL;
  if (n <= 0 || n_valid < 0 || n_valid > n || C <= 0 || Cpad < C || ldx < C || lddx < Cpad) return CAVP_ERR_ARG;
  static const bool no_tiled = getenv("CAVP_CE_BWD_GENERIC") != nullptr;  // force the generic kernel (tests)
  if (hout == 4 * hin && wout == 4 * win && C < 32768 && !no_tiled) {
    auto kern = upsample_ce_bwd_x4_kernel;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, UCE_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int tiles_x = (win + UCE_T - 1) / UCE_T, tiles_y = (hin + UCE_T - 1) / UCE_T;
    kern<<<dim3(tiles_x * tiles_y, n), 256, UCE_SMEM, ST(stream)>>>(x, ldx, hin, win, n_valid, C, Cpad, labels,
                                                                   ignore_index, lse, loss_and_count, gscale, dx, lddx,
                                                                   tiles_x);
    CAVP_LAUNCH_CHECK();
  }
  upsample_ce_bwd_kernel<<<grid_for(static_cast<long long>(n) * hin * win * Cpad, 256, 16), 256, 0, ST(stream)>>>(
      x, ldx, hin, win, hout, wout, n, n_valid, C, Cpad, lerp_scale(hin, hout, 0), lerp_scale(win, wout, 0), labels,
      ignore_index, lse, loss_and_count, gscale, dx, lddx);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_l2norm_gather(const float* f, int ld, const long long* pix, int A, int C, float* anchors, int lda,
                                  float* inv_norm, void* stream) {
  if ((C & 3) || (ld & 3) || (lda & 3)) return CAVP_ERR_ALIGN;
  l2norm_gather_kernel<<<(A + 7) / 8, 256, 0, ST(stream)>>>(f, ld, pix, A, C / 4, anchors, lda, inv_norm);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_l2norm_scatter_bwd(const float* danchors, const float* anchors, int lda, const float* inv_norm,
                                       const long long* pix, int A, int C, float* df, int ld, void* stream) {
  if ((C & 3) || (ld & 3) || (lda & 3)) return CAVP_ERR_ALIGN;
  l2norm_scatter_bwd_kernel<<<(A + 7) / 8, 256, 0, ST(stream)>>>(danchors, anchors, lda, inv_norm, pix, A, C / 4, df,
                                                                 ld);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_infonce_fwd(const float* S, int lds, const long long* labels, int A, float temperature,
                                float* rowmax, float* rowneg, float* rowmean, float* loss, void* stream) {
  infonce_fwd_kernel<<<A, 256, 0, ST(stream)>>>(S, lds, labels, A, 1.f / temperature, rowmax, rowneg, rowmean);
  vec_sum_kernel<<<1, 256, 0, ST(stream)>>>(rowmean, A, -1.f / static_cast<float>(A), loss);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_infonce_bwd(const float* S, int lds, const long long* labels, int A, float temperature,
                                const float* rowmax, const float* rowneg, const float* gscale, float* G, int ldg,
                                void* stream) {
  infonce_bwd_kernel<<<A, 256, 0, ST(stream)>>>(S, lds, labels, A, 1.f / temperature, rowmax, rowneg, gscale, G, ldg);
  CAVP_LAUNCH_CHECK();
}
