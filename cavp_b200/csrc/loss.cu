// Loss kernels.  Reference: loss/losser.py:60-62 (CrossEntropyLoss(ignore_index=255), mean over valid pixels) and
// loss/contrastive_aud.py:17-74 (pixel InfoNCE).  Coalesced, vectorised reductions with warp shuffles; every
// cross-block reduction goes through a partials buffer and a single finishing block, so results are deterministic.
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
    if (lab == ignore_index) continue;
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
    if (lab == ignore_index) {
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
