// AVContrast (loss/av_contrast.py:20-112; SURVEY.md 8(f) N4 - instantiated by loss/losser.py:57 but never called by any
// reference trainer; built because BASELINE.json's north_star names it).
//
//   f_v [b][hw][c] -> F.normalize over dim=1 (the hw axis, as written in the reference :92) -> masked average over the
//   foreground pixels of the 128x128-resized labels (:97-108) -> SupCon-style loss between the b normalised audio
//   vectors and the b pooled visual vectors with "same foreground class" positives (:20-86).
// HBM-bound: two passes over f_v (column statistics forward, gradient backward; 4*b*hw*c bytes each way); everything
// else is 2b x 2b and runs in one block.
//   colstats: per (image, channel)  ss = sum_hw f^2,  ms = sum_hw mask*f   (deterministic two-level reduction)
//   loss    : n = max(sqrt(ss), 1e-12); mv = ms / n / (cnt + eps); a = f_a / max(|f_a|, 1e-12); S = [a; mv][a; mv]^T / T;
//             loss and its gradient w.r.t. f_a, ms and n in the same launch (the backward only rescales)
//   bwd     : df = g * (mask * dms + f * dnn)
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cavp_b200.h"
#include "common.cuh"

namespace cavp {

__global__ void __launch_bounds__(256)
avc_colstats_kernel(const float* __restrict__ fv, const unsigned char* __restrict__ mask, int hw, int C4,
                    float* __restrict__ partials /* [b][nchunk][2][C] */) {
  extern __shared__ float4 sh4[];  // [nrl][2][C4]
  const int nchunk = gridDim.x, chunk = blockIdx.x, img = blockIdx.y;
  const int nrl = blockDim.x / C4;
  const int cl = threadIdx.x % C4, rl = threadIdx.x / C4;
  const int rows_per = (hw + nchunk - 1) / nchunk;
  const int rbeg = chunk * rows_per, rend = min(hw, rbeg + rows_per);
  float4 ss = make_float4(0.f, 0.f, 0.f, 0.f), ms = ss;
  if (rl < nrl) {
    const float* base = fv + (static_cast<size_t>(img) * hw) * C4 * 4;
    const unsigned char* mrow = mask + static_cast<size_t>(img) * hw;
    for (int r = rbeg + rl; r < rend; r += nrl) {
      const float4 f = *reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * C4 * 4 + cl * 4);
      const float m = mrow[r] ? 1.f : 0.f;
      ss.x = fmaf(f.x, f.x, ss.x); ss.y = fmaf(f.y, f.y, ss.y); ss.z = fmaf(f.z, f.z, ss.z); ss.w = fmaf(f.w, f.w, ss.w);
      ms.x = fmaf(m, f.x, ms.x); ms.y = fmaf(m, f.y, ms.y); ms.z = fmaf(m, f.z, ms.z); ms.w = fmaf(m, f.w, ms.w);
    }
    sh4[(rl * 2 + 0) * C4 + cl] = ss;
    sh4[(rl * 2 + 1) * C4 + cl] = ms;
  }
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < nrl; ++k) {
      const float4 a = sh4[(k * 2 + 0) * C4 + cl], b = sh4[(k * 2 + 1) * C4 + cl];
      ss.x += a.x; ss.y += a.y; ss.z += a.z; ss.w += a.w;
      ms.x += b.x; ms.y += b.y; ms.z += b.z; ms.w += b.w;
    }
    float* pp = partials + (static_cast<size_t>(img) * nchunk + chunk) * 2 * C4 * 4;
    *reinterpret_cast<float4*>(pp + cl * 4) = ss;
    *reinterpret_cast<float4*>(pp + C4 * 4 + cl * 4) = ms;
  }
}

// one block; n2 = 2b anchors.  feats [n2][C] scratch, dfeat [n2][C] scratch (global), S / G [n2][n2] in shared memory.
__global__ void __launch_bounds__(1024)
avc_loss_kernel(const float* __restrict__ partials, int nchunk, int b, int C, const float* __restrict__ fa,
                const float* __restrict__ cnt, const int* __restrict__ target, float temperature, float eps,
                float* __restrict__ feats, float* __restrict__ dfeat, float* __restrict__ nrm /* [b][C] */,
                float* __restrict__ msum /* [b][C] */, float* __restrict__ loss, float* __restrict__ d_fa,
                float* __restrict__ dms, float* __restrict__ dnn) {
  extern __shared__ float shf[];
  const int n2 = 2 * b;
  float* S = shf;             // [n2][n2]
  float* G = shf + n2 * n2;   // [n2][n2]
  float* rowv = G + n2 * n2;  // [n2] mean_log_prob_pos, then reused
  float* anorm = rowv + n2;   // [b] |f_a|
  const int tid = threadIdx.x, nt = blockDim.x;
  // 1. column statistics -> n, mv ; audio norms
  for (int i = tid; i < b * C; i += nt) {
    const int img = i / C, ch = i - img * C;
    float ss = 0.f, ms = 0.f;
    for (int k = 0; k < nchunk; ++k) {
      const float* pp = partials + (static_cast<size_t>(img) * nchunk + k) * 2 * C;
      ss += pp[ch];
      ms += pp[C + ch];
    }
    const float n = fmaxf(sqrtf(ss), 1e-12f);
    nrm[i] = n;
    msum[i] = ms;
    feats[(b + img) * C + ch] = ms / n / (cnt[img] + eps);
  }
  for (int img = tid; img < b; img += nt) {
    float s = 0.f;
    for (int ch = 0; ch < C; ++ch) s = fmaf(fa[img * C + ch], fa[img * C + ch], s);
    anorm[img] = fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  for (int i = tid; i < b * C; i += nt) feats[i] = fa[i] / anorm[i / C];
  __syncthreads();
  // 2. S = feats feats^T / T
  for (int e = tid; e < n2 * n2; e += nt) {
    const int i = e / n2, j = e - i * n2;
    const float* fi = feats + i * C;
    const float* fj = feats + j * C;
    float s = 0.f;
    for (int ch = 0; ch < C; ++ch) s = fmaf(fi[ch], fj[ch], s);
    S[e] = s / temperature;
  }
  __syncthreads();
  // 3. per-anchor loss terms and G = dL/dS
  for (int i = tid; i < n2; i += nt) {
    float mx = -INFINITY;
    for (int j = 0; j < n2; ++j) mx = fmaxf(mx, S[i * n2 + j]);
    float esum = 0.f;
    for (int j = 0; j < n2; ++j)
      if (j != i) esum += expf(S[i * n2 + j] - mx);
    const int ti = target[i % b];
    float pos = 0.f, acc = 0.f;
    const float lse = logf(esum);
    for (int j = 0; j < n2; ++j) {
      const bool m = (j != i) && ti >= 0 && target[j % b] == ti;
      if (m) {
        pos += 1.f;
        acc += (S[i * n2 + j] - mx) - lse;
      }
    }
    rowv[i] = acc / (pos + eps);
    const float w = -1.f / (static_cast<float>(n2) * (pos + eps));
    for (int j = 0; j < n2; ++j) {
      const bool m = (j != i) && ti >= 0 && target[j % b] == ti;
      const float p = (j != i) ? expf(S[i * n2 + j] - mx) / esum : 0.f;
      G[i * n2 + j] = w * ((m ? 1.f : 0.f) - pos * p);
    }
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < n2; ++i) s += rowv[i];
    loss[0] = -s / static_cast<float>(n2);
  }
  // 4. dfeat = (G + G^T) feats / T
  for (int e = tid; e < n2 * C; e += nt) {
    const int i = e / C, ch = e - i * C;
    float s = 0.f;
    for (int k = 0; k < n2; ++k) s = fmaf(G[i * n2 + k] + G[k * n2 + i], feats[k * C + ch], s);
    dfeat[e] = s / temperature;
  }
  __syncthreads();
  // 5. back through the two normalisations
  for (int img = tid; img < b; img += nt) {  // audio: a = f / |f|
    float dot = 0.f;
    for (int ch = 0; ch < C; ++ch) dot = fmaf(feats[img * C + ch], dfeat[img * C + ch], dot);
    rowv[img] = dot;
  }
  __syncthreads();
  for (int i = tid; i < b * C; i += nt) {
    const int img = i / C;
    d_fa[i] = (dfeat[i] - feats[i] * rowv[img]) / anorm[img];
    const float n = nrm[i], cp = cnt[img] + eps, dmv = dfeat[b * C + i];
    dms[i] = dmv / (n * cp);
    dnn[i] = n > 1e-12f ? -dmv * msum[i] / (n * n * n * cp) : 0.f;  // (d/dn) / n : df += f * dnn
  }
}

__global__ void avc_bwd_kernel(const float* __restrict__ fv, const unsigned char* __restrict__ mask,
                               const float* __restrict__ dms, const float* __restrict__ dnn,
                               const float* __restrict__ gscale, int hw, int C4, long long total4,
                               float* __restrict__ dfv) {
  const float g = gscale ? gscale[0] : 1.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / C4;
    const int cl = static_cast<int>(i - row * C4);
    const long long img = row / hw;
    const float m = mask[row] ? 1.f : 0.f;
    const float4 f = reinterpret_cast<const float4*>(fv)[i];
    const float4 a = *reinterpret_cast<const float4*>(dms + (img * C4 + cl) * 4);
    const float4 n = *reinterpret_cast<const float4*>(dnn + (img * C4 + cl) * 4);
    reinterpret_cast<float4*>(dfv)[i] = make_float4(g * fmaf(f.x, n.x, m * a.x), g * fmaf(f.y, n.y, m * a.y),
                                                    g * fmaf(f.z, n.z, m * a.z), g * fmaf(f.w, n.w, m * a.w));
  }
}

}  // namespace cavp

using namespace cavp;

extern "C" int cavp_avc_colstats(const float* fv, const unsigned char* mask, int b, int hw, int c, int nchunk,
                                 float* partials, void* stream) {
  if (!fv || !mask || !partials) return CAVP_ERR_NULL;
  if ((c & 3) || (reinterpret_cast<uintptr_t>(fv) & 15) || (reinterpret_cast<uintptr_t>(partials) & 15)) return CAVP_ERR_ALIGN;
  if (b <= 0 || hw <= 0 || nchunk <= 0 || c / 4 > 256) return CAVP_ERR_ARG;
  const int C4 = c / 4, nrl = 256 / C4;
  dim3 grid(nchunk, b);
  avc_colstats_kernel<<<grid, 256, nrl * 2 * C4 * sizeof(float4), static_cast<cudaStream_t>(stream)>>>(fv, mask, hw, C4,
                                                                                                      partials);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_avc_loss(const float* partials, int nchunk, int b, int c, const float* fa, const float* cnt,
                             const int* target, float temperature, float eps, float* feats, float* dfeat, float* nrm,
                             float* msum, float* loss, float* d_fa, float* dms, float* dnn, void* stream) {
  if (!partials || !fa || !cnt || !target || !feats || !dfeat || !nrm || !msum || !loss || !d_fa || !dms || !dnn)
    return CAVP_ERR_NULL;
  if (b <= 0 || c <= 0 || !(temperature > 0.f)) return CAVP_ERR_ARG;
  const int n2 = 2 * b;
  const size_t smem = (static_cast<size_t>(2) * n2 * n2 + n2 + b) * sizeof(float);
  if (smem > 200 * 1024) return CAVP_ERR_ARG;  // b <= ~110
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(avc_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  avc_loss_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(partials, nchunk, b, c, fa, cnt, target, temperature,
                                                                     eps, feats, dfeat, nrm, msum, loss, d_fa, dms, dnn);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_avc_bwd(const float* fv, const unsigned char* mask, const float* dms, const float* dnn,
                            const float* gscale, int b, int hw, int c, float* dfv, void* stream) {
  if (!fv || !mask || !dms || !dnn || !dfv) return CAVP_ERR_NULL;
  if ((c & 3) || (reinterpret_cast<uintptr_t>(fv) & 15) || (reinterpret_cast<uintptr_t>(dfv) & 15)) return CAVP_ERR_ALIGN;
  const long long total4 = static_cast<long long>(b) * hw * (c / 4);
  avc_bwd_kernel<<<grid_for(total4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(fv, mask, dms, dnn, gscale, hw, c / 4,
                                                                                     total4, dfv);
  return static_cast<int>(cudaGetLastError());
}
