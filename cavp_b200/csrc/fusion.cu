// Cross-modal fusion kernels (reference: models/attn.py): LayerNorm, the single-key sigmoid gate that
// models/attn.py:73-106 reduces to when the key/value sequence is the one audio token, GELU backward.
// All are HBM-bound streaming kernels: one warp per token, float4 accesses, warp-shuffle reductions.
#include "common.cuh"
#include "../../include/cavp_b200.h"

namespace cavp {

constexpr int LN_MAX_V4 = 10;  // supports C <= 1280

// ------------------------------------------------------------------------------------------------ LayerNorm
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, long long T, int C4, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invC = 1.f / (C4 * 4);
  for (long long t = warp0; t < T; t += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + t * C4 * 4);
    float4 v[LN_MAX_V4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_V4; ++k) {
      const int f = lane + 32 * k;
      if (f < C4) {
        v[k] = xr[f];
        s += v[k].x + v[k].y + v[k].z + v[k].w;
      }
    }
    const float mean = warp_sum(s) * invC;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_V4; ++k) {
      const int f = lane + 32 * k;
      if (f < C4) {
        const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        ss += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) * invC + eps);
    float4* yr = reinterpret_cast<float4*>(y + t * C4 * 4);
#pragma unroll
    for (int k = 0; k < LN_MAX_V4; ++k) {
      const int f = lane + 32 * k;
      if (f < C4) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[f];
        const float4 b = reinterpret_cast<const float4*>(beta)[f];
        yr[f] = make_float4((v[k].x - mean) * rstd * g.x + b.x, (v[k].y - mean) * rstd * g.y + b.y,
                            (v[k].z - mean) * rstd * g.z + b.z, (v[k].w - mean) * rstd * g.w + b.w);
      }
    }
    if (lane == 0) {
      mean_out[t] = mean;
      rstd_out[t] = rstd;
    }
  }
}
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.  Optionally dx += add (residual gradient).
// Per-warp partial sums of dgamma / dbeta go to partials[warp][2][C].
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, const float* __restrict__ add,
                                     float* __restrict__ dx, float* __restrict__ partials, long long T, int C4) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invC = 1.f / (C4 * 4);
  float4 dg[LN_MAX_V4], db[LN_MAX_V4];
#pragma unroll
  for (int k = 0; k < LN_MAX_V4; ++k) dg[k] = db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long t = warp0; t < T; t += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + t * C4 * 4);
    const float4* gr = reinterpret_cast<const float4*>(dy + t * C4 * 4);
    const float mu = mean[t], rs = rstd[t];
    float4 xh[LN_MAX_V4], g[LN_MAX_V4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_V4; ++k) {
      const int f = lane + 32 * k;
      if (f < C4) {
        const float4 xv = xr[f], d = gr[f], ga = reinterpret_cast<const float4*>(gamma)[f];
        xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[k] = make_float4(d.x * ga.x, d.y * ga.y, d.z * ga.z, d.w * ga.w);
        s1 += g[k].x + g[k].y + g[k].z + g[k].w;
        s2 += g[k].x * xh[k].x + g[k].y * xh[k].y + g[k].z * xh[k].z + g[k].w * xh[k].w;
        dg[k].x += d.x * xh[k].x; dg[k].y += d.y * xh[k].y; dg[k].z += d.z * xh[k].z; dg[k].w += d.w * xh[k].w;
        db[k].x += d.x; db[k].y += d.y; db[k].z += d.z; db[k].w += d.w;
      }
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
    float4* dr = reinterpret_cast<float4*>(dx + t * C4 * 4);
#pragma unroll
    for (int k = 0; k < LN_MAX_V4; ++k) {
      const int f = lane + 32 * k;
      if (f < C4) {
        float4 o = make_float4(rs * (g[k].x - s1 - xh[k].x * s2), rs * (g[k].y - s1 - xh[k].y * s2),
                               rs * (g[k].z - s1 - xh[k].z * s2), rs * (g[k].w - s1 - xh[k].w * s2));
        if (add) {
          const float4 a = reinterpret_cast<const float4*>(add + t * C4 * 4)[f];
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        dr[f] = o;
      }
    }
  }
  float* pp = partials + warp0 * 2 * (C4 * 4);
#pragma unroll
  for (int k = 0; k < LN_MAX_V4; ++k) {
    const int f = lane + 32 * k;
    if (f < C4) {
      reinterpret_cast<float4*>(pp)[f] = dg[k];
      reinterpret_cast<float4*>(pp + C4 * 4)[f] = db[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ sigmoid gate
// models/attn.py:73-106 with one key token per row:  attn[r,h,n] = sigmoid(scale * <q[r%Bq,n,h,:], k[r,h,:]>),
// x[r,n,h,:] = attn[r,h,n] * v[r,h,:].   q rows are shared by rows r, r+Bq, ... (train mode duplicates the visual half,
// models/cavp_model.py:181).  C = H * D, D % 4 == 0; float4 index f belongs to head f / (D/4).
constexpr int GATE_MAX_V4 = 3;   // C <= 384
constexpr int GATE_MAX_REP = 2;  // rows sharing one q row
constexpr int GATE_HEADS = 4;

__global__ void gate_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                float* __restrict__ x, float* __restrict__ attn, int Bq, int rep, int N, int C4, int D4,
                                float scale) {
  extern __shared__ float4 kv_sh[];  // [rep][2][C4]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < rep * 2 * C4; i += blockDim.x) {
    const int rp = i / (2 * C4), which = (i / C4) & 1, f = i % C4;
    const int r = b + rp * Bq;
    kv_sh[i] = reinterpret_cast<const float4*>((which ? v : k) + static_cast<long long>(r) * C4 * 4)[f];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int n = blockIdx.x * nw + warp; n < N; n += gridDim.x * nw) {
    const float4* qr = reinterpret_cast<const float4*>(q + (static_cast<long long>(b) * N + n) * C4 * 4);
    float4 qv[GATE_MAX_V4];
#pragma unroll
    for (int kk = 0; kk < GATE_MAX_V4; ++kk) {
      const int f = lane + 32 * kk;
      if (f < C4) qv[kk] = qr[f];
    }
    for (int rp = 0; rp < rep; ++rp) {
      const float4* ks = kv_sh + rp * 2 * C4;
      const float4* vs = ks + C4;
      float part[GATE_HEADS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < GATE_MAX_V4; ++kk) {
        const int f = lane + 32 * kk;
        if (f < C4) {
          const float4 kf = ks[f];
          const float d = qv[kk].x * kf.x + qv[kk].y * kf.y + qv[kk].z * kf.z + qv[kk].w * kf.w;
          const int h = f / D4;
#pragma unroll
          for (int hh = 0; hh < GATE_HEADS; ++hh) part[hh] += (h == hh) ? d : 0.f;
        }
      }
      float a[GATE_HEADS];
#pragma unroll
      for (int hh = 0; hh < GATE_HEADS; ++hh) a[hh] = 1.f / (1.f + expf(-warp_sum(part[hh]) * scale));
      const long long r = b + static_cast<long long>(rp) * Bq;
      float4* xr = reinterpret_cast<float4*>(x + (r * N + n) * C4 * 4);
#pragma unroll
      for (int kk = 0; kk < GATE_MAX_V4; ++kk) {
        const int f = lane + 32 * kk;
        if (f < C4) {
          const int h = f / D4;
          const float ah = h == 0 ? a[0] : (h == 1 ? a[1] : (h == 2 ? a[2] : a[3]));
          const float4 vf = vs[f];
          xr[f] = make_float4(ah * vf.x, ah * vf.y, ah * vf.z, ah * vf.w);
        }
      }
      if (lane < GATE_HEADS) {
        const float ah = lane == 0 ? a[0] : (lane == 1 ? a[1] : (lane == 2 ? a[2] : a[3]));
        attn[(r * GATE_HEADS + lane) * N + n] = ah;
      }
    }
  }
}

// backward: dq (summed over the rows that share q), dk / dv accumulated with atomics (pre-zeroed [rows][C]).
__global__ void gate_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ q, const float* __restrict__ k,
                                const float* __restrict__ v, const float* __restrict__ attn, float* __restrict__ dq,
                                float* __restrict__ dk, float* __restrict__ dv, int Bq, int rep, int N, int C4, int D4,
                                float scale) {
  extern __shared__ float4 kv_sh[];  // [rep][2][C4] k,v  then [rep][2][C4] block accumulators (as floats)
  float* acc_sh = reinterpret_cast<float*>(kv_sh + rep * 2 * C4);
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < rep * 2 * C4; i += blockDim.x) {
    const int rp = i / (2 * C4), which = (i / C4) & 1, f = i % C4;
    const int r = b + rp * Bq;
    kv_sh[i] = reinterpret_cast<const float4*>((which ? v : k) + static_cast<long long>(r) * C4 * 4)[f];
  }
  for (int i = threadIdx.x; i < rep * 2 * C4 * 4; i += blockDim.x) acc_sh[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float4 dk_acc[GATE_MAX_REP][GATE_MAX_V4], dv_acc[GATE_MAX_REP][GATE_MAX_V4];
#pragma unroll
  for (int rp = 0; rp < GATE_MAX_REP; ++rp)
#pragma unroll
    for (int kk = 0; kk < GATE_MAX_V4; ++kk) dk_acc[rp][kk] = dv_acc[rp][kk] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int n = blockIdx.x * nw + warp; n < N; n += gridDim.x * nw) {
    const float4* qr = reinterpret_cast<const float4*>(q + (static_cast<long long>(b) * N + n) * C4 * 4);
    float4 qv[GATE_MAX_V4], dqv[GATE_MAX_V4];
#pragma unroll
    for (int kk = 0; kk < GATE_MAX_V4; ++kk) {
      const int f = lane + 32 * kk;
      if (f < C4) qv[kk] = qr[f];
      dqv[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int rp = 0; rp < GATE_MAX_REP; ++rp) {
      if (rp < rep) {
        const float4* ks = kv_sh + rp * 2 * C4;
        const float4* vs = ks + C4;
        const long long r = b + static_cast<long long>(rp) * Bq;
        const float4* gr = reinterpret_cast<const float4*>(dx + (r * N + n) * C4 * 4);
        float4 g[GATE_MAX_V4];
        float part[GATE_HEADS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < GATE_MAX_V4; ++kk) {
          const int f = lane + 32 * kk;
          if (f < C4) {
            g[kk] = gr[f];
            const float4 vf = vs[f];
            const float d = g[kk].x * vf.x + g[kk].y * vf.y + g[kk].z * vf.z + g[kk].w * vf.w;
            const int h = f / D4;
#pragma unroll
            for (int hh = 0; hh < GATE_HEADS; ++hh) part[hh] += (h == hh) ? d : 0.f;
          }
        }
        float a[GATE_HEADS], ds[GATE_HEADS];
#pragma unroll
        for (int hh = 0; hh < GATE_HEADS; ++hh) {
          a[hh] = attn[(r * GATE_HEADS + hh) * N + n];
          ds[hh] = warp_sum(part[hh]) * a[hh] * (1.f - a[hh]) * scale;
        }
#pragma unroll
        for (int kk = 0; kk < GATE_MAX_V4; ++kk) {
          const int f = lane + 32 * kk;
          if (f < C4) {
            const int h = f / D4;
            const float dsh = h == 0 ? ds[0] : (h == 1 ? ds[1] : (h == 2 ? ds[2] : ds[3]));
            const float ah = h == 0 ? a[0] : (h == 1 ? a[1] : (h == 2 ? a[2] : a[3]));
            const float4 kf = ks[f];
            dqv[kk].x += dsh * kf.x; dqv[kk].y += dsh * kf.y; dqv[kk].z += dsh * kf.z; dqv[kk].w += dsh * kf.w;
            dk_acc[rp][kk].x += dsh * qv[kk].x; dk_acc[rp][kk].y += dsh * qv[kk].y;
            dk_acc[rp][kk].z += dsh * qv[kk].z; dk_acc[rp][kk].w += dsh * qv[kk].w;
            dv_acc[rp][kk].x += ah * g[kk].x; dv_acc[rp][kk].y += ah * g[kk].y;
            dv_acc[rp][kk].z += ah * g[kk].z; dv_acc[rp][kk].w += ah * g[kk].w;
          }
        }
      }
    }
    float4* dqr = reinterpret_cast<float4*>(dq + (static_cast<long long>(b) * N + n) * C4 * 4);
#pragma unroll
    for (int kk = 0; kk < GATE_MAX_V4; ++kk) {
      const int f = lane + 32 * kk;
      if (f < C4) dqr[f] = dqv[kk];
    }
  }
  // block reduce of dk/dv through shared memory, then one atomic per element per block
#pragma unroll
  for (int rp = 0; rp < GATE_MAX_REP; ++rp) {
    if (rp < rep) {
#pragma unroll
      for (int kk = 0; kk < GATE_MAX_V4; ++kk) {
        const int f = lane + 32 * kk;
        if (f < C4) {
          float* ak = acc_sh + ((rp * 2 + 0) * C4 + f) * 4;
          float* av = acc_sh + ((rp * 2 + 1) * C4 + f) * 4;
          atomicAdd(ak + 0, dk_acc[rp][kk].x); atomicAdd(ak + 1, dk_acc[rp][kk].y);
          atomicAdd(ak + 2, dk_acc[rp][kk].z); atomicAdd(ak + 3, dk_acc[rp][kk].w);
          atomicAdd(av + 0, dv_acc[rp][kk].x); atomicAdd(av + 1, dv_acc[rp][kk].y);
          atomicAdd(av + 2, dv_acc[rp][kk].z); atomicAdd(av + 3, dv_acc[rp][kk].w);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rep * 2 * C4 * 4; i += blockDim.x) {
    const int rp = i / (2 * C4 * 4), which = (i / (C4 * 4)) & 1, ch = i % (C4 * 4);
    const long long r = b + static_cast<long long>(rp) * Bq;
    atomicAdd((which ? dv : dk) + r * C4 * 4 + ch, acc_sh[i]);
  }
}

// ------------------------------------------------------------------------------------------------ GELU backward
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dx,
                                long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 d = reinterpret_cast<const float4*>(dy)[i];
    const float4 x = reinterpret_cast<const float4*>(pre)[i];
    auto f = [](float xv) {
      const float cdf = 0.5f * (1.f + erff(xv * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * expf(-0.5f * xv * xv);
      return cdf + xv * pdf;
    };
    reinterpret_cast<float4*>(dx)[i] = make_float4(d.x * f(x.x), d.y * f(x.y), d.z * f(x.z), d.w * f(x.w));
  }
}

}  // namespace cavp

using namespace cavp;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int cavp_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                  float* rstd, long long T, int C, float eps, void* stream) {
  if ((C & 3) || C / 4 > 32 * LN_MAX_V4) return CAVP_ERR_ARG;
  layernorm_fwd_kernel<<<grid_for(T, 8), 256, 0, ST(stream)>>>(x, gamma, beta, y, mean, rstd, T, C / 4, eps);
  CAVP_LAUNCH_CHECK();
}
// partials must hold cavp_layernorm_bwd_nparts(T) * 2 * C floats
extern "C" int cavp_layernorm_bwd_nparts(long long T) { return grid_for(T, 8, 4) * 8; }
extern "C" int cavp_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                                  const float* rstd, const float* add, float* dx, float* partials, long long T, int C,
                                  void* stream) {
  if ((C & 3) || C / 4 > 32 * LN_MAX_V4) return CAVP_ERR_ARG;
  layernorm_bwd_kernel<<<grid_for(T, 8, 4), 256, 0, ST(stream)>>>(dy, x, gamma, mean, rstd, add, dx, partials, T,
                                                                   C / 4);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_gate_fwd(const float* q, const float* k, const float* v, float* x, float* attn, int Bq, int rep,
                             int N, int C, int heads, void* stream) {
  if (heads != GATE_HEADS || (C & 3) || (C / heads) % 4 || C / 4 > 32 * GATE_MAX_V4 || rep < 1 || rep > GATE_MAX_REP)
    return CAVP_ERR_ARG;
  const int C4 = C / 4, D4 = C / heads / 4;
  int gx = (N + 7) / 8;
  const int cap = (NUM_SMS * 8 + Bq - 1) / Bq;
  if (gx > cap) gx = cap;
  dim3 grid(gx, Bq);
  const float scale = 1.0f / sqrtf(static_cast<float>(C / heads));
  gate_fwd_kernel<<<grid, 256, rep * 2 * C4 * sizeof(float4), ST(stream)>>>(q, k, v, x, attn, Bq, rep, N, C4, D4, scale);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_gate_bwd(const float* dx, const float* q, const float* k, const float* v, const float* attn,
                             float* dq, float* dk, float* dv, int Bq, int rep, int N, int C, int heads, void* stream) {
  if (heads != GATE_HEADS || (C & 3) || (C / heads) % 4 || C / 4 > 32 * GATE_MAX_V4 || rep < 1 || rep > GATE_MAX_REP)
    return CAVP_ERR_ARG;
  const int C4 = C / 4, D4 = C / heads / 4;
  int gx = (N + 7) / 8;
  const int cap = (NUM_SMS * 4 + Bq - 1) / Bq;
  if (gx > cap) gx = cap;
  dim3 grid(gx, Bq);
  const float scale = 1.0f / sqrtf(static_cast<float>(C / heads));
  gate_bwd_kernel<<<grid, 256, 2 * rep * 2 * C4 * sizeof(float4), ST(stream)>>>(dx, q, k, v, attn, dq, dk, dv, Bq, rep,
                                                                                 N, C4, D4, scale);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_gelu_bwd(const float* dy, const float* pre, float* dx, long long n, void* stream) {
  if (n & 3) return CAVP_ERR_ALIGN;
  gelu_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, ST(stream)>>>(dy, pre, dx, n / 4);
  CAVP_LAUNCH_CHECK();
}
