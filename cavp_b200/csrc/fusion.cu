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
// NV4 = float4 slots per lane: 3 covers C <= 384 at ~70 registers (the generic 10-slot instance needs 205 registers = one
// block per SM; measured 1.07 TB/s on the C = 304 fusion block)
template <int NV4>
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, const float* __restrict__ add,
                                     float* __restrict__ dx, float* __restrict__ partials, long long T, int C4) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float invC = 1.f / (C4 * 4);
  float4 dg[NV4], db[NV4];
#pragma unroll
  for (int k = 0; k < NV4; ++k) dg[k] = db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long t = warp0; t < T; t += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + t * C4 * 4);
    const float4* gr = reinterpret_cast<const float4*>(dy + t * C4 * 4);
    const float mu = mean[t], rs = rstd[t];
    float4 xh[NV4], g[NV4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
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
    for (int k = 0; k < NV4; ++k) {
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
  for (int k = 0; k < NV4; ++k) {
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

// Lane mapping: head h <-> lanes 8h..8h+7 (4 heads x 8 lanes); a lane owns float4 chunks l8, l8+8, l8+16 of its head's
// D/4 chunks, so a per-head dot product is a 3-step shuffle reduction inside the quarter-warp, and every load / store
// instruction touches four contiguous 128-byte segments.  Two tokens are processed per iteration to keep more loads
// in flight (the kernel is a pure HBM stream: 1 read + rep writes of N*C floats).
__global__ void __launch_bounds__(256) gate_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                        const float* __restrict__ v, float* __restrict__ x,
                                                        float* __restrict__ attn, int Bq, int rep, int N, int C4,
                                                        int D4, float scale) {
  extern __shared__ float4 kv_sh[];  // [rep][2][C4]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < rep * 2 * C4; i += blockDim.x) {
    const int rp = i / (2 * C4), which = (i / C4) & 1, f = i % C4;
    const int r = b + rp * Bq;
    kv_sh[i] = reinterpret_cast<const float4*>((which ? v : k) + static_cast<long long>(r) * C4 * 4)[f];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int h = lane >> 3, l8 = lane & 7;
  int fidx[GATE_MAX_V4];
  bool fok[GATE_MAX_V4];
#pragma unroll
  for (int j = 0; j < GATE_MAX_V4; ++j) {
    fok[j] = l8 + 8 * j < D4;
    fidx[j] = h * D4 + (fok[j] ? l8 + 8 * j : 0);
  }
  constexpr int TOK = 2;
  for (int n0 = (blockIdx.x * nw + warp) * TOK; n0 < N; n0 += gridDim.x * nw * TOK) {
    float4 qv[TOK][GATE_MAX_V4];
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
      const int n = n0 + t < N ? n0 + t : N - 1;
      const float4* qr = reinterpret_cast<const float4*>(q + (static_cast<long long>(b) * N + n) * C4 * 4);
#pragma unroll
      for (int j = 0; j < GATE_MAX_V4; ++j) qv[t][j] = fok[j] ? __ldg(qr + fidx[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int rp = 0; rp < rep; ++rp) {
      const float4* ks = kv_sh + rp * 2 * C4;
      const float4* vs = ks + C4;
      const long long r = b + static_cast<long long>(rp) * Bq;
      float4 kf[GATE_MAX_V4], vf[GATE_MAX_V4];
#pragma unroll
      for (int j = 0; j < GATE_MAX_V4; ++j) {
        kf[j] = ks[fidx[j]];
        vf[j] = vs[fidx[j]];
      }
#pragma unroll
      for (int t = 0; t < TOK; ++t) {
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < GATE_MAX_V4; ++j)
          if (fok[j]) d += qv[t][j].x * kf[j].x + qv[t][j].y * kf[j].y + qv[t][j].z * kf[j].z + qv[t][j].w * kf[j].w;
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        const float a = 1.f / (1.f + expf(-d * scale));
        if (n0 + t < N) {
          float4* xr = reinterpret_cast<float4*>(x + (r * N + n0 + t) * C4 * 4);
#pragma unroll
          for (int j = 0; j < GATE_MAX_V4; ++j)
            if (fok[j]) xr[fidx[j]] = make_float4(a * vf[j].x, a * vf[j].y, a * vf[j].z, a * vf[j].w);
          if (l8 == 0) attn[(r * GATE_HEADS + h) * N + n0 + t] = a;
        }
      }
    }
  }
}

// backward: dq (summed over the rows that share q), dk / dv accumulated with atomics (pre-zeroed [rows][C]).
__global__ void __launch_bounds__(256, 2) gate_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ q,
                                                        const float* __restrict__ k, const float* __restrict__ v,
                                                        const float* __restrict__ attn, float* __restrict__ dq,
                                                        float* __restrict__ dk, float* __restrict__ dv, int Bq, int rep,
                                                        int N, int C4, int D4, float scale) {
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
  const int h = lane >> 3, l8 = lane & 7;  // same quarter-warp-per-head mapping as the forward kernel
  int fidx[GATE_MAX_V4];
  bool fok[GATE_MAX_V4];
#pragma unroll
  for (int j = 0; j < GATE_MAX_V4; ++j) {
    fok[j] = l8 + 8 * j < D4;
    fidx[j] = h * D4 + (fok[j] ? l8 + 8 * j : 0);
  }
  float4 dk_acc[GATE_MAX_REP][GATE_MAX_V4], dv_acc[GATE_MAX_REP][GATE_MAX_V4];
#pragma unroll
  for (int rp = 0; rp < GATE_MAX_REP; ++rp)
#pragma unroll
    for (int j = 0; j < GATE_MAX_V4; ++j) dk_acc[rp][j] = dv_acc[rp][j] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int n = blockIdx.x * nw + warp; n < N; n += gridDim.x * nw) {
    const float4* qr = reinterpret_cast<const float4*>(q + (static_cast<long long>(b) * N + n) * C4 * 4);
    float4 qv[GATE_MAX_V4], dqv[GATE_MAX_V4], g[GATE_MAX_REP][GATE_MAX_V4];
#pragma unroll
    for (int j = 0; j < GATE_MAX_V4; ++j) {
      qv[j] = fok[j] ? __ldg(qr + fidx[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
      dqv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int rp = 0; rp < GATE_MAX_REP; ++rp) {
      if (rp < rep) {
        const long long r = b + static_cast<long long>(rp) * Bq;
        const float4* gr = reinterpret_cast<const float4*>(dx + (r * N + n) * C4 * 4);
#pragma unroll
        for (int j = 0; j < GATE_MAX_V4; ++j) g[rp][j] = fok[j] ? __ldg(gr + fidx[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int rp = 0; rp < GATE_MAX_REP; ++rp) {
      if (rp < rep) {
        const float4* ks = kv_sh + rp * 2 * C4;
        const float4* vs = ks + C4;
        const long long r = b + static_cast<long long>(rp) * Bq;
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < GATE_MAX_V4; ++j) {
          const float4 vf = vs[fidx[j]];
          d += g[rp][j].x * vf.x + g[rp][j].y * vf.y + g[rp][j].z * vf.z + g[rp][j].w * vf.w;
        }
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        const float a = attn[(r * GATE_HEADS + h) * N + n];
        const float ds = d * a * (1.f - a) * scale;
#pragma unroll
        for (int j = 0; j < GATE_MAX_V4; ++j) {
          const float4 kf = ks[fidx[j]];
          dqv[j].x += ds * kf.x; dqv[j].y += ds * kf.y; dqv[j].z += ds * kf.z; dqv[j].w += ds * kf.w;
          dk_acc[rp][j].x += ds * qv[j].x; dk_acc[rp][j].y += ds * qv[j].y;
          dk_acc[rp][j].z += ds * qv[j].z; dk_acc[rp][j].w += ds * qv[j].w;
          dv_acc[rp][j].x += a * g[rp][j].x; dv_acc[rp][j].y += a * g[rp][j].y;
          dv_acc[rp][j].z += a * g[rp][j].z; dv_acc[rp][j].w += a * g[rp][j].w;
        }
      }
    }
    float4* dqr = reinterpret_cast<float4*>(dq + (static_cast<long long>(b) * N + n) * C4 * 4);
#pragma unroll
    for (int j = 0; j < GATE_MAX_V4; ++j)
      if (fok[j]) dqr[fidx[j]] = dqv[j];
  }
  // block reduce of dk/dv through shared memory, then one atomic per element per block
#pragma unroll
  for (int rp = 0; rp < GATE_MAX_REP; ++rp) {
    if (rp < rep) {
#pragma unroll
      for (int j = 0; j < GATE_MAX_V4; ++j) {
        if (fok[j]) {
          float* ak = acc_sh + ((rp * 2 + 0) * C4 + fidx[j]) * 4;
          float* av = acc_sh + ((rp * 2 + 1) * C4 + fidx[j]) * 4;
          atomicAdd(ak + 0, dk_acc[rp][j].x); atomicAdd(ak + 1, dk_acc[rp][j].y);
          atomicAdd(ak + 2, dk_acc[rp][j].z); atomicAdd(ak + 3, dk_acc[rp][j].w);
          atomicAdd(av + 0, dv_acc[rp][j].x); atomicAdd(av + 1, dv_acc[rp][j].y);
          atomicAdd(av + 2, dv_acc[rp][j].z); atomicAdd(av + 3, dv_acc[rp][j].w);
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

// ------------------------------------------------------------------------------------------------ GELU forward
// y = x * Phi(x) (nn.GELU, exact erf form: timm Mlp act_layer, models/attn.py:138-143, cavp_model.py:123-128).  Runs as its
// own pass over the fc1 output: 64 erff per thread inside the GEMM epilogue cost more than the whole K = 304 main loop
// (one warp per scheduler, nothing to hide the dependent FMA chain), here 2048 threads per SM hide it behind HBM.
__global__ void gelu_fwd_kernel(const float* __restrict__ pre, float* __restrict__ y, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(pre)[i];
    auto f = [](float xv) { return 0.5f * xv * (1.f + erff(xv * 0.70710678118654752440f)); };
    reinterpret_cast<float4*>(y)[i] = make_float4(f(x.x), f(x.y), f(x.z), f(x.w));
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
  if (C / 4 <= 32 * 3)
    layernorm_bwd_kernel<3><<<grid_for(T, 8, 4), 256, 0, ST(stream)>>>(dy, x, gamma, mean, rstd, add, dx, partials, T,
                                                                        C / 4);
  else
    layernorm_bwd_kernel<LN_MAX_V4><<<grid_for(T, 8, 4), 256, 0, ST(stream)>>>(dy, x, gamma, mean, rstd, add, dx,
                                                                                partials, T, C / 4);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_gate_fwd(const float* q, const float* k, const float* v, float* x, float* attn, int Bq, int rep,
                             int N, int C, int heads, void* stream) {
  if (heads != GATE_HEADS || (C & 3) || (C / heads) % 4 || C / 4 > 32 * GATE_MAX_V4 || rep < 1 || rep > GATE_MAX_REP)
    return CAVP_ERR_ARG;
  const int C4 = C / 4, D4 = C / heads / 4;
  if ((C / heads / 4 + 7) / 8 > GATE_MAX_V4) return CAVP_ERR_ARG;
  int gx = (N + 15) / 16;
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
extern "C" int cavp_gelu_fwd(const float* pre, float* y, long long n, void* stream) {
  if (!pre || !y) return CAVP_ERR_NULL;
  if ((n & 3) || (reinterpret_cast<uintptr_t>(pre) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return CAVP_ERR_ALIGN;
  gelu_fwd_kernel<<<grid_for(n / 4, 256), 256, 0, ST(stream)>>>(pre, y, n / 4);
  CAVP_LAUNCH_CHECK();
}
extern "C" int cavp_gelu_bwd(const float* dy, const float* pre, float* dx, long long n, void* stream) {
  if (n & 3) return CAVP_ERR_ALIGN;
  gelu_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, ST(stream)>>>(dy, pre, dx, n / 4);
  CAVP_LAUNCH_CHECK();
}
