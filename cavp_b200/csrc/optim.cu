// Fused optimiser steps over a table of tensors.
//
// The reference builds torch.optim.SGD(momentum, weight_decay) over 12 visual parameter groups and torch.optim.Adam over
// the audio backbone (main_vpo_mono.py:118-125) and calls .step() on both every iteration
// (trainer/trainer_cavp_vpo_mono.py:192-193).  Here ONE launch per optimiser walks every parameter of every group: the
// host uploads a table {param, grad, state pointers, length, lr, weight decay} plus a fixed (tensor, chunk) work list;
// a block owns 16 K-element chunks, so the 50 M-element VGG FC weight and a 64-element BN bias share the grid evenly.
// HBM-bound: SGD reads p, g, buf and writes p, buf (20 B / element); Adam reads p, g, m, v, writes p, m, v (28 B).
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../include/cavp_b200.h"
#include "common.cuh"

namespace cavp {

struct OptTensor {  // mirrors the int64[6] row the host packs (cavp_b200/optim.py)
  float* p;
  const float* g;
  float* s0;  // SGD momentum buffer / Adam exp_avg
  float* s1;  // Adam exp_avg_sq
  long long n;
  float lr;
  float wd;
};
static_assert(sizeof(OptTensor) == 48, "table row layout");

constexpr int OPT_CHUNK = 16384;  // elements per work item
constexpr int OPT_THREADS = 256;

struct SgdOp {
  float momentum;
  __device__ __forceinline__ void operator()(float& p, float g, float& buf, float&, float lr, float wd) const {
    const float gp = fmaf(wd, p, g);                       // grad.add(param, alpha=weight_decay)
    buf = __fadd_rn(__fmul_rn(buf, momentum), gp);         // buf.mul_(momentum).add_(grad)   (first step: buf = 0)
    p = fmaf(-lr, buf, p);                                 // param.add_(buf, alpha=-lr)
  }
};
struct AdamOp {
  float beta2, omb1, omb2, eps, inv_bc1, inv_sqrt_bc2;  // omb = 1 - beta, rounded from the host's double like torch does
  __device__ __forceinline__ void operator()(float& p, float g, float& m, float& v, float lr, float wd) const {
    g = fmaf(wd, p, g);
    m = fmaf(omb1, g - m, m);                               // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(omb2, g * g, __fmul_rn(v, beta2));            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = __fadd_rn(__fmul_rn(sqrtf(v), inv_sqrt_bc2), eps);
    p = fmaf(-(lr * inv_bc1), __fdiv_rn(m, denom), p);      // param.addcdiv_(exp_avg, denom, value=-lr/bc1)
  }
};

template <class Op, bool HAS_S1>
__global__ void __launch_bounds__(OPT_THREADS)
opt_multi_kernel(const OptTensor* __restrict__ table, const int2* __restrict__ work, int nwork, Op op) {
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int2 w = work[wi];
    const OptTensor t = table[w.x];
    if (t.g == nullptr) continue;  // parameter without a gradient this step (torch skips it too)
    const long long off = static_cast<long long>(w.y) * OPT_CHUNK;
    const int n = static_cast<int>(t.n - off < OPT_CHUNK ? t.n - off : OPT_CHUNK);
    float* p = t.p + off;
    const float* g = t.g + off;
    float* s0 = t.s0 + off;
    float* s1 = HAS_S1 ? t.s1 + off : nullptr;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(s0) | (HAS_S1 ? reinterpret_cast<uintptr_t>(s1) : 0);
    const int n4 = (al & 15) == 0 ? (n >> 2) : 0;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 pv = reinterpret_cast<float4*>(p)[i];
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 a = reinterpret_cast<float4*>(s0)[i];
      float4 b = HAS_S1 ? reinterpret_cast<float4*>(s1)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      op(pv.x, gv.x, a.x, b.x, t.lr, t.wd);
      op(pv.y, gv.y, a.y, b.y, t.lr, t.wd);
      op(pv.z, gv.z, a.z, b.z, t.lr, t.wd);
      op(pv.w, gv.w, a.w, b.w, t.lr, t.wd);
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(s0)[i] = a;
      if (HAS_S1) reinterpret_cast<float4*>(s1)[i] = b;
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += OPT_THREADS) {
      float pv = p[i], a = s0[i], b = HAS_S1 ? s1[i] : 0.f;
      op(pv, g[i], a, b, t.lr, t.wd);
      p[i] = pv;
      s0[i] = a;
      if (HAS_S1) s1[i] = b;
    }
  }
}

// ---- TF32 hi | lo split of every weight operand of the model in ONE launch (the per-weight cavp_split_tf32 launches
// were launch-bound: ~230 kernels of a few microseconds per step).  Table rows {src, hi, lo, n}; same work list format.
struct SplitTensor {
  const float* src;
  float* hi;
  float* lo;
  long long n;
};
static_assert(sizeof(SplitTensor) == 32, "table row layout");

__device__ __forceinline__ float tf32_rn_(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(OPT_THREADS)
split_multi_kernel(const SplitTensor* __restrict__ table, const int2* __restrict__ work, int nwork) {
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int2 w = work[wi];
    const SplitTensor t = table[w.x];
    const long long off = static_cast<long long>(w.y) * OPT_CHUNK;
    const int n = static_cast<int>(t.n - off < OPT_CHUNK ? t.n - off : OPT_CHUNK);
    const float* src = t.src + off;
    float* hi = t.hi + off;
    float* lo = t.lo + off;
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo);
    const int n4 = (al & 15) == 0 ? (n >> 2) : 0;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      const float h0 = tf32_rn_(v.x), h1 = tf32_rn_(v.y), h2 = tf32_rn_(v.z), h3 = tf32_rn_(v.w);
      reinterpret_cast<float4*>(hi)[i] = make_float4(h0, h1, h2, h3);
      reinterpret_cast<float4*>(lo)[i] =
          make_float4(tf32_rn_(v.x - h0), tf32_rn_(v.y - h1), tf32_rn_(v.z - h2), tf32_rn_(v.w - h3));
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += OPT_THREADS) {
      const float v = src[i];
      const float h = tf32_rn_(v);
      hi[i] = h;
      lo[i] = tf32_rn_(v - h);
    }
  }
}

// ---- many small device-to-device copies in ONE launch (gradients that were not written straight into the flat
// all-reduce buffer: BatchNorm / LayerNorm / bias gradients, re-packed stems).  Table rows {src, dst, n}.
struct CopyTensor {
  const float* src;
  float* dst;
  long long n;
};
static_assert(sizeof(CopyTensor) == 24, "table row layout");

__global__ void __launch_bounds__(OPT_THREADS)
copy_multi_kernel(const CopyTensor* __restrict__ table, const int2* __restrict__ work, int nwork) {
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int2 w = work[wi];
    const CopyTensor t = table[w.x];
    const long long off = static_cast<long long>(w.y) * OPT_CHUNK;
    const int n = static_cast<int>(t.n - off < OPT_CHUNK ? t.n - off : OPT_CHUNK);
    const float* src = t.src + off;
    float* dst = t.dst + off;
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
    const int n4 = (al & 15) == 0 ? (n >> 2) : 0;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS)
      reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    for (int i = n4 * 4 + threadIdx.x; i < n; i += OPT_THREADS) dst[i] = src[i];
  }
}

// fp32 -> bf16 copies of every weight operand in one launch (bf16 path of BASELINE.json configs[2]); rows {src, dst, n}
__global__ void __launch_bounds__(OPT_THREADS)
cvt_bf16_multi_kernel(const CopyTensor* __restrict__ table, const int2* __restrict__ work, int nwork) {
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int2 w = work[wi];
    const CopyTensor t = table[w.x];
    const long long off = static_cast<long long>(w.y) * OPT_CHUNK;
    const int n = static_cast<int>(t.n - off < OPT_CHUNK ? t.n - off : OPT_CHUNK);
    const float* src = t.src + off;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(t.dst) + off;
    const bool al = ((reinterpret_cast<uintptr_t>(src) & 15) | (reinterpret_cast<uintptr_t>(dst) & 7)) == 0;
    const int n4 = al ? (n >> 2) : 0;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      reinterpret_cast<uint2*>(dst)[i] = o;
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += OPT_THREADS) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

static int opt_grid(int nwork) {
  const int cap = NUM_SMS * 8;
  return nwork < cap ? (nwork < 1 ? 1 : nwork) : cap;
}

}  // namespace cavp

using namespace cavp;

extern "C" int cavp_opt_chunk_elems(void) { return OPT_CHUNK; }

extern "C" int cavp_sgd_multi(const void* table, const int* work, int nwork, float momentum, void* stream) {
  if (!table || !work) return CAVP_ERR_NULL;
  if (nwork <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(table) & 7) || (reinterpret_cast<uintptr_t>(work) & 7)) return CAVP_ERR_ALIGN;
  SgdOp op{momentum};
  opt_multi_kernel<SgdOp, false><<<opt_grid(nwork), OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const OptTensor*>(table), reinterpret_cast<const int2*>(work), nwork, op);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_adam_multi(const void* table, const int* work, int nwork, double beta1, double beta2, double eps,
                               double bias_correction1, double bias_correction2, void* stream) {
  if (!table || !work) return CAVP_ERR_NULL;
  if (nwork <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(table) & 7) || (reinterpret_cast<uintptr_t>(work) & 7)) return CAVP_ERR_ALIGN;
  if (!(bias_correction1 > 0.0) || !(bias_correction2 > 0.0)) return CAVP_ERR_ARG;
  AdamOp op{static_cast<float>(beta2), static_cast<float>(1.0 - beta1), static_cast<float>(1.0 - beta2),
            static_cast<float>(eps), static_cast<float>(1.0 / bias_correction1),
            static_cast<float>(1.0 / sqrt(bias_correction2))};
  opt_multi_kernel<AdamOp, true><<<opt_grid(nwork), OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const OptTensor*>(table), reinterpret_cast<const int2*>(work), nwork, op);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_split_tf32_multi(const void* table, const int* work, int nwork, void* stream) {
  if (!table || !work) return CAVP_ERR_NULL;
  if (nwork <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(table) & 7) || (reinterpret_cast<uintptr_t>(work) & 7)) return CAVP_ERR_ALIGN;
  split_multi_kernel<<<opt_grid(nwork), OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const SplitTensor*>(table), reinterpret_cast<const int2*>(work), nwork);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_copy_multi(const void* table, const int* work, int nwork, void* stream) {
  if (!table || !work) return CAVP_ERR_NULL;
  if (nwork <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(table) & 7) || (reinterpret_cast<uintptr_t>(work) & 7)) return CAVP_ERR_ALIGN;
  copy_multi_kernel<<<opt_grid(nwork), OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const CopyTensor*>(table), reinterpret_cast<const int2*>(work), nwork);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_cvt_bf16_multi(const void* table, const int* work, int nwork, void* stream) {
  if (!table || !work) return CAVP_ERR_NULL;
  if (nwork <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(table) & 7) || (reinterpret_cast<uintptr_t>(work) & 7)) return CAVP_ERR_ALIGN;
  cvt_bf16_multi_kernel<<<opt_grid(nwork), OPT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const CopyTensor*>(table), reinterpret_cast<const int2*>(work), nwork);
  return static_cast<int>(cudaGetLastError());
}
