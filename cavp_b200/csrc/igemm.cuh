// Implicit-GEMM convolution / linear kernels on tcgen05 (sm_100a), fp32 in HBM, TF32 tensor-core products.
//
// Data layout: activations are NHWC with an explicit pixel stride `ld` (so channel slices of a concat buffer are
// ordinary operands); weights are [Cout][R][S][Cin] (= K-major rows of length K = R*S*Cin).
//
//   MODE_ROW  : D[M x N] = A[M x K] * B[N x K]^T, both operands K-major in shared memory.  A is gathered on the fly
//               (im2col for forward, the transposed-stride gather for dgrad), B is the weight matrix.
//               Forward conv / linear / dgrad (with pre-transposed weights) all run through it.
//   MODE_WGRAD: dW[Cout x K] = dY[P x Cout]^T * im2col(X)[P x K]; the reduction runs over pixels, so both operands are
//               MN-major in shared memory (no transposes in HBM).
//
// CTA = 8 producer/epilogue warps + 1 MMA warp.  The two producer groups (4 warps each) alternate k-blocks:
// global -> registers -> TF32 hi/lo split -> swizzled shared memory -> fence.proxy.async -> mbarrier.  One elected
// thread issues tcgen05.mma (M=128, N=BN, K=8) into TMEM; tcgen05.commit releases the stage.
//
// Precision (DESIGN.md "precision"): PREC=1 is plain TF32.  PREC=2 is the fp32-parity mode: three MMAs per k-step
// (hi*hi + lo*hi + hi*lo) AND accumulator promotion - the tensor core adds into TMEM with truncation, which biases
// long accumulation chains (measured: ~8e-9*K relative), so every 2 k-blocks (64 K-elements) the partial sum is
// pulled out of a 4-deep TMEM ring with tcgen05.ld and added, round-to-nearest, into fp32 registers by the
// producer warps while their next global loads are in flight.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace cavp {

constexpr int BM = 128;        // tile rows = TMEM lanes
constexpr int BK = 32;         // fp32 per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;      // tf32
constexpr int PRODUCER_THREADS = 256;
constexpr int GROUP_THREADS = 128;
constexpr int CTA_THREADS = PRODUCER_THREADS + 32;
constexpr int MODE_ROW = 0;
constexpr int MODE_WGRAD = 1;

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_GELU = 3, ACT_SIGMOID = 4 };

struct IgemmParams {
  // MODE_ROW: x = A source (NHWC, pixel stride ldx), w = B [Ncols][ldw] K-major, y = D [M][ldy]
  // MODE_WGRAD: x = NHWC source of the im2col operand, w = dY [P][ldw] (Cout columns), y = dW [Cout][K]
  const float* x;
  const float* w;
  float* y;
  float* y_pre;        // optional copy of D before the activation (same ld)
  const float* scale;  // optional per-column scale (eval-mode BN)
  const float* shift;  // optional per-column shift / bias
  const float* res;    // optional residual [res_mod or M][ldr], added before the activation
  float* stats;        // optional per-(m_tile, warp-quarter) column partial sums: [m_tiles*4][2][ldstat]
  int Nimg, Hs, Ws, C, ldx;
  int Ho, Wo;
  int R, S, stride, pad, dil;
  int dgrad;
  int M, Ncols, K, ldw, ldy, ldr, res_mod, res_div, ldstat;  // residual row = (row / res_div) % res_mod
  int red_len;  // length of the reduction dimension (K for MODE_ROW, P pixels for MODE_WGRAD)
  int act;
  float slope;
  int n_tiles, num_kb, splits;
  int unit_kb;         // k-blocks per promotion unit of the CTA-pair kernel (igemm_ws2.cuh); 2 unless set by the dispatcher
  long long split_slab;  // > 0: deterministic split-K - split i stores its partial product at y + i*split_slab (no atomics)
  FastDiv div_howo, div_wo, div_c, div_s;
};

template <int BN, int PREC>
struct TileCfg {
  static constexpr bool PROMOTE = (PREC == 2);
  static constexpr int NBUF = PROMOTE ? 4 : 1;
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * PREC;
  static constexpr int STAGES = (PREC == 2) ? (BN >= 128 ? 3 : 4) : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*row table*/;
  static constexpr int TMEM_COLS = NBUF * BN < 32 ? 32 : NBUF * BN;
  static constexpr int HALF = BN / 2;  // accumulator columns owned by one epilogue thread
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_LEAKY: return v > 0.f ? v : v * slope;
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// variant for the issue-bound wgrad producers: the lo part is stored unrounded (the tensor core truncates it to TF32,
// |error| <= 2^-21 |x|, sign-symmetric because hi is round-to-nearest) - 3 instead of 5 ALU ops per element
template <int PREC>
__device__ __forceinline__ void store_split_fast(uint32_t hi_addr, uint32_t lo_addr, const float4& v) {
  const float h0 = tf32_rn(v.x), h1 = tf32_rn(v.y), h2 = tf32_rn(v.z), h3 = tf32_rn(v.w);
  st_shared_v4(hi_addr, h0, h1, h2, h3);
  if (PREC == 2) st_shared_v4(lo_addr, v.x - h0, v.y - h1, v.z - h2, v.w - h3);
}
template <int PREC>
__device__ __forceinline__ void store_split(uint32_t hi_addr, uint32_t lo_addr, const float4& v) {
  const float h0 = tf32_rn(v.x), h1 = tf32_rn(v.y), h2 = tf32_rn(v.z), h3 = tf32_rn(v.w);
  st_shared_v4(hi_addr, h0, h1, h2, h3);
  if (PREC == 2) st_shared_v4(lo_addr, tf32_rn(v.x - h0), tf32_rn(v.y - h1), tf32_rn(v.z - h2), tf32_rn(v.w - h3));
}

// ---- epilogue building blocks.  A warp owns a 32 (rows = lanes) x 32 (columns = registers) chunk of the tile and moves
// it through a warp-private shared-memory scratch (row stride 36 floats: 16-byte aligned, conflict-free for
// quarter-warp accesses) so that every global access is four full 128-byte row segments instead of 32 scattered pieces.
constexpr int EPI_LDS = 36;

__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 t;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(addr));
  return t;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float t;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(addr));
  return t;
}
// registers -> scratch[lane][0..31]
__device__ __forceinline__ void stage_chunk(uint32_t scratch, const float (&v)[32], int lane) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    st_shared_v4(scratch + static_cast<uint32_t>((lane * EPI_LDS + j) * 4), v[j], v[j + 1], v[j + 2], v[j + 3]);
  __syncwarp();
}
// scratch -> global rows row0 .. row0+31 (those < M), 32 columns starting at ybase
__device__ __forceinline__ void store_staged(uint32_t scratch, float* ybase, int ldy, int row0, int M, int lane) {
  const int c4 = lane & 7, rsub = lane >> 3;
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const int r = rr * 4 + rsub;
    const float4 t = lds_v4(scratch + static_cast<uint32_t>((r * EPI_LDS + c4 * 4) * 4));
    if (row0 + r < M) *reinterpret_cast<float4*>(ybase + static_cast<size_t>(r) * ldy + c4 * 4) = t;
  }
}
// store_staged with the per-column part of the epilogue applied on the way out: each lane owns ONE column quad, so
// scale / shift are one 16-byte load per lane and chunk instead of 32 scalar loads per thread in the register phase
__device__ __forceinline__ void store_staged_affine(uint32_t scratch, float* ybase, int ldy, int row0, int M, int lane,
                                                    const float* scale, const float* shift, int act, float slope) {
  const int c4 = lane & 7, rsub = lane >> 3;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
  if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const int r = rr * 4 + rsub;
    float4 t = lds_v4(scratch + static_cast<uint32_t>((r * EPI_LDS + c4 * 4) * 4));
    t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y); t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
    if (act == ACT_RELU) {
      t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
    } else if (act == ACT_LEAKY) {
      t.x = t.x > 0.f ? t.x : t.x * slope; t.y = t.y > 0.f ? t.y : t.y * slope;
      t.z = t.z > 0.f ? t.z : t.z * slope; t.w = t.w > 0.f ? t.w : t.w * slope;
    }
    if (row0 + r < M) *reinterpret_cast<float4*>(ybase + static_cast<size_t>(r) * ldy + c4 * 4) = t;
  }
}
// v[j] += res[res_row(row0 + lane)][col0 + j] for a full 32 x 32 chunk: coalesced loads, transposed through the scratch
__device__ __forceinline__ void add_chunk_coalesced(uint32_t scratch, const IgemmParams& p, int row0, int col0,
                                                    float (&v)[32], int lane) {
  __syncwarp();
  const int c4 = lane & 7, rsub = lane >> 3;
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) {
    const int r = rr * 4 + rsub;
    const int rowg = row0 + r;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rowg < p.M) {
      int q_ = p.res_div > 1 ? rowg / p.res_div : rowg;
      if (p.res_mod > 0) q_ %= p.res_mod;
      t = *reinterpret_cast<const float4*>(p.res + static_cast<size_t>(q_) * p.ldr + col0 + c4 * 4);
    }
    st_shared_v4(scratch + static_cast<uint32_t>((r * EPI_LDS + c4 * 4) * 4), t.x, t.y, t.z, t.w);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 t = lds_v4(scratch + static_cast<uint32_t>((lane * EPI_LDS + j) * 4));
    v[j] += t.x;
    v[j + 1] += t.y;
    v[j + 2] += t.z;
    v[j + 3] += t.w;
  }
}

// Epilogue from the fp32 register accumulators of one thread (row = m0 + q*32 + lane, HALF columns starting at
// n0 + group*HALF): scale/shift/residual/activation, optional pre-activation copy, BN sum / sum-of-squares partials,
// or raw accumulation (split-K).  Every option is tested ONCE per chunk (uniform branches around straight-line loops):
// the plain case costs ~60 instructions per 32 x 32 chunk.
// dgrad accumulating into an existing gradient: residual == output, nothing else in the epilogue
__host__ __device__ __forceinline__ bool igemm_inplace_acc(const IgemmParams& p) {
  return p.res != nullptr && p.res == p.y && p.ldr == p.ldy && p.res_mod == 0 && p.res_div <= 1 && p.scale == nullptr &&
         p.shift == nullptr && p.y_pre == nullptr && p.stats == nullptr && p.act == ACT_NONE;
}

template <int HALF>
__device__ __forceinline__ void igemm_epilogue(const IgemmParams& pin, float (&acc)[HALF], int m0, int n0, int m_tile,
                                               int group, int q, int lane, uint32_t scratch, int split = 0) {
  if (m0 >= pin.M) return;
  // deterministic split-K: this split owns a private slab of the output and stores its raw partial product there
  // (plain epilogue, splits = 1 semantics); a fixed-order reduction over the slabs follows on the host's stream
  IgemmParams p = pin;
  if (pin.splits > 1 && pin.split_slab > 0) {
    p.y = pin.y + static_cast<size_t>(split) * pin.split_slab;
    p.splits = 1;
  }  // a pair kernel's second CTA can own a tile past the last row: nothing to store, no stats rows
  const int row = m0 + q * 32 + lane;
  const int row0 = m0 + q * 32;
  const bool row_ok = row < p.M;
  const bool vec_ok = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
  const bool inplace_acc = igemm_inplace_acc(p);
#pragma unroll
  for (int cgrp = 0; cgrp < HALF / 32; ++cgrp) {
    const int col0 = n0 + group * HALF + cgrp * 32;
    if (col0 >= p.Ncols) continue;
    const int ncol = p.Ncols - col0 < 32 ? p.Ncols - col0 : 32;  // valid columns of this chunk (warp-uniform)
    const bool full_chunk = vec_ok && ncol == 32;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = acc[cgrp * 32 + j];
    float* yrow = p.y + static_cast<size_t>(row) * p.ldy + col0;
    if (p.splits > 1 || inplace_acc) {
      // y += v with red.global.add: split-K partial products, and the in-place gradient accumulation of dgrad
      // (res == y; one writer per element, so still deterministic) without waiting for a residual read
      if (full_chunk) {
        stage_chunk(scratch, v, lane);
        const int c4 = lane & 7, rsub = lane >> 3;
        float* ybase = p.y + static_cast<size_t>(row0) * p.ldy + col0;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = rr * 4 + rsub;
          const float4 t = lds_v4(scratch + static_cast<uint32_t>((r * EPI_LDS + c4 * 4) * 4));
          if (row0 + r < p.M) red_add_v4(ybase + static_cast<size_t>(r) * p.ldy + c4 * 4, t.x, t.y, t.z, t.w);
        }
      } else if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncol) atomicAdd(yrow + j, v[j]);
      }
      continue;
    }
    if (ncol < 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j >= ncol) v[j] = 0.f;
    }
    // f(j) for the valid columns; the common full chunk carries no per-column predicate
    auto cols = [&](auto f) {
      if (ncol == 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f(j);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncol) f(j);
      }
    };
    // bias / folded-BN / ReLU only (Linear + bias, eval-mode conv): defer the per-column work to the store phase
    const bool deferred = full_chunk && (p.scale || p.shift) && !p.res && !p.y_pre && !p.stats &&
                          (p.act == ACT_NONE || p.act == ACT_RELU || p.act == ACT_LEAKY) &&
                          (!p.scale || (reinterpret_cast<uintptr_t>(p.scale + col0) & 15) == 0) &&
                          (!p.shift || (reinterpret_cast<uintptr_t>(p.shift + col0) & 15) == 0);
    if (deferred) {
      stage_chunk(scratch, v, lane);
      store_staged_affine(scratch, p.y + static_cast<size_t>(row0) * p.ldy + col0, p.ldy, row0, p.M, lane,
                          p.scale ? p.scale + col0 : nullptr, p.shift ? p.shift + col0 : nullptr, p.act, p.slope);
      continue;
    }
    if (p.scale) cols([&](int j) { v[j] *= __ldg(p.scale + col0 + j); });
    if (p.shift) cols([&](int j) { v[j] += __ldg(p.shift + col0 + j); });
    if (p.res) {
      if (full_chunk && ((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0)) {
        add_chunk_coalesced(scratch, p, row0, col0, v, lane);
      } else if (row_ok) {
        int rr_ = p.res_div > 1 ? row / p.res_div : row;
        if (p.res_mod > 0) rr_ %= p.res_mod;
        const float* rrow = p.res + static_cast<size_t>(rr_) * p.ldr + col0;
        cols([&](int j) { v[j] += __ldg(rrow + j); });
      }
    }
    if (p.y_pre) {
      if (full_chunk && (reinterpret_cast<uintptr_t>(p.y_pre) & 15) == 0) {
        stage_chunk(scratch, v, lane);
        store_staged(scratch, p.y_pre + static_cast<size_t>(row0) * p.ldy + col0, p.ldy, row0, p.M, lane);
      } else if (row_ok) {
        float* prow = p.y_pre + static_cast<size_t>(row) * p.ldy + col0;
        cols([&](int j) { prow[j] = v[j]; });
      }
    }
    switch (p.act) {
      case ACT_RELU:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        break;
      case ACT_LEAKY:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
        break;
      case ACT_GELU:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752440f));
        break;
      case ACT_SIGMOID:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
        break;
      default:
        break;
    }
    if (full_chunk || p.stats) stage_chunk(scratch, v, lane);
    if (full_chunk) {
      store_staged(scratch, p.y + static_cast<size_t>(row0) * p.ldy + col0, p.ldy, row0, p.M, lane);
    } else if (row_ok) {
      cols([&](int j) { yrow[j] = v[j]; });
    }
    if (p.stats) {
      // column sums straight from the staged chunk: lane j adds column j over the valid rows (conflict-free reads)
      const int nrow = p.M - row0 < 32 ? p.M - row0 : 32;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
      for (int r = 0; r < nrow; ++r) {
        const float t = lds_f32(scratch + static_cast<uint32_t>((r * EPI_LDS + lane) * 4));
        s1 += t;
        s2 = fmaf(t, t, s2);
      }
      if (lane < ncol) {
        float* st = p.stats + static_cast<size_t>(m_tile * 4 + q) * 2 * p.ldstat;
        st[col0 + lane] = s1;
        st[p.ldstat + col0 + lane] = s2;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <int BN, int PREC, int MODE, bool BTMA>
__global__ void __launch_bounds__(CTA_THREADS, 1)
igemm_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo) {
  // BTMA: MODE_ROW - the weight operand (B) is pre-split in HBM and fetched by TMA; MODE_WGRAD - the dY operand (A,
  // MN-major) is pre-split and fetched by TMA (four 32-column atoms per k-block), the producers only gather im2col(X)
  using Cfg = TileCfg<BN, PREC>;
  constexpr bool PROMOTE = Cfg::PROMOTE;
  constexpr int NBUF = Cfg::NBUF;
  constexpr int HALF = Cfg::HALF;
  static_assert(BN == 64 || BN == 128, "BN must be 64 or 128");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                             // [STAGES]  producers -> MMA
  uint64_t* empty_bar = bars + Cfg::STAGES;              // [STAGES]  MMA -> producers
  uint64_t* accf_bar = bars + 2 * Cfg::STAGES;           // [NBUF]    MMA -> promotion
  uint64_t* acce_bar = bars + 2 * Cfg::STAGES + NBUF;    // [NBUF]    promotion -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 2 * NBUF);
  int2* rowtab = reinterpret_cast<int2*>(smem_aligned + Cfg::STAGES * Cfg::STAGE_BYTES + 256);  // [BM] (pixel base, packed y/x)

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  const int tile = blockIdx.x;
  const int n_tile = tile % p.n_tiles;
  const int m_tile = tile / p.n_tiles;
  const int m0 = m_tile * BM;
  const int n0 = n_tile * BN;
  // split range of k-blocks (over the reduction dimension) for this CTA
  const int split = blockIdx.y;
  const int kb_begin = static_cast<int>((static_cast<long long>(p.num_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.num_kb) * (split + 1)) / p.splits);
  const int nkb = kb_end - kb_begin;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], GROUP_THREADS + (BTMA ? 1 : 0));
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], PRODUCER_THREADS);
    }
    fence_mbar_init();
  }
  if (MODE == MODE_ROW && tid < BM) {
    const int m = m0 + tid;
    int2 e = make_int2(-1, 0);
    if (m < p.M) {
      uint32_t n, rem, oy, ox;
      p.div_howo.divmod(static_cast<uint32_t>(m), n, rem);
      p.div_wo.divmod(rem, oy, ox);
      int ybase, xbase;
      if (p.dgrad) {
        ybase = static_cast<int>(oy) + p.pad;
        xbase = static_cast<int>(ox) + p.pad;
      } else {
        ybase = static_cast<int>(oy) * p.stride - p.pad;
        xbase = static_cast<int>(ox) * p.stride - p.pad;
      }
      e = make_int2(static_cast<int>(n) * p.Hs * p.Ws, ((ybase + 0x4000) << 16) | (xbase + 0x4000));
    }
    rowtab[tid] = e;
  }
  if (BTMA && tid == 32) {
    tma_prefetch_desc(&tm_b_hi);
    if (PREC == 2) tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == 8) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ===================================================== producers (+ promotion + epilogue)
    const int group = warp >> 2;
    const int gtid = tid & (GROUP_THREADS - 1);
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    float acc[HALF];
#pragma unroll
    for (int j = 0; j < HALF; ++j) acc[j] = 0.f;

    auto promote = [&](int u) {
      const int b = PROMOTE ? (u & (NBUF - 1)) : 0;
      mbar_wait(&accf_bar[b], PROMOTE ? ((u / NBUF) & 1) : 0);
      tc_fence_after();
#pragma unroll
      for (int cgrp = 0; cgrp < HALF / 16; ++cgrp) {
        float v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                      static_cast<uint32_t>(b * BN + group * HALF + cgrp * 16),
                  v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[cgrp * 16 + j] += v[j];
      }
      if (PROMOTE) {
        tc_fence_before();
        mbar_arrive(&acce_bar[b]);
      }
    };

    // ---- per-thread operand addressing
    // MODE_ROW  : chunk column c (16 B) of the 128-byte k-row; rows r0 + 16 i
    // MODE_WGRAD: chunk cc (16 B) along MN (atom = cc>>3); k-rows (pixels) rr*8 + i
    const int c = gtid & 7;
    const int r0 = gtid >> 3;
    const uint32_t swz = static_cast<uint32_t>((c ^ (r0 & 7)) << 4);
    const int cc = gtid & 31;
    const int rr = gtid >> 5;
    // wgrad per-thread constants
    int wg_co = 0, wg_dy = 0, wg_dx = 0;
    uint32_t wg_ci = 0;
    bool wg_co_ok = false, wg_j_ok = false;
    if (MODE == MODE_WGRAD) {
      wg_co = m0 + cc * 4;
      wg_co_ok = wg_co < p.M;
      const int j = n0 + cc * 4;
      wg_j_ok = (cc * 4 < BN) && j < p.Ncols;
      uint32_t tap, ky, kx;
      p.div_c.divmod(static_cast<uint32_t>(wg_j_ok ? j : 0), tap, wg_ci);
      p.div_s.divmod(tap, ky, kx);
      wg_dy = static_cast<int>(ky) * p.dil - p.pad;
      wg_dx = static_cast<int>(kx) * p.dil - p.pad;
    }

    // ---- operand movers
    // A-gather state: element offsets of this thread's 8 rows for the current filter tap (or -1 = padding / out of
    // range).  They only change when the k index crosses into the next tap, so the per-k-block work is 8 predicated
    // 16-byte loads; the tap decomposition and bounds checks run once per tap.
    int a_off[8];
    int a_k = 0, a_ci = 0, a_tap = 0;
    auto a_retap = [&]() {
      uint32_t ky, kx;
      p.div_s.divmod(static_cast<uint32_t>(a_tap), ky, kx);
      const int dy = static_cast<int>(ky) * p.dil;
      const int dx = static_cast<int>(kx) * p.dil;
      const bool kvalid = a_k < p.K;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int iy, ix;
        const int2 ri = rowtab[r0 + 16 * i];
        bool ok = kvalid && ri.x >= 0;
        const int ybase = (ri.y >> 16) - 0x4000, xbase = (ri.y & 0xFFFF) - 0x4000;
        if (p.dgrad) {
          iy = ybase - dy;
          ix = xbase - dx;
          if (p.stride > 1) {
            ok = ok && iy >= 0 && ix >= 0 && (iy % p.stride) == 0 && (ix % p.stride) == 0;
            iy /= p.stride;
            ix /= p.stride;
          }
        } else {
          iy = ybase + dy;
          ix = xbase + dx;
        }
        ok = ok && static_cast<unsigned>(iy) < static_cast<unsigned>(p.Hs) &&
             static_cast<unsigned>(ix) < static_cast<unsigned>(p.Ws);
        a_off[i] = ok ? (ri.x + iy * p.Ws + ix) * p.ldx : -1;
      }
    };
    auto a_seek = [&](int it) {  // position on k-block `it` of this CTA's split
      a_k = (kb_begin + it) * BK + c * 4;
      uint32_t tap, ci;
      p.div_c.divmod(static_cast<uint32_t>(a_k < p.K ? a_k : 0), tap, ci);
      a_tap = static_cast<int>(tap);
      a_ci = static_cast<int>(ci);
      a_retap();
    };
    auto a_advance = [&]() {  // this group's next k-block is two k-blocks further
      a_k += 2 * BK;
      a_ci += 2 * BK;
      if (a_ci >= p.C || a_k >= p.K) {
        while (a_ci >= p.C) {
          a_ci -= p.C;
          ++a_tap;
        }
        a_retap();
      }
    };
    auto load_row_a = [&](float4 (&va)[8]) {
      const float* base = p.x + a_ci;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        va[i] = a_off[i] >= 0 ? ldg_nc_v4(base + a_off[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto load_row_b = [&](int it, float4 (&vb)[8]) {
      const int k = (kb_begin + it) * BK + c * 4;
      const bool kvalid = k < p.K;
#pragma unroll
      for (int j = 0; j < BN / 16; ++j) {
        const int n = n0 + r0 + 16 * j;
        vb[j] = (kvalid && n < p.Ncols) ? ldg_nc_v4(p.w + static_cast<size_t>(n) * p.ldw + k)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto load_wgrad_ab = [&](int it, float4 (&va)[8], float4 (&vb)[8], const bool with_a) {
      // this warp's k-rows are the 8 consecutive pixels rr*8 .. rr*8+7 of the k-block: decode the first one, then walk
      // (ox, oy, image) incrementally instead of 8 divmod pairs
      const int pixb = (kb_begin + it) * BK + rr * 8;
      uint32_t n, rem, oy, ox;
      p.div_howo.divmod(static_cast<uint32_t>(pixb < p.red_len ? pixb : 0), n, rem);
      p.div_wo.divmod(rem, oy, ox);
      int base = static_cast<int>(n) * p.Hs * p.Ws;
      int y = static_cast<int>(oy) * p.stride + wg_dy, x = static_cast<int>(ox) * p.stride + wg_dx;
      const int x_wrap = p.Wo * p.stride + wg_dx, y_wrap = p.Ho * p.stride + wg_dy;
      const float* dyp = p.w + static_cast<size_t>(pixb) * p.ldw + wg_co;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool pok = pixb + i < p.red_len;
        if (with_a)
          va[i] = (pok && wg_co_ok) ? ldg_nc_v4(dyp + static_cast<size_t>(i) * p.ldw) : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool ok = pok && wg_j_ok && static_cast<unsigned>(y) < static_cast<unsigned>(p.Hs) &&
                        static_cast<unsigned>(x) < static_cast<unsigned>(p.Ws);
        vb[i] = ok ? ldg_nc_v4(p.x + static_cast<size_t>(base + y * p.Ws + x) * p.ldx + wg_ci)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
        x += p.stride;
        if (x == x_wrap) {
          x = wg_dx;
          y += p.stride;
          if (y == y_wrap) {
            y = wg_dy;
            base += p.Hs * p.Ws;
          }
        }
      }
    };
    auto load_wgrad = [&](int it, float4 (&va)[8], float4 (&vb)[8]) { load_wgrad_ab(it, va, vb, true); };
    auto store_row_a = [&](int s, const float4 (&va)[8]) {
      const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES, a_lo = a_hi + Cfg::A_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t off = static_cast<uint32_t>((r0 + 16 * i) * 128) + swz;
        store_split_fast<PREC>(a_hi + off, a_lo + off, va[i]);
      }
    };
    auto store_row_b = [&](int s, const float4 (&vb)[8]) {
      const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES * PREC, b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
      for (int j = 0; j < BN / 16; ++j) {
        const uint32_t off = static_cast<uint32_t>((r0 + 16 * j) * 128) + swz;
        store_split<PREC>(b_hi + off, b_lo + off, vb[j]);
      }
    };
    // MN-major tf32 operands use the 128B swizzle with a 32-byte base: atom = 4 k-rows x 128 B, the 32-byte chunk
    // index is XORed with (k-row & 3).
    auto store_wgrad_ab = [&](int s, const float4 (&va)[8], const float4 (&vb)[8], const bool with_a) {
      const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES, a_lo = a_hi + Cfg::A_BYTES;
      const uint32_t b_hi = a_hi + Cfg::A_BYTES * PREC, b_lo = b_hi + Cfg::B_BYTES;
      const uint32_t atom_off = static_cast<uint32_t>((cc >> 3) * 4096);
      const int c16 = cc & 7;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rr * 8 + i;
        const uint32_t off = atom_off + static_cast<uint32_t>(r * 128) +
                             static_cast<uint32_t>((((c16 >> 1) ^ (r & 3)) << 5) | ((c16 & 1) << 4));
        if (with_a) store_split_fast<PREC>(a_hi + off, a_lo + off, va[i]);
        if (cc * 4 < BN) store_split_fast<PREC>(b_hi + off, b_lo + off, vb[i]);
      }
    };
    auto store_wgrad = [&](int s, const float4 (&va)[8], const float4 (&vb)[8]) { store_wgrad_ab(s, va, vb, true); };

    const int npairs = (nkb + 1) >> 1;
    if (BTMA && MODE == MODE_WGRAD) {
      // dY (pre-split, dense [P][Cout]) arrives by TMA as four 32-channel atoms per k-block, already in the MN-major
      // 128B/32B-atom swizzle the descriptors expect; the producers gather only im2col(X), register double-buffered.
      float4 vb0[8], vb1[8], vdummy[8];
      auto body = [&](int u, float4 (&cur)[8], float4 (&nxt)[8]) {
        const int it = 2 * u + group;
        if (it + 2 < nkb) load_wgrad_ab(it + 2, vdummy, nxt, false);
        if (PROMOTE && u >= 2) promote(u - 2);
        if (it < nkb) {
          const int s = it % Cfg::STAGES;
          mbar_wait(&empty_bar[s], (((it / Cfg::STAGES) & 1) ^ 1));
          if (gtid < 32) {  // warp-uniform branch + elect: the TMA issue is not wrapped in per-instruction ELECT loops
            if (elect_one_sync()) {
              const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
              mbar_arrive_expect_tx(&full_bar[s], Cfg::A_BYTES * PREC);
#pragma unroll
              for (int atom = 0; atom < BM / 32; ++atom) {
                tma_load_2d(a_hi + atom * 4096, &tm_b_hi, &full_bar[s], m0 + atom * 32, (kb_begin + it) * BK);
                if (PREC == 2)
                  tma_load_2d(a_hi + Cfg::A_BYTES + atom * 4096, &tm_b_lo, &full_bar[s], m0 + atom * 32,
                              (kb_begin + it) * BK);
              }
            }
            __syncwarp();
          }
          store_wgrad_ab(s, vdummy, cur, false);
          fence_proxy_async();
          mbar_arrive(&full_bar[s]);
        }
      };
      if (group < nkb) load_wgrad_ab(group, vdummy, vb0, false);
      for (int u = 0; u < npairs; u += 2) {
        body(u, vb0, vb1);
        if (u + 1 < npairs) body(u + 1, vb1, vb0);
      }
    } else if (BTMA) {
      // B (pre-split weights) arrives by TMA; A is register double-buffered: the global loads of k-block it+2 are
      // in flight while k-block it is promoted / split / stored.
      float4 va0[8], va1[8];
      auto body = [&](int u, float4 (&cur)[8], float4 (&nxt)[8]) {
        const int it = 2 * u + group;
        if (it + 2 < nkb) {
          a_advance();
          load_row_a(nxt);
        }
        if (PROMOTE && u >= 2) promote(u - 2);
        if (it < nkb) {
          const int s = it % Cfg::STAGES;
          mbar_wait(&empty_bar[s], (((it / Cfg::STAGES) & 1) ^ 1));
          if (gtid < 32) {
            if (elect_one_sync()) {
              const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES * PREC;
              mbar_arrive_expect_tx(&full_bar[s], Cfg::B_BYTES * PREC);
              tma_load_2d(b_hi, &tm_b_hi, &full_bar[s], (kb_begin + it) * BK, n0);
              if (PREC == 2) tma_load_2d(b_hi + Cfg::B_BYTES, &tm_b_lo, &full_bar[s], (kb_begin + it) * BK, n0);
            }
            __syncwarp();
          }
          store_row_a(s, cur);
          fence_proxy_async();
          mbar_arrive(&full_bar[s]);
        }
      };
      if (group < nkb) {
        a_seek(group);
        load_row_a(va0);
      }
      for (int u = 0; u < npairs; u += 2) {
        body(u, va0, va1);
        if (u + 1 < npairs) body(u + 1, va1, va0);
      }
    } else {
      for (int u = 0; u < npairs; ++u) {
        const int it = 2 * u + group;
        const bool active = it < nkb;
        const int s = it % Cfg::STAGES;
        float4 va[8];
        float4 vb[8];
        if (active) {
          mbar_wait(&empty_bar[s], (((it / Cfg::STAGES) & 1) ^ 1));
          if (MODE == MODE_ROW) {
            if (u == 0) a_seek(it); else a_advance();
            load_row_a(va);
            load_row_b(it, vb);
          } else {
            load_wgrad(it, va, vb);
          }
        }
        if (PROMOTE && u >= 2) promote(u - 2);  // overlaps with the global loads issued above
        if (active) {
          if (MODE == MODE_ROW) {
            store_row_a(s, va);
            store_row_b(s, vb);
          } else {
            store_wgrad(s, va, vb);
          }
          fence_proxy_async();
          mbar_arrive(&full_bar[s]);
        }
      }
    }
    if (PROMOTE) {
      for (int u = (npairs > 2 ? npairs - 2 : 0); u < npairs; ++u) promote(u);
    } else {
      promote(0);
    }

    igemm_epilogue<HALF>(p, acc, m0, n0, m_tile, group, q, lane, smem_base + static_cast<uint32_t>(warp * 4608), split);
  } else {
    // ===================================================== MMA issuer (warp 8: converged loop, one elected lane issues)
    {
      constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, MODE == MODE_WGRAD, MODE == MODE_WGRAD);
      // K-major: 128B swizzle (type 2), LBO field 1, SBO = 1024 (8 rows x 128 B), k-step = +32 B inside the row.
      // MN-major tf32: 128B swizzle with 32B base (type 1), LBO = 4096 (next 32-wide MN atom), SBO = 512 (next 4 k-rows),
      // k-step = 8 k-rows = +1024 B.
      constexpr uint32_t LAYOUT = (MODE == MODE_WGRAD) ? 1u : 2u;
      constexpr uint32_t LBO = (MODE == MODE_WGRAD) ? 4096u : 16u;
      constexpr uint32_t SBO = (MODE == MODE_WGRAD) ? 512u : 1024u;
      constexpr uint32_t KSTEP = (MODE == MODE_WGRAD) ? 1024u : 32u;
      // descriptors of stage 0; stage s / k-step kk only add to the 14-bit start-address field (bytes >> 4)
      const uint64_t d_a_hi0 = umma_desc(smem_base, LBO, SBO, LAYOUT);
      const uint64_t d_a_lo0 = umma_desc(smem_base + Cfg::A_BYTES, LBO, SBO, LAYOUT);
      const uint64_t d_b_hi0 = umma_desc(smem_base + Cfg::A_BYTES * PREC, LBO, SBO, LAYOUT);
      const uint64_t d_b_lo0 = umma_desc(smem_base + Cfg::A_BYTES * PREC + Cfg::B_BYTES, LBO, SBO, LAYOUT);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % Cfg::STAGES;
        const int u = PROMOTE ? (it >> 1) : 0;
        const int b = u & (NBUF - 1);
        const bool unit_first = PROMOTE ? ((it & 1) == 0) : (it == 0);
        const bool unit_last = PROMOTE ? ((it & 1) == 1 || it == nkb - 1) : (it == nkb - 1);
        if (PROMOTE && unit_first) {
          mbar_wait(&acce_bar[b], (((u / NBUF) & 1) ^ 1));
          tc_fence_after();
        }
        mbar_wait(&full_bar[s], (it / Cfg::STAGES) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t tacc = tmem_base + static_cast<uint32_t>(b * BN);
          const uint64_t soff = static_cast<uint64_t>((s * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint64_t off = soff + ((kk * KSTEP) >> 4);
            mma_tf32_ss(tacc, d_a_hi0 + off, d_b_hi0 + off, idesc, !(unit_first && kk == 0));
            if (PREC == 2) {
              mma_tf32_ss(tacc, d_a_lo0 + off, d_b_hi0 + off, idesc, 1);
              mma_tf32_ss(tacc, d_a_hi0 + off, d_b_lo0 + off, idesc, 1);
            }
          }
          tc_commit(&empty_bar[s]);
          if (unit_last) tc_commit(&accf_bar[b]);
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace cavp
