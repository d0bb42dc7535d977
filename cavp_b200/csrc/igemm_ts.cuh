// Forward / dgrad implicit-GEMM tile kernel with the A operand in TENSOR MEMORY.
//
// Same math and epilogue as igemm_kernel<.., MODE_ROW, BTMA=true> (igemm.cuh), different operand plumbing:
//   * B (pre-split weights, hi | lo) : TMA (cp.async.bulk.tensor, 128B swizzle) into a shared-memory ring;
//   * A (gathered activations)       : one producer thread per tile row loads the 32 K-elements of its row (128 B
//                                      contiguous in NHWC), splits them into TF32 hi / lo and writes them with
//                                      tcgen05.st straight into TMEM (lane = row, column = k);
//   * MMA                            : tcgen05.mma.kind::tf32 with A from TMEM ([a_tmem]) and B from shared memory.
// With PREC = 2 every k-step issues three MMAs; in the SS form each of them re-reads A and B from shared memory and the
// kernel sits on the shared-memory bandwidth (160 KB per k-block of a 128x128 tile).  Here shared memory only carries B
// (80 KB per k-block), the TMEM reads are free, and the producers issue 4 tcgen05.st instead of 16 swizzled STS.
//
// TMEM map (512 columns): [0, 2*BN) two accumulator buffers (promotion ring), [256, 256 + 4*64) four A stages
// (32 hi + 32 lo columns each).
#pragma once
#include "igemm.cuh"

namespace cavp {

template <int BN, int PREC>
struct TsCfg {
  static constexpr bool PROMOTE = (PREC == 2);
  static constexpr int NBUF = PROMOTE ? 2 : 1;
  static constexpr int STAGES = 4;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = B_BYTES * PREC;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 8 * 4608 /*epilogue scratch*/;
  static constexpr int TMEM_COLS = 512;
  static constexpr int A_COL0 = 256;
  static constexpr int A_STAGE_COLS = 64;
  static constexpr int HALF = BN / 2;
};

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive columns <- 16 registers per thread
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int BN, int PREC>
__global__ void __launch_bounds__(CTA_THREADS, 1)
igemm_ts_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_b_hi,
                const __grid_constant__ CUtensorMap tm_b_lo) {
  using Cfg = TsCfg<BN, PREC>;
  constexpr bool PROMOTE = Cfg::PROMOTE;
  constexpr int NBUF = Cfg::NBUF;
  constexpr int HALF = Cfg::HALF;
  static_assert(BN == 64 || BN == 128, "BN must be 64 or 128");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                           // [STAGES] A rows stored (128 arrivals) + B bytes landed (+1)
  uint64_t* empty_bar = bars + Cfg::STAGES;            // [STAGES] MMA -> producers
  uint64_t* accf_bar = bars + 2 * Cfg::STAGES;         // [NBUF]
  uint64_t* acce_bar = bars + 2 * Cfg::STAGES + NBUF;  // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 2 * NBUF);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int tile = blockIdx.x;
  const int n_tile = tile % p.n_tiles;
  const int m_tile = tile / p.n_tiles;
  const int m0 = m_tile * BM;
  const int n0 = n_tile * BN;
  const int split = blockIdx.y;
  const int kb_begin = static_cast<int>((static_cast<long long>(p.num_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.num_kb) * (split + 1)) / p.splits);
  const int nkb = kb_end - kb_begin;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], GROUP_THREADS + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], PRODUCER_THREADS);
    }
    fence_mbar_init();
  }
  if (tid == 32) {
    tma_prefetch_desc(&tm_b_hi);
    if (PREC == 2) tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == 8) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    const int group = warp >> 2;
    const int q = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
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
        tmem_ld16(tmem_base + lane_base + static_cast<uint32_t>(b * BN + group * HALF + cgrp * 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[cgrp * 16 + j] += v[j];
      }
      if (PROMOTE) {
        tc_fence_before();
        mbar_arrive(&acce_bar[b]);
      }
    };

    // ---- this thread's tile row (one GEMM row = one output pixel for fwd, one input pixel for dgrad)
    const int m = m0 + q * 32 + lane;
    int r_base = -1, r_y = 0, r_x = 0;
    if (m < p.M) {
      uint32_t n, rem, oy, ox;
      p.div_howo.divmod(static_cast<uint32_t>(m), n, rem);
      p.div_wo.divmod(rem, oy, ox);
      r_base = static_cast<int>(n) * p.Hs * p.Ws;
      if (p.dgrad) {
        r_y = static_cast<int>(oy) + p.pad;
        r_x = static_cast<int>(ox) + p.pad;
      } else {
        r_y = static_cast<int>(oy) * p.stride - p.pad;
        r_x = static_cast<int>(ox) * p.stride - p.pad;
      }
    }
    const int taps = p.R * p.S;
    // element offset of this row's source pixel for filter tap `tap`, or -1 (padding / stride hole / no such tap)
    auto tap_offset = [&](int tap) -> int {
      if (r_base < 0 || tap >= taps) return -1;
      uint32_t ky, kx;
      p.div_s.divmod(static_cast<uint32_t>(tap), ky, kx);
      int iy, ix;
      bool ok = true;
      if (p.dgrad) {
        iy = r_y - static_cast<int>(ky) * p.dil;
        ix = r_x - static_cast<int>(kx) * p.dil;
        if (p.stride > 1) {
          ok = iy >= 0 && ix >= 0 && (iy % p.stride) == 0 && (ix % p.stride) == 0;
          iy /= p.stride;
          ix /= p.stride;
        }
      } else {
        iy = r_y + static_cast<int>(ky) * p.dil;
        ix = r_x + static_cast<int>(kx) * p.dil;
      }
      ok = ok && static_cast<unsigned>(iy) < static_cast<unsigned>(p.Hs) &&
           static_cast<unsigned>(ix) < static_cast<unsigned>(p.Ws);
      return ok ? (r_base + iy * p.Ws + ix) * p.ldx : -1;
    };
    // gather state for the k-block being loaded: k0 = first k, (tap0, ci0) its decomposition, offsets of tap0 / tap0+1
    int a_k0 = 0, a_tap0 = 0, a_ci0 = 0, a_off0 = -1, a_off1 = -1;
    const bool wide_c = p.C >= BK;  // a k-block then spans at most two taps
    auto a_seek = [&](int it) {
      a_k0 = (kb_begin + it) * BK;
      uint32_t tap, ci;
      p.div_c.divmod(static_cast<uint32_t>(a_k0 < p.K ? a_k0 : 0), tap, ci);
      a_tap0 = static_cast<int>(tap);
      a_ci0 = static_cast<int>(ci);
      a_off0 = tap_offset(a_tap0);
      a_off1 = tap_offset(a_tap0 + 1);
    };
    auto a_advance = [&]() {
      a_k0 += 2 * BK;
      a_ci0 += 2 * BK;
      if (a_ci0 >= p.C) {
        if (wide_c) {
          while (a_ci0 >= p.C) {
            a_ci0 -= p.C;
            ++a_tap0;
          }
        } else {
          uint32_t tap, ci;
          p.div_c.divmod(static_cast<uint32_t>(a_k0 < p.K ? a_k0 : 0), tap, ci);
          a_tap0 = static_cast<int>(tap);
          a_ci0 = static_cast<int>(ci);
        }
        a_off0 = tap_offset(a_tap0);
        a_off1 = tap_offset(a_tap0 + 1);
      }
    };
    auto load_row = [&](float4 (&va)[8]) {
      if (wide_c) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int cij = a_ci0 + 4 * j;
          const bool second = cij >= p.C;
          const int off = second ? a_off1 : a_off0;
          const int ci = second ? cij - p.C : cij;
          va[j] = (off >= 0 && a_k0 + 4 * j < p.K) ? ldg_nc_v4(p.x + off + ci) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {  // tiny channel counts (stems): decode every 16-byte chunk
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = a_k0 + 4 * j;
          int off = -1;
          uint32_t tap = 0, ci = 0;
          if (k < p.K) {
            p.div_c.divmod(static_cast<uint32_t>(k), tap, ci);
            off = tap_offset(static_cast<int>(tap));
          }
          va[j] = off >= 0 ? ldg_nc_v4(p.x + off + static_cast<int>(ci)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    // split into TF32 hi / lo and write this row's 32 k-columns into A stage s of tensor memory
    auto store_row = [&](int s, const float4 (&va)[8]) {
      const uint32_t a_hi = tmem_base + lane_base + static_cast<uint32_t>(Cfg::A_COL0 + s * Cfg::A_STAGE_COLS);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = va[h * 4 + j];
          hi[4 * j + 0] = tf32_rn(v.x); hi[4 * j + 1] = tf32_rn(v.y);
          hi[4 * j + 2] = tf32_rn(v.z); hi[4 * j + 3] = tf32_rn(v.w);
          if (PREC == 2) {
            lo[4 * j + 0] = tf32_rn(v.x - hi[4 * j + 0]); lo[4 * j + 1] = tf32_rn(v.y - hi[4 * j + 1]);
            lo[4 * j + 2] = tf32_rn(v.z - hi[4 * j + 2]); lo[4 * j + 3] = tf32_rn(v.w - hi[4 * j + 3]);
          }
        }
        tmem_st16(a_hi + h * 16, hi);
        if (PREC == 2) tmem_st16(a_hi + 32 + h * 16, lo);
      }
      tmem_wait_st();
    };

    const int npairs = (nkb + 1) >> 1;
    float4 va0[8], va1[8];
    auto body = [&](int u, float4 (&cur)[8], float4 (&nxt)[8]) {
      const int it = 2 * u + group;
      if (it + 2 < nkb) {
        a_advance();
        load_row(nxt);
      }
      if (PROMOTE && u >= 2) promote(u - 2);
      if (it < nkb) {
        const int s = it % Cfg::STAGES;
        mbar_wait(&empty_bar[s], (((it / Cfg::STAGES) & 1) ^ 1));
        tc_fence_after();
        if ((tid & (GROUP_THREADS - 1)) == 0) {
          const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], Cfg::B_BYTES * PREC);
          tma_load_2d(b_hi, &tm_b_hi, &full_bar[s], (kb_begin + it) * BK, n0);
          if (PREC == 2) tma_load_2d(b_hi + Cfg::B_BYTES, &tm_b_lo, &full_bar[s], (kb_begin + it) * BK, n0);
        }
        store_row(s, cur);
        tc_fence_before();
        mbar_arrive(&full_bar[s]);
      }
    };
    if (group < nkb) {
      a_seek(group);
      load_row(va0);
    }
    for (int u = 0; u < npairs; u += 2) {
      body(u, va0, va1);
      if (u + 1 < npairs) body(u + 1, va1, va0);
    }
    if (PROMOTE) {
      for (int u = (npairs > 2 ? npairs - 2 : 0); u < npairs; ++u) promote(u);
    } else {
      promote(0);
    }
    igemm_epilogue<HALF>(p, acc, m0, n0, m_tile, group, q, lane, smem_base + static_cast<uint32_t>(Cfg::STAGES * Cfg::STAGE_BYTES + 256 + warp * 4608));
  } else {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, 0, 0);
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
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(b * BN);
        const uint32_t ta_hi = tmem_base + static_cast<uint32_t>(Cfg::A_COL0 + s * Cfg::A_STAGE_COLS);
        const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES, b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint64_t db_hi = umma_desc(b_hi + kk * 32, 16, 1024, 2);
          mma_tf32_ts(tacc, ta_hi + kk * UMMA_K, db_hi, idesc, !(unit_first && kk == 0));
          if (PREC == 2) {
            const uint64_t db_lo = umma_desc(b_lo + kk * 32, 16, 1024, 2);
            mma_tf32_ts(tacc, ta_hi + 32 + kk * UMMA_K, db_hi, idesc, 1);
            mma_tf32_ts(tacc, ta_hi + kk * UMMA_K, db_lo, idesc, 1);
          }
        }
        tc_commit(&empty_bar[s]);
        if (unit_last) tc_commit(&accf_bar[b]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace cavp
