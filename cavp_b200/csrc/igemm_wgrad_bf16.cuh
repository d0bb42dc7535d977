// bf16 CTA-pair weight-gradient kernel of the bf16 row (cavp_prec = 3, BASELINE.json configs[2]):
//   dW[Cout x K] (+)= dY[P x Cout]^T * im2col(X)[P x K],   tcgen05.mma.cta_group::2.kind::f16, fp32 accumulation in TMEM.
//
// Structure of igemm_wgrad2.cuh (M = 256 output channels per pair, N = 128 weight columns: each CTA gathers 64 of them)
// with bf16 MN-major operands.  Both operands are stored the way they lie in HBM - a pixel (the reduction index) per
// 128-byte shared-memory row, 64 channels along the row - which is the canonical MN-major SWIZZLE_128B layout of a 16-bit
// UMMA operand ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 8 pixel rows = 1024 B (SBO), the next 64-channel atom
// one 64-row block further (LBO = 8192 B).  A k-block is 64 pixels = four K = 16 MMAs.
//   * dY arrives pre-converted to bf16 (dense [P][Cout], written by the kernel that produced dY or by one conversion
//     pass) through TMA: two 64-channel x 64-pixel boxes per CTA and k-block, 128-byte swizzle;
//   * im2col(X) is gathered by the producer warps (two groups alternating k-blocks): 8 consecutive channels of one filter
//     tap = 2 x 16-byte fp32 loads -> cvt.rn.bf16x2 -> one 16-byte shared-memory store with the same swizzle;
//   * no promotion: the whole pixel range of a split accumulates in TMEM (the splits are added with red.global.add,
//     as in the TF32 kernels), then the eight warps run the common epilogue.
// Envelope: C % 8 == 0 (a chunk never straddles a tap), Cout % 8 == 0 (TMA stride), Cout > 128; everything else stays on
// the TF32 weight-gradient kernels.
#pragma once
#include "igemm_bf16.cuh"

namespace cavp {

struct WgBf16Cfg {
  static constexpr int BN = 128;
  static constexpr int BH = 64;             // im2col columns gathered by one CTA (one 64-element atom)
  static constexpr int KPIX = 64;           // pixels per k-block
  static constexpr int A_BYTES = 2 * KPIX * 128;  // two 64-channel atoms x 64 pixel rows x 128 B
  static constexpr int B_BYTES = KPIX * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = BN;
  static constexpr int HALF = BN / 2;
  static_assert(8 * 4608 <= RING_BYTES, "epilogue scratch fits in the ring");
};

// instruction descriptor: kind::f16, bf16 A/B, fp32 D, both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
  return umma_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CTA_THREADS, 1)
igemm_wgrad_bf16_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_a) {
  using Cfg = WgBf16Cfg;
  constexpr int HALF = Cfg::HALF;
  constexpr int BN = Cfg::BN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + Cfg::RING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* accf_bar = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();

  const int tile = blockIdx.x >> 1;  // pair index: (m_pair, n_tile), n fastest
  const int n_tile = tile % p.n_tiles;
  const int m_pair = tile / p.n_tiles;
  const int m_tile = m_pair * 2 + static_cast<int>(rank);
  const int m0 = m_tile * BM;                             // this CTA's output channels
  const int n0 = n_tile * BN;                             // the pair's weight columns
  const int nb0 = n0 + static_cast<int>(rank) * Cfg::BH;  // the half this CTA gathers
  const int split = blockIdx.y;
  const int kb_begin = static_cast<int>((static_cast<long long>(p.num_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.num_kb) * (split + 1)) / p.splits);
  const int nkb = kb_end - kb_begin;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 2 * (GROUP_THREADS / 32) + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accf_bar[0], 1);
    fence_mbar_init();
  }
  if (tid == 32) tma_prefetch_desc(&tm_a);
  if (warp == 8) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ===================================================== producers (+ epilogue)
    const int group = warp >> 2;
    const int gtid = tid & (GROUP_THREADS - 1);
    const int q = warp & 3;

    // this thread's part of the 64-column x 64-pixel im2col half tile: 16-byte bf16 chunk cc (8 columns), pixels rr*4 .. +3
    const int cc = gtid & 7;
    const int rr = gtid >> 3;
    const int j = nb0 + cc * 8;
    const bool j_ok = j < p.Ncols;
    uint32_t tap, ky, kx, wg_ci;
    p.div_c.divmod(static_cast<uint32_t>(j_ok ? j : 0), tap, wg_ci);
    p.div_s.divmod(tap, ky, kx);
    const int wg_dy = static_cast<int>(ky) * p.dil - p.pad;
    const int wg_dx = static_cast<int>(kx) * p.dil - p.pad;

    auto load_b = [&](int it, float4 (&vb)[8]) {
      const int pixb = (kb_begin + it) * Cfg::KPIX + rr * 4;
      uint32_t n, rem, oy, ox;
      p.div_howo.divmod(static_cast<uint32_t>(pixb < p.red_len ? pixb : 0), n, rem);
      p.div_wo.divmod(rem, oy, ox);
      int base = static_cast<int>(n) * p.Hs * p.Ws;
      int y = static_cast<int>(oy) * p.stride + wg_dy, x = static_cast<int>(ox) * p.stride + wg_dx;
      const int x_wrap = p.Wo * p.stride + wg_dx, y_wrap = p.Ho * p.stride + wg_dy;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = (pixb + i < p.red_len) && j_ok && static_cast<unsigned>(y) < static_cast<unsigned>(p.Hs) &&
                        static_cast<unsigned>(x) < static_cast<unsigned>(p.Ws);
        if (ok) {
          const float* src = p.x + static_cast<size_t>(base + y * p.Ws + x) * p.ldx + wg_ci;
          vb[2 * i] = ldg_nc_v4(src);
          vb[2 * i + 1] = ldg_nc_v4(src + 4);
        } else {
          vb[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
          vb[2 * i + 1] = vb[2 * i];
        }
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
    auto store_b = [&](int s, const float4 (&vb)[8]) {
      const uint32_t b_st = smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rr * 4 + i;
        const uint32_t off = static_cast<uint32_t>(r * 128) + static_cast<uint32_t>((cc ^ (r & 7)) << 4);
        const float4 lo = vb[2 * i], hi = vb[2 * i + 1];
        st_shared_v4_b32(b_st + off, cvt_bf16x2(lo.x, lo.y), cvt_bf16x2(lo.z, lo.w), cvt_bf16x2(hi.x, hi.y),
                         cvt_bf16x2(hi.z, hi.w));
      }
    };

    const int npairs = (nkb + 1) >> 1;
    float4 vb0[8], vb1[8];
    auto body = [&](int u, float4 (&cur)[8], float4 (&nxt)[8]) {
      const int it = 2 * u + group;
      if (it + 2 < nkb) load_b(it + 2, nxt);
      if (it < nkb) {
        const int s = it % Cfg::STAGES;
        mbar_wait(&empty_bar[s], (((it / Cfg::STAGES) & 1) ^ 1));
        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[s]), 0);
        if (gtid < 32) {
          if (elect_one_sync()) {
            const uint32_t a_st = smem_base + s * Cfg::STAGE_BYTES;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::A_BYTES);
            tma_load_2d_pair(a_st, &tm_a, full_leader, m0, (kb_begin + it) * Cfg::KPIX);
            tma_load_2d_pair(a_st + Cfg::KPIX * 128, &tm_a, full_leader, m0 + 64, (kb_begin + it) * Cfg::KPIX);
          }
          __syncwarp();
        }
        store_b(s, cur);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(full_leader);
      }
    };
    if (group < nkb) load_b(group, vb0);
    for (int u = 0; u < npairs; u += 2) {
      body(u, vb0, vb1);
      if (u + 1 < npairs) body(u + 1, vb1, vb0);
    }
    // all MMAs of this split are done: accumulators -> registers -> common epilogue (red.add when splits > 1)
    mbar_wait(&accf_bar[0], 0);
    tc_fence_after();
    float acc[HALF];
#pragma unroll
    for (int cgrp = 0; cgrp < HALF / 32; ++cgrp) {
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(group * HALF + cgrp * 32), v);
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) acc[cgrp * 32 + jj] = v[jj];
    }
    igemm_epilogue<HALF>(p, acc, m0, n0, m_tile, group, q, lane, smem_base + static_cast<uint32_t>(warp * 4608));
  } else {
    // ===================================================== MMA issuer (leader CTA, converged warp + elect)
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_mn(2 * BM, BN);
      const uint64_t d_a0 = umma_desc(smem_base, Cfg::KPIX * 128, 1024, 2);                 // LBO = next 64-channel atom
      const uint64_t d_b0 = umma_desc(smem_base + Cfg::A_BYTES, Cfg::KPIX * 128, 1024, 2);  // (one atom: LBO unused)
      for (int it = 0; it < nkb; ++it) {
        const int s = it % Cfg::STAGES;
        mbar_wait(&full_bar[s], (it / Cfg::STAGES) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t soff = static_cast<uint64_t>((s * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < Cfg::KPIX / UMMA_K16; ++kk) {
            const uint64_t off = soff + ((kk * 2048) >> 4);  // 16 pixel rows per k-step
            mma_bf16_ss_pair(tmem_base, d_a0 + off, d_b0 + off, idesc, !(it == 0 && kk == 0));
          }
          tc_commit_pair(&empty_bar[s], 3);
          if (it == nkb - 1) tc_commit_pair(&accf_bar[0], 3);
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 8) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace cavp
