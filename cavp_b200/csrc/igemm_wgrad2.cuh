// CTA-pair weight-gradient kernel: dW[Cout x K] (+)= dY[P x Cout]^T * im2col(X)[P x K] with ONE
// tcgen05.mma.cta_group::2 per k-step (M = 256 output channels: 128 per CTA; N = 128 weight columns: 64 per CTA).
//
// Same operand layouts and precision scheme as igemm_kernel<128, PREC, MODE_WGRAD, BTMA=true> (igemm.cuh): dY is
// pre-split (hi | lo, dense) and arrives by TMA in the MN-major 128B/32B-atom swizzle, im2col(X) is gathered by the
// producer warps, products are 3xTF32 with promotion every 64 pixels.  What the pair changes: each CTA gathers, splits
// and stores only HALF of the im2col tile (64 of the 128 columns) - the producer work that bounds the single-CTA
// kernel (54 % tensor-pipe active) - and reads 72 KB instead of 96 KB of operands per k-block from shared memory.
// Cross-CTA protocol as in igemm_ws2.cuh: full[s] in the leader (one arrival per producing warp of both CTAs + the
// leader's expect_tx for both dY tiles), empty[s] / accf[b] armed in both CTAs by multicast commits, acce[b] in the
// leader (one arrival per promotion warp of both CTAs).
#pragma once
#include "igemm.cuh"

namespace cavp {

template <int PREC>
struct Wg2Cfg {
  static constexpr int BN = 128;
  static constexpr int BH = 64;  // im2col columns gathered by one CTA
  static constexpr bool PROMOTE = (PREC == 2);
  static constexpr int NBUF = PROMOTE ? 4 : 1;
  static constexpr int A_BYTES = BM * 128;  // 128 channels x 32 pixels x 4 B
  static constexpr int B_BYTES = BH * 128;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * PREC;
  static constexpr int STAGES = 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  // the epilogue scratch (8 warps x 4608 B) reuses the stage ring once all MMAs are done, as in igemm_kernel
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = NBUF * BN < 32 ? 32 : NBUF * BN;
  static constexpr int HALF = BN / 2;
  static_assert(8 * 4608 <= RING_BYTES, "epilogue scratch fits in the ring");
};

template <int PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CTA_THREADS, 1)
igemm_wgrad2_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_a_hi,
                    const __grid_constant__ CUtensorMap tm_a_lo) {
  using Cfg = Wg2Cfg<PREC>;
  constexpr bool PROMOTE = Cfg::PROMOTE;
  constexpr int NBUF = Cfg::NBUF;
  constexpr int HALF = Cfg::HALF;
  constexpr int BN = Cfg::BN;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + Cfg::RING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* accf_bar = bars + 2 * Cfg::STAGES;
  uint64_t* acce_bar = bars + 2 * Cfg::STAGES + NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 2 * NBUF);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();

  const int tile = blockIdx.x >> 1;  // pair index: (m_pair, n_tile), n fastest
  const int n_tile = tile % p.n_tiles;
  const int m_pair = tile / p.n_tiles;
  const int m_tile = m_pair * 2 + static_cast<int>(rank);
  const int m0 = m_tile * BM;                               // this CTA's output channels
  const int n0 = n_tile * BN;                               // the pair's weight columns
  const int nb0 = n0 + static_cast<int>(rank) * Cfg::BH;    // the half this CTA gathers
  const int split = blockIdx.y;
  const int kb_begin = static_cast<int>((static_cast<long long>(p.num_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.num_kb) * (split + 1)) / p.splits);
  const int nkb = kb_end - kb_begin;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 2 * (GROUP_THREADS / 32) + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], 2 * (PRODUCER_THREADS / 32));
    }
    fence_mbar_init();
  }
  if (tid == 32) {
    tma_prefetch_desc(&tm_a_hi);
    if (PREC == 2) tma_prefetch_desc(&tm_a_lo);
  }
  if (warp == 8) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ===================================================== producers (+ promotion + epilogue)
    const int group = warp >> 2;
    const int gtid = tid & (GROUP_THREADS - 1);
    const int q = warp & 3;
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
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acce_bar[b]), 0));
      }
    };

    // this thread's part of the 64-column x 32-pixel im2col half tile: 16-byte chunk cc (4 columns), pixels rr*4 .. +3
    const int cc = gtid & 15;
    const int rr = gtid >> 4;
    const int j = nb0 + cc * 4;
    const bool j_ok = j < p.Ncols;
    uint32_t tap, ky, kx, wg_ci;
    p.div_c.divmod(static_cast<uint32_t>(j_ok ? j : 0), tap, wg_ci);
    p.div_s.divmod(tap, ky, kx);
    const int wg_dy = static_cast<int>(ky) * p.dil - p.pad;
    const int wg_dx = static_cast<int>(kx) * p.dil - p.pad;

    auto load_b = [&](int it, float4 (&vb)[4]) {
      const int pixb = (kb_begin + it) * BK + rr * 4;
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
    auto store_b = [&](int s, const float4 (&vb)[4]) {
      const uint32_t b_hi = smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES * PREC, b_lo = b_hi + Cfg::B_BYTES;
      const uint32_t atom_off = static_cast<uint32_t>((cc >> 3) * 4096);
      const int c16 = cc & 7;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rr * 4 + i;
        const uint32_t off = atom_off + static_cast<uint32_t>(r * 128) +
                             static_cast<uint32_t>((((c16 >> 1) ^ (r & 3)) << 5) | ((c16 & 1) << 4));
        store_split_fast<PREC>(b_hi + off, b_lo + off, vb[i]);
      }
    };

    const int npairs = (nkb + 1) >> 1;
    float4 vb0[4], vb1[4], vb2[4];
    // three register buffers of 4 float4: the gather of k-blocks it+2 and it+4 is in flight while it is stored
    auto body = [&](int u, float4 (&cur)[4], float4 (&nxt2)[4]) {
      const int it = 2 * u + group;
      if (it + 4 < nkb) load_b(it + 4, nxt2);
      if (PROMOTE && u >= 2) promote(u - 2);
      if (it < nkb) {
        const int s = it % Cfg::STAGES;
        mbar_wait(&empty_bar[s], (((it / Cfg::STAGES) & 1) ^ 1));
        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[s]), 0);
        if (gtid < 32) {
          if (elect_one_sync()) {
            const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::A_BYTES * PREC);
#pragma unroll
            for (int atom = 0; atom < BM / 32; ++atom) {
              tma_load_2d_pair(a_hi + atom * 4096, &tm_a_hi, full_leader, m0 + atom * 32, (kb_begin + it) * BK);
              if (PREC == 2)
                tma_load_2d_pair(a_hi + Cfg::A_BYTES + atom * 4096, &tm_a_lo, full_leader, m0 + atom * 32,
                                 (kb_begin + it) * BK);
            }
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
    if (group + 2 < nkb) load_b(group + 2, vb1);
    for (int u = 0; u < npairs; u += 3) {
      body(u, vb0, vb2);
      if (u + 1 < npairs) body(u + 1, vb1, vb0);
      if (u + 2 < npairs) body(u + 2, vb2, vb1);
    }
    if (PROMOTE) {
      for (int u = (npairs > 2 ? npairs - 2 : 0); u < npairs; ++u) promote(u);
    } else {
      promote(0);
    }
    igemm_epilogue<HALF>(p, acc, m0, n0, m_tile, group, q, lane, smem_base + static_cast<uint32_t>(warp * 4608));
  } else {
    // ===================================================== MMA issuer (leader CTA, converged warp + elect)
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(2 * BM, BN, 1, 1);
      const uint64_t d_a_hi0 = umma_desc(smem_base, 4096, 512, 1);
      const uint64_t d_a_lo0 = umma_desc(smem_base + Cfg::A_BYTES, 4096, 512, 1);
      const uint64_t d_b_hi0 = umma_desc(smem_base + Cfg::A_BYTES * PREC, 4096, 512, 1);
      const uint64_t d_b_lo0 = umma_desc(smem_base + Cfg::A_BYTES * PREC + Cfg::B_BYTES, 4096, 512, 1);
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
            const uint64_t off = soff + ((kk * 1024) >> 4);  // 8 pixel rows per k-step
            mma_tf32_ss_pair(tacc, d_a_hi0 + off, d_b_hi0 + off, idesc, !(unit_first && kk == 0));
            if (PREC == 2) {
              mma_tf32_ss_pair(tacc, d_a_lo0 + off, d_b_hi0 + off, idesc, 1);
              mma_tf32_ss_pair(tacc, d_a_hi0 + off, d_b_lo0 + off, idesc, 1);
            }
          }
          tc_commit_pair(&empty_bar[s], 3);
          if (unit_last) tc_commit_pair(&accf_bar[b], 3);
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
