// CTA-pair version of the persistent warp-specialised forward / dgrad implicit-GEMM kernel (igemm_ws.cuh).
//
// Why: at PREC=2 a 128 x 128 tile moves 160 KB through shared memory per k-block (64 KB of hi/lo operand writes + 96 KB
// of tcgen05.mma operand reads) against 768 clk of MMA time - the kernel sits on the 128 B/clk shared-memory port, not
// on the tensor pipe (measured 49 % tensor-pipe active).  A pair of CTAs on the two SMs of a TPC issuing ONE
// tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA, N = 128: 64 weight rows per CTA) halves the B-operand traffic
// per SM: 32 KB A writes + 16 KB B writes + 72 KB operand reads = 120 KB per k-block and SM.
//
// Tiles are 128 or 160 columns wide (64 / 80 weight rows per CTA; 160 where it pads N less, e.g. N = 304, or barely more on
// long-K GEMMs): the wider tile amortises the activation tile over more columns - measured 80 % vs 66 % tensor-pipe active.
// Linear layers (1x1, stride 1) take a producer fast path whose register double buffer runs across tile boundaries.
//
// Roles per CTA (same thread layout as igemm_ws): warps 0-3 promotion + epilogue of the CTA's own 128 rows, warps 4-11
// activation producers (two groups alternating k-blocks), warp 12 = MMA issuer (leader CTA only; idle in the peer).
// Cross-CTA protocol (all barriers at the same shared-memory offsets in both CTAs):
//   full[s]   leader only.  9 arrivals per phase: lane 0 of each of the 4 warps of the producing group of EACH CTA (after
//             fence.proxy.async + __syncwarp) + the leader's arrive.expect_tx covering both weight halves; the peer's
//             TMA completes its bytes on the leader's barrier (cp.async.bulk.tensor...cta_group::2).
//   empty[s]  both CTAs, armed by ONE multicast tcgen05.commit of the leader.
//   accf[b]   both CTAs, multicast commit: the 64-wide K unit in TMEM buffer b is complete (each CTA holds its rows).
//   acce[b]   leader only.  8 arrivals: one per promotion warp of both CTAs once the unit has been read out of TMEM.
#pragma once
#include "igemm_ws.cuh"

namespace cavp {

template <int BN, int PREC>
struct Ws2Cfg {
  static constexpr bool PROMOTE = (PREC == 2);
  static constexpr int NBUF = 512 / BN >= 4 ? 4 : 512 / BN;  // BN = 160: three 160-column accumulators
  static constexpr int BH = BN / 2;  // weight rows held by one CTA
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BH * 128;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * PREC;
  static constexpr int STAGES = (PREC == 2) ? (BN > 128 ? 3 : 4) : 6;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int ROWTAB_BYTES = BM * 8;
  static constexpr int SCRATCH_BYTES = WS_EPI_WARPS * 4608;
  static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + ROWTAB_BYTES + SCRATCH_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(2 * STAGES + 2 * NBUF + 1 <= BAR_BYTES / 8, "barrier area");
};

struct Ws2Work {
  int m_pair, n_tile, kb_begin, nkb, split;
};
__device__ __forceinline__ Ws2Work ws2_decode(const IgemmParams& p, int w, int m_pairs) {
  Ws2Work r;
  const int tiles = p.n_tiles * m_pairs;
  const int split = w / tiles;
  const int tile = w - split * tiles;
  r.n_tile = tile % p.n_tiles;
  r.m_pair = tile / p.n_tiles;
  r.split = split;
  r.kb_begin = static_cast<int>((static_cast<long long>(p.num_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.num_kb) * (split + 1)) / p.splits);
  r.nkb = kb_end - r.kb_begin;
  return r;
}

template <int BN, int PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WS_THREADS, 1)
igemm_ws2_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_b_hi,
                 const __grid_constant__ CUtensorMap tm_b_lo, int total_work, int m_pairs) {
  using Cfg = Ws2Cfg<BN, PREC>;
  constexpr bool PROMOTE = Cfg::PROMOTE;
  constexpr int NBUF = Cfg::NBUF;
  static_assert(BN == 128 || BN == 160, "pair kernel: 128- or 160-column tiles (64 / 80 weight rows per CTA)");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + Cfg::RING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* accf_bar = bars + 2 * Cfg::STAGES;
  uint64_t* acce_bar = bars + 2 * Cfg::STAGES + NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 2 * NBUF);
  int2* rowtab = reinterpret_cast<int2*>(smem_aligned + Cfg::RING_BYTES + Cfg::BAR_BYTES);
  const uint32_t scratch_base = smem_base + Cfg::RING_BYTES + Cfg::BAR_BYTES + Cfg::ROWTAB_BYTES;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 2 * (GROUP_THREADS / 32) + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], 2 * WS_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (tid == 32) {
    tma_prefetch_desc(&tm_b_hi);
    if (PREC == 2) tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == WS_MMA_WARP) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < WS_EPI_WARPS) {
    // ================================================================= promotion + epilogue (thread = tile row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t scratch = scratch_base + static_cast<uint32_t>(warp * 4608);
    int ubase = 0;
    for (int w = pair_id; w < total_work; w += num_pairs) {
      const Ws2Work wk = ws2_decode(p, w, m_pairs);
      const int m_tile = wk.m_pair * 2 + static_cast<int>(rank);
      const int nunits = PROMOTE ? ((wk.nkb + p.unit_kb - 1) / p.unit_kb) : 1;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      for (int u = 0; u < nunits; ++u) {
        const int U = ubase + u;
        const int b = U % NBUF;
        mbar_wait(&accf_bar[b], (U / NBUF) & 1);
        tc_fence_after();
        if constexpr (BN > 128) {  // 160 accumulators per thread: 16-column loads keep the temporaries small
#pragma unroll
          for (int cg = 0; cg < BN / 16; ++cg) {
            float v[16];
            tmem_ld16(tmem_base + lane_base + static_cast<uint32_t>(b * BN + cg * 16), v);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[cg * 16 + j] += v[j];
          }
        } else {
#pragma unroll
          for (int cg = 0; cg < BN / 32; ++cg) {
            float v[32];
            tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(b * BN + cg * 32), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[cg * 32 + j] += v[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acce_bar[b]), 0));
      }
      ubase += nunits;
      igemm_epilogue<BN>(p, acc, m_tile * BM, wk.n_tile * BN, m_tile, 0, q, lane, scratch, wk.split);
    }
  } else if (warp < WS_MMA_WARP) {
    // ================================================================= producers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    const int ptid = tid - WS_PROD_WARP0 * 32;  // 0..255
    const int group = ptid >> 7;
    const int gtid = ptid & (GROUP_THREADS - 1);
    const int c = gtid & 7;
    const int r0 = gtid >> 3;
    const uint32_t swz = static_cast<uint32_t>((c ^ (r0 & 7)) << 4);
    int gbase = 0;  // global k-block counter at the start of the current work item
    // ---- linear layers (1x1, stride 1, no padding, no split-K, even k-block count): the A row of output row m is row m
    // of x, so there is no row table to rebuild and the register double buffer keeps running ACROSS tiles - the gather
    // of the next tile's first k-blocks is in flight while the current tile's last ones are stored.  For K = 304 (10
    // k-blocks per tile) the per-tile restart (two named barriers, table, exposed first loads) cost ~20 % of the tile.
    const bool linear = p.R == 1 && p.S == 1 && p.stride == 1 && p.pad == 0 && p.splits == 1 && (p.num_kb & 1) == 0 &&
                        p.Hs == p.Ho && p.Ws == p.Wo;
    if (linear) {
      const int nkb = p.num_kb;
      struct Pos { int w, it, m0, nb0; };
      auto decode = [&](int w, int it) {
        Pos q_{w, it, 0, 0};
        if (w < total_work) {
          const Ws2Work wk = ws2_decode(p, w, m_pairs);
          q_.m0 = (wk.m_pair * 2 + static_cast<int>(rank)) * BM;
          q_.nb0 = wk.n_tile * BN + static_cast<int>(rank) * Cfg::BH;
        }
        return q_;
      };
      auto advance = [&](const Pos& a) {
        if (a.it + 2 < nkb) return Pos{a.w, a.it + 2, a.m0, a.nb0};
        return decode(a.w + num_pairs, group);
      };
      auto load_lin = [&](const Pos& a, float4 (&va)[8]) {
        const int k = a.it * BK + c * 4;
        const bool kvalid = k < p.K;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = a.m0 + r0 + 16 * i;
          va[i] = (kvalid && m < p.M) ? ldg_nc_v4(p.x + static_cast<size_t>(m) * p.ldx + k)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      int G = group;  // this group's k-blocks are every other one of the CTA's global k-block sequence
      auto body = [&](Pos& cur_pos, float4 (&cur)[8], float4 (&nxt)[8]) {
        const Pos nxt_pos = advance(cur_pos);
        if (nxt_pos.w < total_work) load_lin(nxt_pos, nxt);
        const int s = G % Cfg::STAGES;
        mbar_wait(&empty_bar[s], (((G / Cfg::STAGES) & 1) ^ 1));
        const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES, a_lo = a_hi + Cfg::A_BYTES;
        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[s]), 0);
        if (gtid < 32) {
          if (elect_one_sync()) {
            const uint32_t b_hi = a_hi + Cfg::A_BYTES * PREC;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::B_BYTES * PREC);
            tma_load_2d_pair(b_hi, &tm_b_hi, full_leader, cur_pos.it * BK, cur_pos.nb0);
            if (PREC == 2) tma_load_2d_pair(b_hi + Cfg::B_BYTES, &tm_b_lo, full_leader, cur_pos.it * BK, cur_pos.nb0);
          }
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t off = static_cast<uint32_t>((r0 + 16 * i) * 128) + swz;
          store_split_fast<PREC>(a_hi + off, a_lo + off, cur[i]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(full_leader);
        G += 2;
        cur_pos = nxt_pos;
      };
      float4 va0[8], va1[8];
      Pos pos = decode(pair_id, group);
      if (pos.w < total_work) load_lin(pos, va0);
      while (pos.w < total_work) {
        body(pos, va0, va1);
        if (pos.w < total_work) body(pos, va1, va0);
      }
    } else
    for (int w = pair_id; w < total_work; w += num_pairs) {
      const Ws2Work wk = ws2_decode(p, w, m_pairs);
      const int m0 = (wk.m_pair * 2 + static_cast<int>(rank)) * BM;
      const int nb0 = wk.n_tile * BN + static_cast<int>(rank) * Cfg::BH;  // first weight row of this CTA's half
      named_bar_sync(1, PRODUCER_THREADS);
      if (ptid < BM) {
        const int m = m0 + ptid;
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
        rowtab[ptid] = e;
      }
      named_bar_sync(1, PRODUCER_THREADS);

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
      auto a_seek = [&](int it) {
        a_k = (wk.kb_begin + it) * BK + c * 4;
        uint32_t tap, ci;
        p.div_c.divmod(static_cast<uint32_t>(a_k < p.K ? a_k : 0), tap, ci);
        a_tap = static_cast<int>(tap);
        a_ci = static_cast<int>(ci);
        a_retap();
      };
      auto a_advance = [&]() {
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
      float4 va0[8], va1[8];
      auto body = [&](int it, float4 (&cur)[8], float4 (&nxt)[8]) {
        if (it + 2 < wk.nkb) {
          a_advance();
          load_row_a(nxt);
        }
        const int G = gbase + it;
        const int s = G % Cfg::STAGES;
        mbar_wait(&empty_bar[s], (((G / Cfg::STAGES) & 1) ^ 1));
        const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES, a_lo = a_hi + Cfg::A_BYTES;
        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[s]), 0);
        if (gtid < 32) {
          if (elect_one_sync()) {
            const uint32_t b_hi = a_hi + Cfg::A_BYTES * PREC;
            // the leader accounts for the weight bytes of both halves; each CTA fetches its own 64 rows
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::B_BYTES * PREC);
            tma_load_2d_pair(b_hi, &tm_b_hi, full_leader, (wk.kb_begin + it) * BK, nb0);
            if (PREC == 2) tma_load_2d_pair(b_hi + Cfg::B_BYTES, &tm_b_lo, full_leader, (wk.kb_begin + it) * BK, nb0);
          }
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t off = static_cast<uint32_t>((r0 + 16 * i) * 128) + swz;
          store_split_fast<PREC>(a_hi + off, a_lo + off, cur[i]);
        }
        fence_proxy_async();
        __syncwarp();  // every lane of the warp has written and fenced its rows
        if (lane == 0) mbar_arrive_cluster(full_leader);
      };
      if (group < wk.nkb) {
        a_seek(group);
        load_row_a(va0);
      }
      for (int it = group; it < wk.nkb; it += 4) {
        body(it, va0, va1);
        if (it + 2 < wk.nkb) body(it + 2, va1, va0);
      }
      gbase += wk.nkb;
    }
  } else {
    // ================================================================= MMA issuer (leader CTA, warp 12, one thread)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (rank == 0 && warp == WS_MMA_WARP) {  // converged loop; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = umma_idesc_tf32(2 * BM, BN, 0, 0);
      const uint64_t d_a_hi0 = umma_desc(smem_base, 16, 1024, 2);
      const uint64_t d_a_lo0 = umma_desc(smem_base + Cfg::A_BYTES, 16, 1024, 2);
      const uint64_t d_b_hi0 = umma_desc(smem_base + Cfg::A_BYTES * PREC, 16, 1024, 2);
      const uint64_t d_b_lo0 = umma_desc(smem_base + Cfg::A_BYTES * PREC + Cfg::B_BYTES, 16, 1024, 2);
      int gbase = 0, ubase = 0;
      for (int w = pair_id; w < total_work; w += num_pairs) {
        const Ws2Work wk = ws2_decode(p, w, m_pairs);
        // a promotion unit = p.unit_kb k-blocks (2 = 64 K-elements; short-K linear layers use half of their K extent,
        // see csrc/igemm.cu) accumulated in one TMEM buffer before the promotion warps add it into registers
        int uin = 0, ucur = 0;  // position inside the unit, unit index inside the tile
        for (int it = 0; it < wk.nkb; ++it) {
          const int G = gbase + it;
          const int s = G % Cfg::STAGES;
          const int U = ubase + (PROMOTE ? ucur : 0);
          const int b = U % NBUF;
          const bool unit_first = PROMOTE ? (uin == 0) : (it == 0);
          const bool unit_last = PROMOTE ? (uin == p.unit_kb - 1 || it == wk.nkb - 1) : (it == wk.nkb - 1);
          if (++uin == p.unit_kb) {
            uin = 0;
            ++ucur;
          }
          if (unit_first) {
            mbar_wait_cluster(&acce_bar[b], (((U / NBUF) & 1) ^ 1));
            tc_fence_after();
          }
          mbar_wait_cluster(&full_bar[s], (G / Cfg::STAGES) & 1);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t tacc = tmem_base + static_cast<uint32_t>(b * BN);
            const uint64_t soff = static_cast<uint64_t>((s * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
              const uint64_t off = soff + kk * 2;
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
        gbase += wk.nkb;
        ubase += PROMOTE ? ((wk.nkb + p.unit_kb - 1) / p.unit_kb) : 1;
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading TMEM / receiving commits until both CTAs are through
  if (warp == WS_MMA_WARP) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace cavp
