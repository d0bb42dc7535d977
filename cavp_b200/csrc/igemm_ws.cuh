// Persistent, fully warp-specialised forward / dgrad implicit-GEMM kernel (weights by TMA, activations gathered).
//
// Same math, operand layouts and epilogue as igemm_kernel<BN, PREC, MODE_ROW, BTMA=true> (igemm.cuh); what changes is
// the schedule.  In igemm_kernel the producer warps also own the fp32 accumulators, so the pipeline fills and drains
// once per tile and nothing overlaps the epilogue - expensive for the many short-K GEMMs of the fusion block (K = 304:
// 10 k-blocks per tile).  Here one CTA per SM loops over tiles and the roles never stop:
//
//   warps  0-3  (WG0, 232 regs): promotion + epilogue.  Thread = tile row; pulls every finished 64-wide K unit out of the
//                                4-deep TMEM ring (tcgen05.ld), adds it round-to-nearest into BN fp32 registers, runs the
//                                epilogue of tile t while the MMAs of tile t+1 already fill the ring.
//   warps  4-11 (WG1-2, 112 regs): producers, two groups alternating k-blocks: cached-offset gather -> register double
//                                buffer -> TF32 hi/lo split -> swizzled smem; one thread per group issues the weight TMA.
//   warp   12   (WG3, 40 regs):  MMA issuer.  warps 13-15 idle (they only exist so that setmaxnreg can rebalance registers
//                                between warpgroups).
// The smem stage ring, the TMEM ring and all mbarrier phases run continuously across tiles (global k-block / unit counters).
#pragma once
#include "igemm.cuh"

namespace cavp {

constexpr int WS_THREADS = 512;
constexpr int WS_EPI_WARPS = 4;
constexpr int WS_PROD_WARP0 = 4;
constexpr int WS_MMA_WARP = 12;

template <int BN, int PREC>
struct WsCfg {
  static constexpr bool PROMOTE = (PREC == 2);
  static constexpr int NBUF = 512 / BN >= 4 ? 4 : 512 / BN;
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * PREC;
  static constexpr int STAGES = (PREC == 2) ? (BN >= 128 ? 3 : 4) : 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int ROWTAB_BYTES = BM * 8;
  static constexpr int SCRATCH_BYTES = WS_EPI_WARPS * 4608;
  static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + ROWTAB_BYTES + SCRATCH_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct WsWork {
  int m_tile, n_tile, kb_begin, nkb, split;
};
__device__ __forceinline__ WsWork ws_decode(const IgemmParams& p, int w) {
  // work item = (tile, split); tiles ordered n-fastest so that concurrently running CTAs share their A rows in L2
  WsWork r;
  const int tiles = p.n_tiles * ((p.M + BM - 1) / BM);
  const int split = w / tiles;
  const int tile = w - split * tiles;
  r.n_tile = tile % p.n_tiles;
  r.m_tile = tile / p.n_tiles;
  r.split = split;
  r.kb_begin = static_cast<int>((static_cast<long long>(p.num_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.num_kb) * (split + 1)) / p.splits);
  r.nkb = kb_end - r.kb_begin;
  return r;
}

template <int BN, int PREC>
__global__ void __launch_bounds__(WS_THREADS, 1)
igemm_ws_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_b_hi,
                const __grid_constant__ CUtensorMap tm_b_lo, int total_work) {
  using Cfg = WsCfg<BN, PREC>;
  constexpr bool PROMOTE = Cfg::PROMOTE;
  constexpr int NBUF = Cfg::NBUF;
  static_assert(BN == 64 || BN == 128, "BN must be 64 or 128");

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

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], GROUP_THREADS + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], WS_EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (tid == 32) {
    tma_prefetch_desc(&tm_b_hi);
    if (PREC == 2) tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == WS_MMA_WARP) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < WS_EPI_WARPS) {
    // ================================================================= promotion + epilogue (thread = tile row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t scratch = scratch_base + static_cast<uint32_t>(warp * 4608);
    int ubase = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WsWork wk = ws_decode(p, w);
      const int nunits = PROMOTE ? ((wk.nkb + 1) >> 1) : 1;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      for (int u = 0; u < nunits; ++u) {
        const int U = ubase + u;
        const int b = U & (NBUF - 1);
        mbar_wait(&accf_bar[b], (U / NBUF) & 1);
        tc_fence_after();
#pragma unroll
        for (int cg = 0; cg < BN / 32; ++cg) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(b * BN + cg * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cg * 32 + j] += v[j];
        }
        tc_fence_before();
        mbar_arrive(&acce_bar[b]);
      }
      ubase += nunits;
      igemm_epilogue<BN>(p, acc, wk.m_tile * BM, wk.n_tile * BN, wk.m_tile, 0, q, lane, scratch, wk.split);
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
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WsWork wk = ws_decode(p, w);
      const int m0 = wk.m_tile * BM, n0 = wk.n_tile * BN;
      // ---- row table of this tile (all producers have finished reading the previous one)
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
        if (gtid < 32) {
          if (elect_one_sync()) {
            const uint32_t b_hi = a_hi + Cfg::A_BYTES * PREC;
            mbar_arrive_expect_tx(&full_bar[s], Cfg::B_BYTES * PREC);
            tma_load_2d(b_hi, &tm_b_hi, &full_bar[s], (wk.kb_begin + it) * BK, n0);
            if (PREC == 2) tma_load_2d(b_hi + Cfg::B_BYTES, &tm_b_lo, &full_bar[s], (wk.kb_begin + it) * BK, n0);
          }
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t off = static_cast<uint32_t>((r0 + 16 * i) * 128) + swz;
          store_split_fast<PREC>(a_hi + off, a_lo + off, cur[i]);
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
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
    // ================================================================= MMA issuer (warp 12, one thread) + idle warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == WS_MMA_WARP) {  // converged loop; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, 0, 0);
      const uint64_t d_a_hi0 = umma_desc(smem_base, 16, 1024, 2);
      const uint64_t d_a_lo0 = umma_desc(smem_base + Cfg::A_BYTES, 16, 1024, 2);
      const uint64_t d_b_hi0 = umma_desc(smem_base + Cfg::A_BYTES * PREC, 16, 1024, 2);
      const uint64_t d_b_lo0 = umma_desc(smem_base + Cfg::A_BYTES * PREC + Cfg::B_BYTES, 16, 1024, 2);
      int gbase = 0, ubase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WsWork wk = ws_decode(p, w);
        for (int it = 0; it < wk.nkb; ++it) {
          const int G = gbase + it;
          const int s = G % Cfg::STAGES;
          const int U = ubase + (PROMOTE ? (it >> 1) : 0);
          const int b = U & (NBUF - 1);
          const bool unit_first = PROMOTE ? ((it & 1) == 0) : (it == 0);
          const bool unit_last = PROMOTE ? ((it & 1) == 1 || it == wk.nkb - 1) : (it == wk.nkb - 1);
          if (unit_first) {
            mbar_wait(&acce_bar[b], (((U / NBUF) & 1) ^ 1));
            tc_fence_after();
          }
          mbar_wait(&full_bar[s], (G / Cfg::STAGES) & 1);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t tacc = tmem_base + static_cast<uint32_t>(b * BN);
            const uint64_t soff = static_cast<uint64_t>((s * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
              const uint64_t off = soff + kk * 2;
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
        gbase += wk.nkb;
        ubase += PROMOTE ? ((wk.nkb + 1) >> 1) : 1;
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WS_MMA_WARP) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace cavp
