// 256-column CTA-pair tile for the fp32-parity (3xTF32 + promotion) forward / dgrad GEMMs whose N is a multiple of 256.
//
// Why: with 128-column pair tiles (igemm_ws2.cuh) every k-block moves 120 KB through an SM's shared memory (32 KB of
// hi | lo activation-tile writes, 16 KB of weight TMA, 72 KB of tensor-core operand reads) against 768 clk of MMA - the
// 128 B/clk shared-memory port, not the tensor pipe, sets the pace (measured 66 % tensor-pipe active, data pipe 85 %).
// ONE tcgen05.mma.cta_group::2 with M = 256, N = 256 (each CTA holds 128 rows of A and 128 weight rows) reads the
// activation tile once for twice the columns: 32 + 32 + 96 = 160 KB per 1536 clk of MMA, i.e. 0.81 of the port.
//
// What has to change for that: 256 fp32 accumulators per tile row do not fit one thread, so promotion and epilogue
// run on EIGHT warps (warp w and w + 4 share TMEM lane quarter w % 4 and own one 128-column half each, 176 registers)
// and the activation gather moves to FOUR producer warps that handle every k-block (the per-thread rate is unchanged:
// a k-block now lasts twice as long).  TMEM holds two 256-column accumulator units (a unit = 2 k-blocks = 64
// K-elements, pulled out and added round-to-nearest into registers like in the 128-column kernel).  Shared memory:
// 3 stages x 64 KB + eight 32 x 16 epilogue scratch pads; the epilogue works on 16-column chunks for that reason and
// supports what the N % 256 == 0 layers need (scale / shift / ReLU / LeakyReLU, BatchNorm partials, split-K slabs or
// red.add, in-place gradient accumulation) - launches with other epilogue options stay on the 128-column kernel.
//
// Cross-CTA protocol as in igemm_ws2.cuh: full[s] in the leader (one arrival per producer warp of both CTAs + the
// leader's expect_tx for both weight halves), empty[s] / accf[b] armed in both CTAs by multicast commits, acce[b] in the
// leader (one arrival per promotion warp of both CTAs).
#pragma once
#include "igemm_ws2.cuh"

namespace cavp {

constexpr int WX_BN = 256;
constexpr int WX_EPI_WARPS = 8;
constexpr int WX_PROD_WARP0 = 8;
constexpr int WX_PROD_THREADS = 128;
constexpr int WX_MMA_WARP = 12;
constexpr int WX_LDS = 20;  // scratch row stride in floats (16 columns + pad; 16-byte aligned rows)

struct WxCfg {
  static constexpr int NBUF = 2;
  static constexpr int BH = WX_BN / 2;  // weight rows held by one CTA
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BH * 128;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * 2;
  static constexpr int STAGES = 3;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int ROWTAB_BYTES = BM * 8;
  static constexpr int SCRATCH_PER_WARP = 32 * WX_LDS * 4;
  static constexpr int SCRATCH_BYTES = WX_EPI_WARPS * SCRATCH_PER_WARP;
  static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + ROWTAB_BYTES + SCRATCH_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(2 * STAGES + 2 * NBUF + 1 <= BAR_BYTES / 8, "barrier area");
};

// can the 256-column kernel's epilogue serve this launch?
__host__ __device__ __forceinline__ bool ws2x_epilogue_ok(const IgemmParams& p) {
  const bool plain_res = p.res == nullptr || igemm_inplace_acc(p);
  return plain_res && p.y_pre == nullptr && (p.act == ACT_NONE || p.act == ACT_RELU || p.act == ACT_LEAKY) &&
         (p.Ncols % WX_BN) == 0 && (p.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0;
}

// Epilogue of one thread's 128 accumulators (row = m0 + q*32 + lane, columns n0 .. n0+127) in 16-column chunks.
__device__ __forceinline__ void ws2x_epilogue(const IgemmParams& pin, float (&acc)[128], int m0, int n0, int m_tile,
                                              int q, int lane, uint32_t scratch, int split) {
  if (m0 >= pin.M) return;
  IgemmParams p = pin;
  if (pin.splits > 1 && pin.split_slab > 0) {  // deterministic split-K: private slab, plain store
    p.y = pin.y + static_cast<size_t>(split) * pin.split_slab;
    p.splits = 1;
  }
  const int row0 = m0 + q * 32;
  const bool accumulate = p.splits > 1 || igemm_inplace_acc(p);
  const int c4 = lane & 3, rsub = lane >> 2;  // store phase: 4 lanes per row (16 columns), 8 rows per pass
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    const int col0 = n0 + ch * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = acc[ch * 16 + j];
    if (!accumulate) {
      if (p.scale) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= __ldg(p.scale + col0 + j);
      }
      if (p.shift) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += __ldg(p.shift + col0 + j);
      }
      if (p.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (p.act == ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      st_shared_v4(scratch + static_cast<uint32_t>((lane * WX_LDS + j) * 4), v[j], v[j + 1], v[j + 2], v[j + 3]);
    __syncwarp();
    float* ybase = p.y + static_cast<size_t>(row0) * p.ldy + col0;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = rr * 8 + rsub;
      const float4 t = lds_v4(scratch + static_cast<uint32_t>((r * WX_LDS + c4 * 4) * 4));
      if (row0 + r < p.M) {
        float* dst = ybase + static_cast<size_t>(r) * p.ldy + c4 * 4;
        if (accumulate)
          red_add_v4(dst, t.x, t.y, t.z, t.w);
        else
          *reinterpret_cast<float4*>(dst) = t;
      }
    }
    if (p.stats && !accumulate) {
      // column sums from the staged chunk: lane = (row half, column): 16 rows each, then one shuffle
      const int col = lane & 15, half = lane >> 4;
      const int nrow = p.M - row0 < 32 ? p.M - row0 : 32;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
      for (int r = half * 16; r < half * 16 + 16; ++r) {
        if (r < nrow) {
          const float t = lds_f32(scratch + static_cast<uint32_t>((r * WX_LDS + col) * 4));
          s1 += t;
          s2 = fmaf(t, t, s2);
        }
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
      if (half == 0) {
        float* st = p.stats + static_cast<size_t>(m_tile * 4 + q) * 2 * p.ldstat;
        st[col0 + col] = s1;
        st[p.ldstat + col0 + col] = s2;
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WS_THREADS, 1)
igemm_ws2x_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_b_hi,
                  const __grid_constant__ CUtensorMap tm_b_lo, int total_work, int m_pairs) {
  using Cfg = WxCfg;
  constexpr int NBUF = Cfg::NBUF;
  constexpr int BN = WX_BN;

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
      mbar_init(&full_bar[s], 2 * (WX_PROD_THREADS / 32) + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], 2 * WX_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (tid == 32) {
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == WX_MMA_WARP) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < WX_EPI_WARPS) {
    // ================================================================= promotion + epilogue (warpgroups 0 and 1)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    const int q = warp & 3;       // TMEM lane quarter = tile rows q*32 .. q*32+31
    const int half = warp >> 2;   // column half: 0 -> 0..127, 1 -> 128..255
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t scratch = scratch_base + static_cast<uint32_t>(warp * Cfg::SCRATCH_PER_WARP);
    int ubase = 0;
    for (int w = pair_id; w < total_work; w += num_pairs) {
      const Ws2Work wk = ws2_decode(p, w, m_pairs);
      const int m_tile = wk.m_pair * 2 + static_cast<int>(rank);
      const int nunits = (wk.nkb + 1) >> 1;
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      for (int u = 0; u < nunits; ++u) {
        const int U = ubase + u;
        const int b = U & (NBUF - 1);
        mbar_wait(&accf_bar[b], (U / NBUF) & 1);
        tc_fence_after();
#pragma unroll
        for (int cg = 0; cg < 8; ++cg) {
          float v[16];
          tmem_ld16(tmem_base + lane_base + static_cast<uint32_t>(b * BN + half * 128 + cg * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[cg * 16 + j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acce_bar[b]), 0));
      }
      ubase += nunits;
      ws2x_epilogue(p, acc, m_tile * BM, wk.n_tile * BN + half * 128, m_tile, q, lane, scratch, wk.split);
    }
  } else if (warp < WX_MMA_WARP) {
    // ================================================================= producers (4 warps, every k-block)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    const int gtid = tid - WX_PROD_WARP0 * 32;  // 0..127
    const int c = gtid & 7;
    const int r0 = gtid >> 3;
    const uint32_t swz = static_cast<uint32_t>((c ^ (r0 & 7)) << 4);
    int gbase = 0;
    for (int w = pair_id; w < total_work; w += num_pairs) {
      const Ws2Work wk = ws2_decode(p, w, m_pairs);
      const int m0 = (wk.m_pair * 2 + static_cast<int>(rank)) * BM;
      const int nb0 = wk.n_tile * BN + static_cast<int>(rank) * Cfg::BH;
      named_bar_sync(1, WX_PROD_THREADS);
      {
        const int m = m0 + gtid;
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
        rowtab[gtid] = e;
      }
      named_bar_sync(1, WX_PROD_THREADS);

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
        a_k += BK;
        a_ci += BK;
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
        if (it + 1 < wk.nkb) {
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
            const uint32_t b_hi = a_hi + Cfg::A_BYTES * 2;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::B_BYTES * 2);
            tma_load_2d_pair(b_hi, &tm_b_hi, full_leader, (wk.kb_begin + it) * BK, nb0);
            tma_load_2d_pair(b_hi + Cfg::B_BYTES, &tm_b_lo, full_leader, (wk.kb_begin + it) * BK, nb0);
          }
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t off = static_cast<uint32_t>((r0 + 16 * i) * 128) + swz;
          store_split_fast<2>(a_hi + off, a_lo + off, cur[i]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(full_leader);
      };
      a_seek(0);
      load_row_a(va0);
      for (int it = 0; it < wk.nkb; it += 2) {
        body(it, va0, va1);
        if (it + 1 < wk.nkb) body(it + 1, va1, va0);
      }
      gbase += wk.nkb;
    }
  } else {
    // ================================================================= MMA issuer (leader CTA, warp 12)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (rank == 0 && warp == WX_MMA_WARP) {
      constexpr uint32_t idesc = umma_idesc_tf32(2 * BM, BN, 0, 0);
      const uint64_t d_a_hi0 = umma_desc(smem_base, 16, 1024, 2);
      const uint64_t d_a_lo0 = umma_desc(smem_base + Cfg::A_BYTES, 16, 1024, 2);
      const uint64_t d_b_hi0 = umma_desc(smem_base + Cfg::A_BYTES * 2, 16, 1024, 2);
      const uint64_t d_b_lo0 = umma_desc(smem_base + Cfg::A_BYTES * 2 + Cfg::B_BYTES, 16, 1024, 2);
      int gbase = 0, ubase = 0;
      for (int w = pair_id; w < total_work; w += num_pairs) {
        const Ws2Work wk = ws2_decode(p, w, m_pairs);
        for (int it = 0; it < wk.nkb; ++it) {
          const int G = gbase + it;
          const int s = G % Cfg::STAGES;
          const int U = ubase + (it >> 1);
          const int b = U & (NBUF - 1);
          const bool unit_first = (it & 1) == 0;
          const bool unit_last = (it & 1) == 1 || it == wk.nkb - 1;
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
              mma_tf32_ss_pair(tacc, d_a_lo0 + off, d_b_hi0 + off, idesc, 1);
              mma_tf32_ss_pair(tacc, d_a_hi0 + off, d_b_lo0 + off, idesc, 1);
            }
            tc_commit_pair(&empty_bar[s], 3);
            if (unit_last) tc_commit_pair(&accf_bar[b], 3);
          }
          __syncwarp();
        }
        gbase += wk.nkb;
        ubase += (wk.nkb + 1) >> 1;
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == WX_MMA_WARP) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace cavp
