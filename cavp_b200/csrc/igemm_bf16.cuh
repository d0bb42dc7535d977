// bf16-operand version of the CTA-pair forward / dgrad implicit-GEMM kernel (igemm_ws2.cuh) - the "bf16 training" path
// of BASELINE.json configs[2]: tcgen05.mma.kind::f16 with bf16 A/B, fp32 accumulation in TMEM, fp32 activations and
// epilogue (BatchNorm statistics, LayerNorm, losses stay fp32).
//
// What changes against the 3xTF32 kernel:
//   * one MMA per k-step instead of three and no accumulator promotion: the whole K extent accumulates in TMEM
//     (2 accumulator buffers, so the epilogue of tile t overlaps the main loop of tile t+1); warps 0-3 only run epilogues;
//   * a k-block is still one 128-byte swizzle row per tile row, i.e. 64 bf16 K-elements (UMMA_K = 16 -> 4 MMAs);
//   * the activation tile is gathered as fp32 (2 x 16 bytes per 8-element chunk), converted with cvt.rn.bf16x2.f32 and
//     stored as one 16-byte shared-memory vector: a quarter of the shared-memory write traffic per K-element.  Both
//     producer groups work on EVERY k-block (group g owns rows 64g .. 64g+63; 4 chunks x 2 float4 per thread and
//     buffer, the same 64 registers of double buffer as the fp32 kernel);
//   * the weight operand is a bf16 copy of the K-major matrix ([N][ldw] bf16, refreshed once per step by
//     cavp_cvt_bf16_multi / cavp_transpose_bf16_multi) fetched by TMA with a 64-element box;
//   * 256-column pair tiles (each CTA holds 128 weight rows) where N is a multiple of 256: the activation tile is
//     amortised over twice the columns - with 6x less MMA time per K-element the gather is what bounds this kernel.
// Channel counts must be multiples of 8 (a chunk never straddles a filter tap); the dispatcher falls back to the TF32
// kernels otherwise (the 3/4-channel stems).
#pragma once
#include "igemm_ws2.cuh"

namespace cavp {

constexpr int BK16 = 64;      // bf16 K-elements per k-block (one 128-byte row)
constexpr int UMMA_K16 = 16;  // kind::f16

template <int BN>
struct Bf16Cfg {
  static constexpr int NBUF = 512 / BN >= 4 ? 4 : 512 / BN;
  static constexpr int BH = BN / 2;  // weight rows held by one CTA
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BH * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN > 128 ? 6 : 7;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int ROWTAB_BYTES = BM * 8;
  static constexpr int SCRATCH_BYTES = WS_EPI_WARPS * 4608;
  static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + ROWTAB_BYTES + SCRATCH_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(2 * STAGES + 2 * NBUF + 1 <= BAR_BYTES / 8, "barrier area");
};

// two fp32 -> packed bf16x2 (lo half = first argument: lower address in shared memory), round to nearest even
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void st_shared_v4_b32(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 32-bit instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate (same fields as umma_idesc_tf32;
// A/B format 1 = bf16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WS_THREADS, 1)
igemm_bf16_pair_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_b, int total_work, int m_pairs) {
  using Cfg = Bf16Cfg<BN>;
  constexpr int NBUF = Cfg::NBUF;
  static_assert(BN == 128 || BN == 256, "bf16 pair kernel: 128- or 256-column tiles");

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
      mbar_init(&full_bar[s], 2 * (PRODUCER_THREADS / 32) + 1);  // every producer warp of both CTAs + the expect_tx
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], 2 * WS_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (tid == 32) tma_prefetch_desc(&tm_b);
  if (warp == WS_MMA_WARP) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < WS_EPI_WARPS) {
    // ================================================================= epilogue (thread = tile row)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t scratch = scratch_base + static_cast<uint32_t>(warp * 4608);
    int U = 0;
    for (int w = pair_id; w < total_work; w += num_pairs, ++U) {
      const Ws2Work wk = ws2_decode(p, w, m_pairs);
      const int m_tile = wk.m_pair * 2 + static_cast<int>(rank);
      const int b = U % NBUF;
      mbar_wait(&accf_bar[b], (U / NBUF) & 1);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < BN / 128; ++half) {
        float acc[128];
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(b * BN + half * 128 + cg * 32), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cg * 32 + j] = v[j];
        }
        if (half == BN / 128 - 1) {  // the accumulator buffer is free as soon as its last column left TMEM
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acce_bar[b]), 0));
        }
        igemm_epilogue<128>(p, acc, m_tile * BM, wk.n_tile * BN + half * 128, m_tile, 0, q, lane, scratch, wk.split);
      }
    }
  } else if (warp < WS_MMA_WARP) {
    // ================================================================= producers (both groups on every k-block)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    const int ptid = tid - WS_PROD_WARP0 * 32;  // 0..255
    const int c = ptid & 7;                     // 16-byte bf16 chunk = K-elements [8c, 8c+8) of the k-block
    const int r0 = ptid >> 3;                   // rows r0 + 32 i, i < 4
    const uint32_t swz = static_cast<uint32_t>((c ^ (r0 & 7)) << 4);
    int gbase = 0;
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

      int a_off[4];
      int a_k = 0, a_ci = 0, a_tap = 0;
      auto a_retap = [&]() {
        uint32_t ky, kx;
        p.div_s.divmod(static_cast<uint32_t>(a_tap), ky, kx);
        const int dy = static_cast<int>(ky) * p.dil;
        const int dx = static_cast<int>(kx) * p.dil;
        const bool kvalid = a_k < p.K;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int iy, ix;
          const int2 ri = rowtab[r0 + 32 * i];
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
        a_k = (wk.kb_begin + it) * BK16 + c * 8;
        uint32_t tap, ci;
        p.div_c.divmod(static_cast<uint32_t>(a_k < p.K ? a_k : 0), tap, ci);
        a_tap = static_cast<int>(tap);
        a_ci = static_cast<int>(ci);
        a_retap();
      };
      auto a_advance = [&]() {
        a_k += BK16;
        a_ci += BK16;
        if (a_ci >= p.C || a_k >= p.K) {
          while (a_ci >= p.C) {
            a_ci -= p.C;
            ++a_tap;
          }
          a_retap();
        }
      };
      auto load_rows = [&](float4 (&va)[8]) {
        const float* base = p.x + a_ci;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (a_off[i] >= 0) {
            va[2 * i] = ldg_nc_v4(base + a_off[i]);
            va[2 * i + 1] = ldg_nc_v4(base + a_off[i] + 4);
          } else {
            va[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
            va[2 * i + 1] = va[2 * i];
          }
        }
      };
      float4 va0[8], va1[8];
      auto body = [&](int it, float4 (&cur)[8], float4 (&nxt)[8]) {
        if (it + 1 < wk.nkb) {
          a_advance();
          load_rows(nxt);
        }
        const int G = gbase + it;
        const int s = G % Cfg::STAGES;
        mbar_wait(&empty_bar[s], (((G / Cfg::STAGES) & 1) ^ 1));
        const uint32_t a_st = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[s]), 0);
        if (ptid < 32) {
          if (elect_one_sync()) {
            // the leader accounts for the weight bytes of both halves; each CTA fetches its own BN/2 rows
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::B_BYTES);
            tma_load_2d_pair(a_st + Cfg::A_BYTES, &tm_b, full_leader, (wk.kb_begin + it) * BK16, nb0);
          }
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = static_cast<uint32_t>((r0 + 32 * i) * 128) + swz;
          const float4 lo = cur[2 * i], hi = cur[2 * i + 1];
          st_shared_v4_b32(a_st + off, cvt_bf16x2(lo.x, lo.y), cvt_bf16x2(lo.z, lo.w), cvt_bf16x2(hi.x, hi.y),
                           cvt_bf16x2(hi.z, hi.w));
        }
        fence_proxy_async();
        __syncwarp();  // every lane of the warp has written and fenced its rows
        if (lane == 0) mbar_arrive_cluster(full_leader);
      };
      if (wk.nkb > 0) {
        a_seek(0);
        load_rows(va0);
      }
      for (int it = 0; it < wk.nkb; it += 2) {
        body(it, va0, va1);
        if (it + 1 < wk.nkb) body(it + 1, va1, va0);
      }
      gbase += wk.nkb;
    }
  } else {
    // ================================================================= MMA issuer (leader CTA, warp 12)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (rank == 0 && warp == WS_MMA_WARP) {  // converged loop; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      const uint64_t d_a0 = umma_desc(smem_base, 16, 1024, 2);
      const uint64_t d_b0 = umma_desc(smem_base + Cfg::A_BYTES, 16, 1024, 2);
      int gbase = 0, U = 0;
      for (int w = pair_id; w < total_work; w += num_pairs, ++U) {
        const Ws2Work wk = ws2_decode(p, w, m_pairs);
        const int b = U % NBUF;
        mbar_wait_cluster(&acce_bar[b], (((U / NBUF) & 1) ^ 1));
        tc_fence_after();
        for (int it = 0; it < wk.nkb; ++it) {
          const int G = gbase + it;
          const int s = G % Cfg::STAGES;
          mbar_wait_cluster(&full_bar[s], (G / Cfg::STAGES) & 1);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t tacc = tmem_base + static_cast<uint32_t>(b * BN);
            const uint64_t soff = static_cast<uint64_t>((s * Cfg::STAGE_BYTES) >> 4);
#pragma unroll
            for (int kk = 0; kk < BK16 / UMMA_K16; ++kk) {
              const uint64_t off = soff + kk * 2;
              mma_bf16_ss_pair(tacc, d_a0 + off, d_b0 + off, idesc, !(it == 0 && kk == 0));
            }
            tc_commit_pair(&empty_bar[s], 3);
            if (it == wk.nkb - 1) tc_commit_pair(&accf_bar[b], 3);
          }
          __syncwarp();
        }
        gbase += wk.nkb;
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
