// CTA-pair kernel for the short-K LINEAR layers of the fusion block (1x1, stride 1: Linear 304 <-> 1216 / 304 / 256 over
// 100-200 K tokens; fp32-parity mode).
//
// Why a separate schedule: with K = 304 a tile has only 10 k-blocks, so per tile the promotion + epilogue work (pull
// the accumulator units out of TMEM, add, transpose through shared memory, store 80 KB) is as long as the main loop.  In
// igemm_ws2_kernel those four warps were ~100 % busy while producers and the MMA warp waited (ncu source view:
// profiles/r02_ncu_full_ws2_128_fc2dgrad_k304_after.csv, tensor pipe 53 %).  Here that role gets EIGHT warps (warp w and
// w + 4 share TMEM lane quarter w % 4 and own half of the tile's columns each, as in igemm_ws2x_kernel) while the
// producer side keeps its eight warps and the linear fast path of igemm_ws2_kernel (no row table, the register double
// buffer runs across tile boundaries): 20 warps per CTA, registers rebalanced with setmaxnreg (120 / 96 / 40; the CTA is
// launched with 640 x 96 registers, and 256 x 120 + 256 x 96 + 128 x 40 stays inside that allocation).
//
// Tiles are 256 rows x BN columns per CTA pair (BN = 128 or 160: 160 for N = 304 / 1216); the epilogue is the 16-column
// chunk epilogue of igemm_ws2x.cuh (bias / folded BN / ReLU / LeakyReLU in the store phase, or BatchNorm partials of a raw
// conv output, or in-place accumulate; column tails in multiples of 16).  Launches that need anything else stay on igemm_ws2_kernel.
#pragma once
#include "igemm_ws2x.cuh"

namespace cavp {

constexpr int LIN_THREADS = 640;
constexpr int LIN_EPI_WARPS = 8;
constexpr int LIN_PROD_WARP0 = 8;
constexpr int LIN_MMA_WARP = 16;

template <int BN>
struct LinCfg {
  static constexpr int NBUF = 512 / BN >= 4 ? 4 : 512 / BN;
  static constexpr int BH = BN / 2;   // weight rows held by one CTA
  static constexpr int HC = BN / 2;   // accumulator columns owned by one epilogue warp
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BH * 128;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * 2;
  static constexpr int STAGES = BN > 128 ? 3 : 4;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SCRATCH_PER_WARP = 32 * WX_LDS * 4;
  static constexpr int SCRATCH_BYTES = LIN_EPI_WARPS * SCRATCH_PER_WARP;
  static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + SCRATCH_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;
  static_assert(HC % 16 == 0, "16-column epilogue chunks");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(2 * STAGES + 2 * NBUF + 1 <= BAR_BYTES / 8, "barrier area");
};

// can this launch run on the linear kernel?  (shape: the dispatcher; epilogue: here)
__host__ __device__ __forceinline__ bool lin_epilogue_ok(const IgemmParams& p) {
  const bool plain_res = p.res == nullptr || igemm_inplace_acc(p);
  return plain_res && p.y_pre == nullptr && (p.act == ACT_NONE || p.act == ACT_RELU || p.act == ACT_LEAKY) &&
         (p.stats == nullptr || (p.scale == nullptr && p.shift == nullptr && p.act == ACT_NONE)) &&
         (p.Ncols % 16) == 0 && (p.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 && p.splits == 1;
}

// ws2x_epilogue generalised to HC columns per thread with a column tail (whole 16-column chunks)
template <int HC>
__device__ __forceinline__ void lin_epilogue(const IgemmParams& p, float (&acc)[HC], int m0, int n0, int m_tile, int q,
                                             int lane, uint32_t scratch) {
  if (m0 >= p.M) return;
  const int row0 = m0 + q * 32;
  const bool accumulate = igemm_inplace_acc(p);
  const int c4 = lane & 3, rsub = lane >> 2;
#pragma unroll
  for (int ch = 0; ch < HC / 16; ++ch) {
    const int col0 = n0 + ch * 16;
    if (col0 >= p.Ncols) continue;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = acc[ch * 16 + j];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      st_shared_v4(scratch + static_cast<uint32_t>((lane * WX_LDS + j) * 4), v[j], v[j + 1], v[j + 2], v[j + 3]);
    __syncwarp();
    // per-column work in the store phase: a lane owns one column quad of the chunk
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!accumulate) {
      if (p.scale) sc = __ldg(reinterpret_cast<const float4*>(p.scale + col0) + c4);
      if (p.shift) sh = __ldg(reinterpret_cast<const float4*>(p.shift + col0) + c4);
    }
    float* ybase = p.y + static_cast<size_t>(row0) * p.ldy + col0;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = rr * 8 + rsub;
      float4 t = lds_v4(scratch + static_cast<uint32_t>((r * WX_LDS + c4 * 4) * 4));
      if (row0 + r < p.M) {
        float* dst = ybase + static_cast<size_t>(r) * p.ldy + c4 * 4;
        if (accumulate) {
          red_add_v4(dst, t.x, t.y, t.z, t.w);
        } else {
          t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y); t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
          if (p.act == ACT_RELU) {
            t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
          } else if (p.act == ACT_LEAKY) {
            t.x = t.x > 0.f ? t.x : t.x * p.slope; t.y = t.y > 0.f ? t.y : t.y * p.slope;
            t.z = t.z > 0.f ? t.z : t.z * p.slope; t.w = t.w > 0.f ? t.w : t.w * p.slope;
          }
          *reinterpret_cast<float4*>(dst) = t;
        }
      }
    }
    if (p.stats && !accumulate) {
      // BatchNorm partials of a train-mode conv (no per-column work in that case: the staged values are final):
      // lane = (row half, column), 16 rows each, then one shuffle - as in ws2x_epilogue
      const int col = lane & 15, hf = lane >> 4;
      const int nrow = p.M - row0 < 32 ? p.M - row0 : 32;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
      for (int r = hf * 16; r < hf * 16 + 16; ++r) {
        if (r < nrow) {
          const float t = lds_f32(scratch + static_cast<uint32_t>((r * WX_LDS + col) * 4));
          s1 += t;
          s2 = fmaf(t, t, s2);
        }
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
      if (hf == 0) {
        float* st = p.stats + static_cast<size_t>(m_tile * 4 + q) * 2 * p.ldstat;
        st[col0 + col] = s1;
        st[p.ldstat + col0 + col] = s2;
      }
    }
  }
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LIN_THREADS, 1)
igemm_lin_kernel(const IgemmParams p, const __grid_constant__ CUtensorMap tm_b_hi,
                 const __grid_constant__ CUtensorMap tm_b_lo, int total_work, int m_pairs) {
  using Cfg = LinCfg<BN>;
  constexpr int NBUF = Cfg::NBUF;
  constexpr int HC = Cfg::HC;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + Cfg::RING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* accf_bar = bars + 2 * Cfg::STAGES;
  uint64_t* acce_bar = bars + 2 * Cfg::STAGES + NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 2 * NBUF);
  const uint32_t scratch_base = smem_base + Cfg::RING_BYTES + Cfg::BAR_BYTES;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int nkb = p.num_kb;
  const int UK = p.unit_kb;  // k-blocks per promotion unit

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 2 * (GROUP_THREADS / 32) + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&accf_bar[b], 1);
      mbar_init(&acce_bar[b], 2 * LIN_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (tid == 32) {
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
  }
  if (warp == LIN_MMA_WARP) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < LIN_EPI_WARPS) {
    // ================================================================= promotion + epilogue (warpgroups 0 and 1)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    const int q = warp & 3;
    const int half = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t scratch = scratch_base + static_cast<uint32_t>(warp * Cfg::SCRATCH_PER_WARP);
    const int nunits = (nkb + UK - 1) / UK;
    int ubase = 0;
    for (int w = pair_id; w < total_work; w += num_pairs) {
      const Ws2Work wk = ws2_decode(p, w, m_pairs);
      const int m_tile = wk.m_pair * 2 + static_cast<int>(rank);
      float acc[HC];
#pragma unroll
      for (int j = 0; j < HC; ++j) acc[j] = 0.f;
      for (int u = 0; u < nunits; ++u) {
        const int U = ubase + u;
        const int b = U % NBUF;
        mbar_wait(&accf_bar[b], (U / NBUF) & 1);
        tc_fence_after();
#pragma unroll
        for (int cg = 0; cg < HC / 16; ++cg) {
          float v[16];
          tmem_ld16(tmem_base + lane_base + static_cast<uint32_t>(b * BN + half * HC + cg * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[cg * 16 + j] += v[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acce_bar[b]), 0));
      }
      ubase += nunits;
      lin_epilogue<HC>(p, acc, m_tile * BM, wk.n_tile * BN + half * HC, m_tile, q, lane, scratch);
    }
  } else if (warp < LIN_MMA_WARP) {
    // ================================================================= producers: the linear fast path of igemm_ws2
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    const int ptid = tid - LIN_PROD_WARP0 * 32;  // 0..255
    const int group = ptid >> 7;
    const int gtid = ptid & (GROUP_THREADS - 1);
    const int c = gtid & 7;
    const int r0 = gtid >> 3;
    const uint32_t swz = static_cast<uint32_t>((c ^ (r0 & 7)) << 4);
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
    int G = group;
    auto body = [&](Pos& cur_pos, float4 (&cur)[8], float4 (&nxt)[8]) {
      const Pos nxt_pos = advance(cur_pos);
      if (nxt_pos.w < total_work) load_lin(nxt_pos, nxt);
      const int s = G % Cfg::STAGES;
      mbar_wait(&empty_bar[s], (((G / Cfg::STAGES) & 1) ^ 1));
      const uint32_t a_hi = smem_base + s * Cfg::STAGE_BYTES, a_lo = a_hi + Cfg::A_BYTES;
      const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[s]), 0);
      if (gtid < 32) {
        if (elect_one_sync()) {
          const uint32_t b_hi = a_hi + Cfg::A_BYTES * 2;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::B_BYTES * 2);
          tma_load_2d_pair(b_hi, &tm_b_hi, full_leader, cur_pos.it * BK, cur_pos.nb0);
          tma_load_2d_pair(b_hi + Cfg::B_BYTES, &tm_b_lo, full_leader, cur_pos.it * BK, cur_pos.nb0);
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
  } else {
    // ================================================================= MMA issuer (leader CTA, warp 16)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (rank == 0 && warp == LIN_MMA_WARP) {
      constexpr uint32_t idesc = umma_idesc_tf32(2 * BM, BN, 0, 0);
      const uint64_t d_a_hi0 = umma_desc(smem_base, 16, 1024, 2);
      const uint64_t d_a_lo0 = umma_desc(smem_base + Cfg::A_BYTES, 16, 1024, 2);
      const uint64_t d_b_hi0 = umma_desc(smem_base + Cfg::A_BYTES * 2, 16, 1024, 2);
      const uint64_t d_b_lo0 = umma_desc(smem_base + Cfg::A_BYTES * 2 + Cfg::B_BYTES, 16, 1024, 2);
      const int nunits = (nkb + UK - 1) / UK;
      int gbase = 0, ubase = 0;
      for (int w = pair_id; w < total_work; w += num_pairs) {
        int uin = 0, ucur = 0;
        for (int it = 0; it < nkb; ++it) {
          const int G = gbase + it;
          const int s = G % Cfg::STAGES;
          const int U = ubase + ucur;
          const int b = U % NBUF;
          const bool unit_first = uin == 0;
          const bool unit_last = uin == UK - 1 || it == nkb - 1;
          if (++uin == UK) {
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
              mma_tf32_ss_pair(tacc, d_a_lo0 + off, d_b_hi0 + off, idesc, 1);
              mma_tf32_ss_pair(tacc, d_a_hi0 + off, d_b_lo0 + off, idesc, 1);
            }
            tc_commit_pair(&empty_bar[s], 3);
            if (unit_last) tc_commit_pair(&accf_bar[b], 3);
          }
          __syncwarp();
        }
        gbase += nkb;
        ubase += nunits;
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == LIN_MMA_WARP) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace cavp
