// C-ABI launchers for the tcgen05 implicit-GEMM kernels (igemm.cuh).  See include/cavp_b200.h for the contract.
#include <cstring>
#include <cstdlib>
#include "igemm.cuh"
#include "igemm_ws.cuh"
#include "igemm_ws2.cuh"
#include "igemm_wgrad2.cuh"
#include "igemm_bf16.cuh"
#include "igemm_ws2x.cuh"
#include "igemm_wgrad_bf16.cuh"
#include "igemm_lin.cuh"
#include <cuda_bf16.h>
#include "../../include/cavp_b200.h"

namespace cavp {

// ---- TMA descriptors for the weight operand (driver entry point resolved through the runtime; no -lcuda needed)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}
// [rows][ld] fp32 matrix, box = 32 (K) x bn rows, 128-byte swizzle, out-of-bounds elements read as zero
static int make_weight_tmap(CUtensorMap* tm, const float* w, int rows, int K, int ld, int bn) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CAVP_ERR_ARG;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + static_cast<int>(r);
}

// dense [P][cout] fp32 matrix (the pre-split dY of a weight gradient), box = 32 channels x 32 pixels, 128-byte swizzle
// with 32-byte atoms = the shared-memory layout of an MN-major tf32 UMMA operand
static int make_dy_tmap(CUtensorMap* tm, const float* dy, long long P, int cout) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CAVP_ERR_ARG;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cout), static_cast<cuuint64_t>(P)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cout) * 4};
  cuuint32_t box[2] = {32, static_cast<cuuint32_t>(BK)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dy), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + static_cast<int>(r);
}

// cudaFuncSetAttribute is per device: one process may drive several GPUs (nn.DataParallel, cuda:1 after cuda:0)
static constexpr int MAX_DEVICES = 64;
static inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < MAX_DEVICES) ? d : 0;
}

template <int BN, int PREC, int MODE, bool BTMA>
static int launch_igemm(const IgemmParams& p, const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, cudaStream_t st) {
  using Cfg = TileCfg<BN, PREC>;
  auto kern = igemm_kernel<BN, PREC, MODE, BTMA>;
  static bool configured_dev[MAX_DEVICES] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  dim3 grid(static_cast<unsigned>(m_tiles * p.n_tiles), static_cast<unsigned>(p.splits), 1);
  kern<<<grid, CTA_THREADS, Cfg::SMEM_BYTES, st>>>(p, tm_hi, tm_lo);
  return static_cast<int>(cudaGetLastError());
}

template <int BN, int PREC>
static int launch_igemm_ws(const IgemmParams& p, const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, cudaStream_t st) {
  using Cfg = WsCfg<BN, PREC>;
  auto kern = igemm_ws_kernel<BN, PREC>;
  static bool configured_dev[MAX_DEVICES] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  const int total_work = m_tiles * p.n_tiles * p.splits;
  const int grid = total_work < 148 ? total_work : 148;
  kern<<<grid, WS_THREADS, Cfg::SMEM_BYTES, st>>>(p, tm_hi, tm_lo, total_work);
  return static_cast<int>(cudaGetLastError());
}

// CTA-pair kernel: one cluster of 2 per TPC, persistent over (256-row pair tile, n tile, split) work items
template <int BN, int PREC>
static int launch_igemm_ws2(const IgemmParams& p, const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, cudaStream_t st) {
  using Cfg = Ws2Cfg<BN, PREC>;
  auto kern = igemm_ws2_kernel<BN, PREC>;
  static int max_pairs_dev[MAX_DEVICES] = {};
  int& max_pairs = max_pairs_dev[current_device()];
  if (max_pairs == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148, 1, 1);
    cfg.blockDim = dim3(WS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);  // cluster size comes from the kernel's __cluster_dims__
    if (e != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 74;
    }
    max_pairs = n;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_work = m_pairs * p.n_tiles * p.splits;
  const int pairs = total_work < max_pairs ? total_work : max_pairs;
  kern<<<2 * pairs, WS_THREADS, Cfg::SMEM_BYTES, st>>>(p, tm_hi, tm_lo, total_work, m_pairs);
  return static_cast<int>(cudaGetLastError());
}

// 256-column CTA-pair kernel (fp32-parity mode, N % 256 == 0)
static int launch_igemm_ws2x(const IgemmParams& p, const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, cudaStream_t st) {
  auto kern = igemm_ws2x_kernel;
  static int max_pairs_dev[MAX_DEVICES] = {};
  int& max_pairs = max_pairs_dev[current_device()];
  if (max_pairs == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WxCfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148, 1, 1);
    cfg.blockDim = dim3(WS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = WxCfg::SMEM_BYTES;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 74;
    }
    max_pairs = n;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_work = m_pairs * p.n_tiles * p.splits;
  const int pairs = total_work < max_pairs ? total_work : max_pairs;
  kern<<<2 * pairs, WS_THREADS, WxCfg::SMEM_BYTES, st>>>(p, tm_hi, tm_lo, total_work, m_pairs);
  return static_cast<int>(cudaGetLastError());
}

// short-K linear layers: 8 promotion / epilogue warps + 8 producer warps (igemm_lin.cuh)
template <int BN>
static int launch_igemm_lin(const IgemmParams& p, const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, cudaStream_t st) {
  using Cfg = LinCfg<BN>;
  auto kern = igemm_lin_kernel<BN>;
  static int max_pairs_dev[MAX_DEVICES] = {};
  int& max_pairs = max_pairs_dev[current_device()];
  if (max_pairs == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148, 1, 1);
    cfg.blockDim = dim3(LIN_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 74;
    }
    max_pairs = n;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_work = m_pairs * p.n_tiles;
  const int pairs = total_work < max_pairs ? total_work : max_pairs;
  kern<<<2 * pairs, LIN_THREADS, Cfg::SMEM_BYTES, st>>>(p, tm_hi, tm_lo, total_work, m_pairs);
  return static_cast<int>(cudaGetLastError());
}

template <int PREC>
static int launch_wgrad2(const IgemmParams& p, const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, cudaStream_t st) {
  using Cfg = Wg2Cfg<PREC>;
  auto kern = igemm_wgrad2_kernel<PREC>;
  static bool configured_dev[MAX_DEVICES] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  const int m_pairs = (m_tiles + 1) / 2;
  dim3 grid(static_cast<unsigned>(2 * m_pairs * p.n_tiles), static_cast<unsigned>(p.splits), 1);
  kern<<<grid, CTA_THREADS, Cfg::SMEM_BYTES, st>>>(p, tm_hi, tm_lo);
  return static_cast<int>(cudaGetLastError());
}

// [rows][ld] bf16 matrix, box = 64 (K) x bn rows, 128-byte swizzle, out-of-bounds elements read as zero
static int make_weight_tmap_bf16(CUtensorMap* tm, const void* w, int rows, int K, int ld, int bn) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CAVP_ERR_ARG;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK16), static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + static_cast<int>(r);
}

// bf16 CTA-pair kernel: persistent over (256-row pair tile, n tile, split) work items, one cluster of 2 per TPC
template <int BN>
static int launch_igemm_bf16(IgemmParams& p, const void* w_bf16, cudaStream_t st) {
  using Cfg = Bf16Cfg<BN>;
  auto kern = igemm_bf16_pair_kernel<BN>;
  static int max_pairs_dev[MAX_DEVICES] = {};
  int& max_pairs = max_pairs_dev[current_device()];
  if (max_pairs == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148, 1, 1);
    cfg.blockDim = dim3(WS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 74;
    }
    max_pairs = n;
  }
  p.n_tiles = (p.Ncols + BN - 1) / BN;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int rc = make_weight_tmap_bf16(&tm, w_bf16, p.Ncols, p.K, p.ldw, BN / 2);
  if (rc) return rc;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_work = m_pairs * p.n_tiles * p.splits;
  const int pairs = total_work < max_pairs ? total_work : max_pairs;
  kern<<<2 * pairs, WS_THREADS, Cfg::SMEM_BYTES, st>>>(p, tm, total_work, m_pairs);
  return static_cast<int>(cudaGetLastError());
}

// b_lo_off > 0: the B operand is pre-split ([hi | lo], lo at w + b_lo_off) and is fetched by TMA
template <int MODE>
static int dispatch(IgemmParams& p, int prec, long long b_lo_off, cudaStream_t st) {
  const int bn = p.Ncols > 64 ? 128 : 64;
  p.n_tiles = (p.Ncols + bn - 1) / bn;
  CUtensorMap tm_hi, tm_lo;
  memset(&tm_hi, 0, sizeof(tm_hi));
  memset(&tm_lo, 0, sizeof(tm_lo));
  if constexpr (MODE == MODE_ROW) {
    if (b_lo_off > 0) {
      int rc = make_weight_tmap(&tm_hi, p.w, p.Ncols, p.K, p.ldw, bn);
      if (rc) return rc;
      rc = make_weight_tmap(&tm_lo, p.w + b_lo_off, p.Ncols, p.K, p.ldw, bn);
      if (rc) return rc;
      // schedule selection: the persistent warp-specialised kernel wins at PREC=1 (+20 %); at PREC=2
      // it depends on the shape (below).  CAVP_IGEMM_WS=1/0 forces it on/off.
      static const char* ws_env = getenv("CAVP_IGEMM_WS");
      // Schedule choice, measured per shape at PREC=2 (profiles/r01_shape_sweep_sched.txt):
      //  * the persistent kernel (ws) beats the one-tile-per-CTA kernel by 5-30 % (its epilogue overlaps the next tile's
      //    main loop; short-K GEMMs gain most) except when the epilogue fetches a residual or evaluates erf/exp;
      //  * the CTA-pair kernel (ws2) adds 5-15 % on top when there are enough 256-row pair tiles to give every one of the
      //    74 TPCs an item (each SM reads only half of the weight tile from shared memory; round 2: the VGG 512-channel
      //    convs, M = 6144, went from 228 to 303 TFLOP/s when the threshold dropped from two waves to one); with 160-column tiles where they pad N
      //    less than 128-column tiles (N = 304: +35 %) or barely more on long-K GEMMs (N = 2048, K >= 1024: +13 %).
      // CAVP_IGEMM_WS=0/1/2 forces tile / ws / pair.
      const int m_tiles_ = (p.M + BM - 1) / BM;
      const bool heavy_epilogue = (p.res != nullptr && !igemm_inplace_acc(p)) || p.act == ACT_GELU || p.act == ACT_SIGMOID;
      const int pair_items = ((m_tiles_ + 1) / 2) * ((p.Ncols + 127) / 128) * p.splits;
      int sched = prec == 1 ? 1 : (heavy_epilogue ? 0 : ((bn == 128 && m_tiles_ >= 2 && pair_items >= 74) ? 2 : 1));
      if (ws_env) sched = ws_env[0] - '0';
      {
        // short-K linear layers (the fusion block): the schedule with eight promotion / epilogue warps
        static const char* lin_env = getenv("CAVP_IGEMM_LIN");
        const bool linear = p.R == 1 && p.S == 1 && p.stride == 1 && p.pad == 0 && p.Hs == p.Ho && p.Ws == p.Wo;
        // layers the 256-column kernel could take as well (N % 256 == 0) go there unless K is short (<= LIN_X_KB k-blocks)
        static const char* linx_env = getenv("CAVP_IGEMM_LIN_X_KB");
        const int lin_x_kb = linx_env ? atoi(linx_env) : 0;
        const bool x_ok = ws2x_epilogue_ok(p);
        const bool lin_ok = sched == 2 && bn == 128 && prec == 2 && linear && (p.num_kb & 1) == 0 && p.num_kb <= 40 &&
                            lin_epilogue_ok(p) && !(lin_env && lin_env[0] == '0') && (!x_ok || p.num_kb <= lin_x_kb);
        if (lin_ok) {
          const int pad128 = (p.Ncols + 127) / 128 * 128, pad160 = (p.Ncols + 159) / 160 * 160;
          const bool bn160 = pad160 <= pad128;
          const int pbn = bn160 ? 160 : 128;
          p.n_tiles = (p.Ncols + pbn - 1) / pbn;
          rc = make_weight_tmap(&tm_hi, p.w, p.Ncols, p.K, p.ldw, pbn / 2);
          if (rc) return rc;
          rc = make_weight_tmap(&tm_lo, p.w + b_lo_off, p.Ncols, p.K, p.ldw, pbn / 2);
          if (rc) return rc;
          return bn160 ? launch_igemm_lin<160>(p, tm_hi, tm_lo, st) : launch_igemm_lin<128>(p, tm_hi, tm_lo, st);
        }
      }
      if (sched == 2 && bn == 128 && prec == 2 && ws2x_epilogue_ok(p)) {
        // 256-column pair tiles (igemm_ws2x.cuh) when the halved number of work items still fills the 74 TPCs evenly
        static const char* x_env = getenv("CAVP_IGEMM_BN256");
        const int items = ((m_tiles_ + 1) / 2) * (p.Ncols / 256) * p.splits;
        const int waves = (items + 73) / 74;
        const bool fill_ok = items >= 74 * 6 || items * 100 >= waves * 74 * 85;
        if (x_env ? x_env[0] != '0' : fill_ok) {
          p.n_tiles = p.Ncols / 256;
          rc = make_weight_tmap(&tm_hi, p.w, p.Ncols, p.K, p.ldw, 128);
          if (rc) return rc;
          rc = make_weight_tmap(&tm_lo, p.w + b_lo_off, p.Ncols, p.K, p.ldw, 128);
          if (rc) return rc;
          return launch_igemm_ws2x(p, tm_hi, tm_lo, st);
        }
      }
      if (sched == 2 && bn == 128) {  // CTA-pair kernel: each CTA fetches half of the weight rows
        // 160-column tiles when they pad N less than 128-column tiles do (N = 304: 2 x 160 = 320 instead of 3 x 128 = 384)
        static const char* bn_env = getenv("CAVP_IGEMM_BN160");
        const int pad128 = (p.Ncols + 127) / 128 * 128, pad160 = (p.Ncols + 159) / 160 * 160;
        const bool bn160 = prec == 2 && (bn_env ? bn_env[0] != '0'
                                                : (pad160 < pad128 || (pad160 * 100 <= pad128 * 102 && p.num_kb >= 32)));
        const int pbn = bn160 ? 160 : 128;
        p.n_tiles = (p.Ncols + pbn - 1) / pbn;
        rc = make_weight_tmap(&tm_hi, p.w, p.Ncols, p.K, p.ldw, pbn / 2);
        if (rc) return rc;
        rc = make_weight_tmap(&tm_lo, p.w + b_lo_off, p.Ncols, p.K, p.ldw, pbn / 2);
        if (rc) return rc;
        if (bn160) return launch_igemm_ws2<160, 2>(p, tm_hi, tm_lo, st);
        return prec == 2 ? launch_igemm_ws2<128, 2>(p, tm_hi, tm_lo, st) : launch_igemm_ws2<128, 1>(p, tm_hi, tm_lo, st);
      }
      const bool use_ws = sched >= 1;
      if (use_ws) {
        if (prec == 2)
          return bn == 128 ? launch_igemm_ws<128, 2>(p, tm_hi, tm_lo, st) : launch_igemm_ws<64, 2>(p, tm_hi, tm_lo, st);
        return bn == 128 ? launch_igemm_ws<128, 1>(p, tm_hi, tm_lo, st) : launch_igemm_ws<64, 1>(p, tm_hi, tm_lo, st);
      }
      if (prec == 2)
        return bn == 128 ? launch_igemm<128, 2, MODE, true>(p, tm_hi, tm_lo, st)
                         : launch_igemm<64, 2, MODE, true>(p, tm_hi, tm_lo, st);
      return bn == 128 ? launch_igemm<128, 1, MODE, true>(p, tm_hi, tm_lo, st)
                       : launch_igemm<64, 1, MODE, true>(p, tm_hi, tm_lo, st);
    }
  }
  if constexpr (MODE == MODE_WGRAD) {
    if (b_lo_off > 0) {  // dY pre-split and dense: fetched by TMA
      int rc = make_dy_tmap(&tm_hi, p.w, p.red_len, p.M);
      if (rc) return rc;
      rc = make_dy_tmap(&tm_lo, p.w + b_lo_off, p.red_len, p.M);
      if (rc) return rc;
      // CTA-pair kernel (256 output channels per pair, each CTA gathers half of the im2col tile) when there are at
      // least two channel tiles; CAVP_WGRAD_PAIR=0/1 forces it off/on
      static const char* pair_env = getenv("CAVP_WGRAD_PAIR");
      const bool pair = bn == 128 && p.M > BM && (pair_env ? pair_env[0] != '0' : true);
      if (pair) return prec == 2 ? launch_wgrad2<2>(p, tm_hi, tm_lo, st) : launch_wgrad2<1>(p, tm_hi, tm_lo, st);
      if (prec == 2)
        return bn == 128 ? launch_igemm<128, 2, MODE, true>(p, tm_hi, tm_lo, st)
                         : launch_igemm<64, 2, MODE, true>(p, tm_hi, tm_lo, st);
      return bn == 128 ? launch_igemm<128, 1, MODE, true>(p, tm_hi, tm_lo, st)
                       : launch_igemm<64, 1, MODE, true>(p, tm_hi, tm_lo, st);
    }
  }
  if (prec == 2)
    return bn == 128 ? launch_igemm<128, 2, MODE, false>(p, tm_hi, tm_lo, st)
                     : launch_igemm<64, 2, MODE, false>(p, tm_hi, tm_lo, st);
  return bn == 128 ? launch_igemm<128, 1, MODE, false>(p, tm_hi, tm_lo, st)
                   : launch_igemm<64, 1, MODE, false>(p, tm_hi, tm_lo, st);
}

// hi = rn_tf32(w), lo = rn_tf32(w - hi): done once per step per weight instead of once per CTA per k-block
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                  long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(w)[i];
    const float h0 = tf32_rn(v.x), h1 = tf32_rn(v.y), h2 = tf32_rn(v.z), h3 = tf32_rn(v.w);
    reinterpret_cast<float4*>(hi)[i] = make_float4(h0, h1, h2, h3);
    reinterpret_cast<float4*>(lo)[i] =
        make_float4(tf32_rn(v.x - h0), tf32_rn(v.y - h1), tf32_rn(v.z - h2), tf32_rn(v.w - h3));
  }
}

// strided source [rows][ld] (a channel window of a gradient buffer) -> dense hi / lo [rows][cols]
__global__ void split_tf32_2d_kernel(const float* __restrict__ src, int ld, long long rows, int cols4,
                                     float* __restrict__ hi, float* __restrict__ lo) {
  const long long n4 = rows * cols4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols4;
    const int c = static_cast<int>(i - r * cols4);
    const float4 v = *reinterpret_cast<const float4*>(src + r * ld + c * 4);
    const float h0 = tf32_rn(v.x), h1 = tf32_rn(v.y), h2 = tf32_rn(v.z), h3 = tf32_rn(v.w);
    reinterpret_cast<float4*>(hi)[i] = make_float4(h0, h1, h2, h3);
    reinterpret_cast<float4*>(lo)[i] =
        make_float4(tf32_rn(v.x - h0), tf32_rn(v.y - h1), tf32_rn(v.z - h2), tf32_rn(v.w - h3));
  }
}

// strided fp32 [rows][ld] -> dense bf16 [rows][cols] (the dY operand of the bf16 weight-gradient kernel)
__global__ void cvt_bf16_2d_kernel(const float* __restrict__ src, int ld, long long rows, int cols4,
                                   __nv_bfloat16* __restrict__ dst) {
  const long long n4 = rows * cols4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols4;
    const int c = static_cast<int>(i - r * cols4);
    const float4 v = *reinterpret_cast<const float4*>(src + r * ld + c * 4);
    uint2 o;
    o.x = cvt_bf16x2(v.x, v.y);
    o.y = cvt_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
}

// dense bf16 [P][cout] matrix, box = 64 channels x 64 pixels, 128-byte swizzle = the MN-major bf16 UMMA operand layout
static int make_dy_tmap_bf16(CUtensorMap* tm, const void* dy, long long P, int cout) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CAVP_ERR_ARG;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cout), static_cast<cuuint64_t>(P)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cout) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(WgBf16Cfg::KPIX)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1000 + static_cast<int>(r);
}

static void fill_divs(IgemmParams& p) {
  p.div_howo = make_fastdiv(static_cast<uint32_t>(p.Ho * p.Wo));
  p.div_wo = make_fastdiv(static_cast<uint32_t>(p.Wo));
  p.div_c = make_fastdiv(static_cast<uint32_t>(p.C));
  p.div_s = make_fastdiv(static_cast<uint32_t>(p.S));
}

}  // namespace cavp

using namespace cavp;

static int igemm_row_impl(const float* x, const float* w, const void* w_bf16, float* y, float* y_pre,
                          const float* scale, const float* shift, const float* res, float* stats, int nimg, int hs,
                          int ws, int c, int ldx, int ho, int wo, int r, int s, int stride, int pad, int dil, int dgrad,
                          int ncols, int ldw, int ldy, int ldr, int res_mod, int res_div, int ldstat, int act,
                          float slope, int splits, int prec, long long b_lo_off, void* stream);

extern "C" int cavp_igemm(const float* x, const float* w, float* y, float* y_pre, const float* scale,
                          const float* shift, const float* res, float* stats, int nimg, int hs, int ws, int c, int ldx,
                          int ho, int wo, int r, int s, int stride, int pad, int dil, int dgrad, int ncols, int ldw,
                          int ldy, int ldr, int res_mod, int res_div, int ldstat, int act, float slope, int splits,
                          int prec, long long b_lo_off, void* stream) {
  if (prec != 1 && prec != 2) return CAVP_ERR_ARG;
  return igemm_row_impl(x, w, nullptr, y, y_pre, scale, shift, res, stats, nimg, hs, ws, c, ldx, ho, wo, r, s, stride,
                        pad, dil, dgrad, ncols, ldw, ldy, ldr, res_mod, res_div, ldstat, act, slope, splits, prec,
                        b_lo_off, stream);
}

// bf16-operand GEMM (BASELINE.json configs[2]): `w_bf16` is the bf16 copy of the K-major weight operand ([ncols][ldw]);
// `w` / `b_lo_off` are the fp32 operand used when the shape is outside the bf16 kernel's envelope (plain TF32 then)
extern "C" int cavp_igemm_bf16(const float* x, const float* w, const void* w_bf16, float* y, float* y_pre,
                               const float* scale, const float* shift, const float* res, float* stats, int nimg, int hs,
                               int ws, int c, int ldx, int ho, int wo, int r, int s, int stride, int pad, int dil,
                               int dgrad, int ncols, int ldw, int ldy, int ldr, int res_mod, int res_div, int ldstat,
                               int act, float slope, int splits, long long b_lo_off, void* stream) {
  return igemm_row_impl(x, w, w_bf16, y, y_pre, scale, shift, res, stats, nimg, hs, ws, c, ldx, ho, wo, r, s, stride,
                        pad, dil, dgrad, ncols, ldw, ldy, ldr, res_mod, res_div, ldstat, act, slope, splits, 1, b_lo_off,
                        stream);
}

static int igemm_row_impl(const float* x, const float* w, const void* w_bf16, float* y, float* y_pre,
                          const float* scale, const float* shift, const float* res, float* stats, int nimg, int hs,
                          int ws, int c, int ldx, int ho, int wo, int r, int s, int stride, int pad, int dil, int dgrad,
                          int ncols, int ldw, int ldy, int ldr, int res_mod, int res_div, int ldstat, int act,
                          float slope, int splits, int prec, long long b_lo_off, void* stream) {
  if (!x || !w || !y) return CAVP_ERR_NULL;
  if ((c & 3) || (ldx & 3) || (ldw & 3) || c <= 0 || ncols <= 0) return CAVP_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return CAVP_ERR_ALIGN;
  if (stride < 1 || dil < 1 || r < 1 || s < 1) return CAVP_ERR_ARG;
  const long long M = static_cast<long long>(nimg) * ho * wo;
  if (M <= 0 || M >= (1ll << 31) || static_cast<long long>(nimg) * hs * ws * ldx >= (1ll << 31)) return CAVP_ERR_ARG;
  IgemmParams p{};
  p.x = x; p.w = w; p.y = y; p.y_pre = y_pre; p.scale = scale; p.shift = shift; p.res = res; p.stats = stats;
  p.Nimg = nimg; p.Hs = hs; p.Ws = ws; p.C = c; p.ldx = ldx; p.Ho = ho; p.Wo = wo;
  p.R = r; p.S = s; p.stride = stride; p.pad = pad; p.dil = dil; p.dgrad = dgrad;
  p.M = static_cast<int>(M); p.Ncols = ncols; p.K = r * s * c; p.ldw = ldw; p.ldy = ldy; p.ldr = ldr;
  p.res_mod = res_mod; p.res_div = res_div; p.ldstat = ldstat; p.act = act; p.slope = slope;
  p.red_len = p.K;
  // bf16 kernel envelope: 8-element chunks never straddle a filter tap, TMA strides are multiples of 16 bytes
  const bool use_bf16 = w_bf16 != nullptr && (c & 7) == 0 && (ldw & 7) == 0 &&
                        (reinterpret_cast<uintptr_t>(w_bf16) & 15) == 0;
  const int bk = use_bf16 ? BK16 : BK;
  p.num_kb = (p.K + bk - 1) / bk;
  // splits < 0: deterministic split-K - y holds |splits| slabs of M*ldy floats, split i stores its raw partial product
  // in slab i (no atomics, no pre-zeroing); the caller sums the slabs in a fixed order
  const bool slabs = splits < -1;
  if (slabs) splits = -splits;
  p.splits = splits < 1 ? 1 : (splits > p.num_kb ? p.num_kb : splits);
  p.split_slab = (slabs && p.splits > 1) ? static_cast<long long>(p.M) * ldy : 0;
  if (slabs && p.splits != splits) return CAVP_ERR_ARG;  // the caller sized y for exactly |splits| slabs
  if (p.splits > 1 && (y_pre || scale || shift || res || stats || act != ACT_NONE)) return CAVP_ERR_ARG;
  fill_divs(p);
  // Promotion unit of the CTA-pair kernel: 2 k-blocks (64 K-elements, error ~5e-7 for every K).  Short-K linear layers
  // (the fusion block, K = 304: 10 k-blocks per tile) are bound by their promotion / epilogue warps (ncu: ~90 % busy,
  // tensor pipe 49 %): they take half of their K extent per unit (<= 6 k-blocks = 192 K-elements, error ~1.5e-6),
  // which cuts the tcgen05.ld + add work per tile 2.5x.  CAVP_IGEMM_UNIT=2 restores the 64-element units.
  p.unit_kb = 2;
  {
    static const char* unit_env = getenv("CAVP_IGEMM_UNIT");
    const bool linear = r == 1 && s == 1 && stride == 1 && pad == 0 && hs == ho && ws == wo;
    if (linear && p.num_kb > 4 && p.num_kb <= 12 && p.splits == 1 && !(unit_env && unit_env[0] == '2'))
      p.unit_kb = (p.num_kb + 1) / 2;
  }
  if (b_lo_off > 0 && ((b_lo_off & 3) || (reinterpret_cast<uintptr_t>(w) & 15))) return CAVP_ERR_ALIGN;
  if (use_bf16) {
    // 256-column pair tiles where they pad N no more than 128-column tiles do (N = 256, 512, 1024, 2048 ...)
    const int pad128 = (ncols + 127) / 128 * 128, pad256 = (ncols + 255) / 256 * 256;
    static const char* bn_env = getenv("CAVP_BF16_BN");
    const bool bn256 = bn_env ? bn_env[0] == '2' : pad256 == pad128;
    return bn256 ? launch_igemm_bf16<256>(p, w_bf16, static_cast<cudaStream_t>(stream))
                 : launch_igemm_bf16<128>(p, w_bf16, static_cast<cudaStream_t>(stream));
  }
  return dispatch<MODE_ROW>(p, prec, b_lo_off, static_cast<cudaStream_t>(stream));
}

static int wgrad_impl(const float* dy, const float* x, float* dw, int nimg, int hs, int ws, int c, int ldx, int ho,
                      int wo, int r, int s, int stride, int pad, int dil, int cout, int lddy, int splits, int prec,
                      long long dy_lo_off, void* stream);

extern "C" int cavp_igemm_wgrad(const float* dy, const float* x, float* dw, int nimg, int hs, int ws, int c, int ldx,
                                int ho, int wo, int r, int s, int stride, int pad, int dil, int cout, int lddy,
                                int splits, int prec, void* stream) {
  return wgrad_impl(dy, x, dw, nimg, hs, ws, c, ldx, ho, wo, r, s, stride, pad, dil, cout, lddy, splits, prec, 0, stream);
}

extern "C" int cavp_igemm_wgrad_tma(const float* dy_hi, long long dy_lo_off, const float* x, float* dw, int nimg, int hs,
                                    int ws, int c, int ldx, int ho, int wo, int r, int s, int stride, int pad, int dil,
                                    int cout, int splits, int prec, void* stream) {
  if (dy_lo_off <= 0 || (dy_lo_off & 3)) return CAVP_ERR_ARG;
  return wgrad_impl(dy_hi, x, dw, nimg, hs, ws, c, ldx, ho, wo, r, s, stride, pad, dil, cout, cout, splits, prec,
                    dy_lo_off, stream);
}

extern "C" int cavp_split_tf32_2d(const float* src, int ld, long long rows, int cols, float* hi, float* lo,
                                  void* stream) {
  if (!src || !hi || !lo) return CAVP_ERR_NULL;
  if ((cols & 3) || (ld & 3) || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(hi) & 15) ||
      (reinterpret_cast<uintptr_t>(lo) & 15))
    return CAVP_ERR_ALIGN;
  long long blocks = (rows * (cols / 4) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  split_tf32_2d_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, ld, rows,
                                                                                                    cols / 4, hi, lo);
  return static_cast<int>(cudaGetLastError());
}

static int wgrad_impl(const float* dy, const float* x, float* dw, int nimg, int hs, int ws, int c, int ldx, int ho,
                      int wo, int r, int s, int stride, int pad, int dil, int cout, int lddy, int splits, int prec,
                      long long dy_lo_off, void* stream) {
  if (!dy || !x || !dw) return CAVP_ERR_NULL;
  if ((c & 3) || (ldx & 3) || (cout & 3) || (lddy & 3)) return CAVP_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15)) return CAVP_ERR_ALIGN;
  if (prec != 1 && prec != 2) return CAVP_ERR_ARG;
  const long long P = static_cast<long long>(nimg) * ho * wo;
  if (P <= 0 || P >= (1ll << 31) || static_cast<long long>(nimg) * hs * ws >= (1ll << 31)) return CAVP_ERR_ARG;
  IgemmParams p{};
  p.x = x; p.w = dy; p.y = dw;
  p.Nimg = nimg; p.Hs = hs; p.Ws = ws; p.C = c; p.ldx = ldx; p.Ho = ho; p.Wo = wo;
  p.R = r; p.S = s; p.stride = stride; p.pad = pad; p.dil = dil;
  p.M = cout; p.Ncols = r * s * c; p.K = r * s * c; p.ldw = lddy; p.ldy = r * s * c;
  p.red_len = static_cast<int>(P);
  p.num_kb = (p.red_len + BK - 1) / BK;
  p.splits = splits < 1 ? 1 : (splits > p.num_kb ? p.num_kb : splits);
  p.unit_kb = 2;
  fill_divs(p);
  return dispatch<MODE_WGRAD>(p, prec, dy_lo_off, static_cast<cudaStream_t>(stream));
}

extern "C" int cavp_split_tf32(const float* w, float* hi, float* lo, long long n, void* stream) {
  if ((n & 3) || (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(hi) & 15) ||
      (reinterpret_cast<uintptr_t>(lo) & 15))
    return CAVP_ERR_ALIGN;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  split_tf32_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, hi, lo, n / 4);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_cvt_bf16_2d(const float* src, int ld, long long rows, int cols, void* dst, void* stream) {
  if (!src || !dst) return CAVP_ERR_NULL;
  if ((cols & 3) || (ld & 3) || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 7))
    return CAVP_ERR_ALIGN;
  long long blocks = (rows * (cols / 4) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  cvt_bf16_2d_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, ld, rows, cols / 4, static_cast<__nv_bfloat16*>(dst));
  return static_cast<int>(cudaGetLastError());
}

// bf16 weight gradient (cavp_prec = 3): dy_bf16 = dense bf16 [P][cout]; x = fp32 NHWC source of the im2col operand
extern "C" int cavp_igemm_wgrad_bf16(const void* dy_bf16, const float* x, float* dw, int nimg, int hs, int ws, int c,
                                     int ldx, int ho, int wo, int r, int s, int stride, int pad, int dil, int cout,
                                     int splits, void* stream) {
  if (!dy_bf16 || !x || !dw) return CAVP_ERR_NULL;
  if ((c & 7) || (ldx & 3) || (cout & 7)) return CAVP_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy_bf16) & 15)) return CAVP_ERR_ALIGN;
  if (cout <= BM) return CAVP_ERR_ARG;  // the pair tile is 256 output channels; small layers use the TF32 kernels
  const long long P = static_cast<long long>(nimg) * ho * wo;
  if (P <= 0 || P >= (1ll << 31) || static_cast<long long>(nimg) * hs * ws >= (1ll << 31)) return CAVP_ERR_ARG;
  IgemmParams p{};
  p.x = x; p.w = nullptr; p.y = dw;
  p.Nimg = nimg; p.Hs = hs; p.Ws = ws; p.C = c; p.ldx = ldx; p.Ho = ho; p.Wo = wo;
  p.R = r; p.S = s; p.stride = stride; p.pad = pad; p.dil = dil;
  p.M = cout; p.Ncols = r * s * c; p.K = r * s * c; p.ldw = cout; p.ldy = r * s * c;
  p.red_len = static_cast<int>(P);
  p.num_kb = (p.red_len + WgBf16Cfg::KPIX - 1) / WgBf16Cfg::KPIX;
  p.splits = splits < 1 ? 1 : (splits > p.num_kb ? p.num_kb : splits);
  p.n_tiles = (p.Ncols + WgBf16Cfg::BN - 1) / WgBf16Cfg::BN;
  fill_divs(p);
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int rc = make_dy_tmap_bf16(&tm, dy_bf16, P, cout);
  if (rc) return rc;
  auto kern = igemm_wgrad_bf16_kernel;
  static bool configured_dev[MAX_DEVICES] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WgBf16Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  const int m_pairs = (m_tiles + 1) / 2;
  dim3 grid(static_cast<unsigned>(2 * m_pairs * p.n_tiles), static_cast<unsigned>(p.splits), 1);
  kern<<<grid, CTA_THREADS, WgBf16Cfg::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(p, tm);
  return static_cast<int>(cudaGetLastError());
}
