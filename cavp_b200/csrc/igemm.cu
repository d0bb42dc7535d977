// C-ABI launchers for the tcgen05 implicit-GEMM kernels (igemm.cuh).  See include/cavp_b200.h for the contract.
#include "igemm.cuh"
#include "../../include/cavp_b200.h"

namespace cavp {

template <int BN, int PREC, int MODE>
static int launch_igemm(const IgemmParams& p, cudaStream_t st) {
  using Cfg = TileCfg<BN, PREC>;
  auto kern = igemm_kernel<BN, PREC, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM;
  dim3 grid(static_cast<unsigned>(m_tiles * p.n_tiles), static_cast<unsigned>(p.splits), 1);
  kern<<<grid, CTA_THREADS, Cfg::SMEM_BYTES, st>>>(p);
  return static_cast<int>(cudaGetLastError());
}

template <int MODE>
static int dispatch(IgemmParams& p, int prec, cudaStream_t st) {
  const int bn = p.Ncols > 64 ? 128 : 64;
  p.n_tiles = (p.Ncols + bn - 1) / bn;
  if (prec == 2) return bn == 128 ? launch_igemm<128, 2, MODE>(p, st) : launch_igemm<64, 2, MODE>(p, st);
  return bn == 128 ? launch_igemm<128, 1, MODE>(p, st) : launch_igemm<64, 1, MODE>(p, st);
}

static void fill_divs(IgemmParams& p) {
  p.div_howo = make_fastdiv(static_cast<uint32_t>(p.Ho * p.Wo));
  p.div_wo = make_fastdiv(static_cast<uint32_t>(p.Wo));
  p.div_c = make_fastdiv(static_cast<uint32_t>(p.C));
  p.div_s = make_fastdiv(static_cast<uint32_t>(p.S));
}

}  // namespace cavp

using namespace cavp;

extern "C" int cavp_igemm(const float* x, const float* w, float* y, float* y_pre, const float* scale,
                          const float* shift, const float* res, float* stats, int nimg, int hs, int ws, int c, int ldx,
                          int ho, int wo, int r, int s, int stride, int pad, int dil, int dgrad, int ncols, int ldw,
                          int ldy, int ldr, int res_mod, int res_div, int ldstat, int act, float slope, int splits,
                          int prec,
                          void* stream) {
  if (!x || !w || !y) return CAVP_ERR_NULL;
  if ((c & 3) || (ldx & 3) || (ldw & 3) || c <= 0 || ncols <= 0) return CAVP_ERR_ALIGN;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return CAVP_ERR_ALIGN;
  if (prec != 1 && prec != 2) return CAVP_ERR_ARG;
  if (stride < 1 || dil < 1 || r < 1 || s < 1) return CAVP_ERR_ARG;
  const long long M = static_cast<long long>(nimg) * ho * wo;
  if (M <= 0 || M >= (1ll << 31) || static_cast<long long>(nimg) * hs * ws >= (1ll << 31)) return CAVP_ERR_ARG;
  IgemmParams p{};
  p.x = x; p.w = w; p.y = y; p.y_pre = y_pre; p.scale = scale; p.shift = shift; p.res = res; p.stats = stats;
  p.Nimg = nimg; p.Hs = hs; p.Ws = ws; p.C = c; p.ldx = ldx; p.Ho = ho; p.Wo = wo;
  p.R = r; p.S = s; p.stride = stride; p.pad = pad; p.dil = dil; p.dgrad = dgrad;
  p.M = static_cast<int>(M); p.Ncols = ncols; p.K = r * s * c; p.ldw = ldw; p.ldy = ldy; p.ldr = ldr;
  p.res_mod = res_mod; p.res_div = res_div; p.ldstat = ldstat; p.act = act; p.slope = slope;
  p.red_len = p.K;
  p.num_kb = (p.K + BK - 1) / BK;
  p.splits = splits < 1 ? 1 : (splits > p.num_kb ? p.num_kb : splits);
  if (p.splits > 1 && (y_pre || scale || shift || res || stats || act != ACT_NONE)) return CAVP_ERR_ARG;
  fill_divs(p);
  return dispatch<MODE_ROW>(p, prec, static_cast<cudaStream_t>(stream));
}

extern "C" int cavp_igemm_wgrad(const float* dy, const float* x, float* dw, int nimg, int hs, int ws, int c, int ldx,
                                int ho, int wo, int r, int s, int stride, int pad, int dil, int cout, int lddy,
                                int splits, int prec, void* stream) {
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
  fill_divs(p);
  return dispatch<MODE_WGRAD>(p, prec, static_cast<cudaStream_t>(stream));
}
