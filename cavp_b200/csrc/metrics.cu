// Eval epilogue: per-pixel argmax of the segmentation logits + the (label, prediction) confusion
// counts every reference metric is derived from, in one pass.
//
// Reference (utils/eval_utils.py): MIoU.batch_pix_accuracy / batch_intersection_union (:73-97: torch.max over classes,
// three torch.histc calls) and ForegroundDetect.__call__ / _fast_hist (:107-117,151-155: argmax -> .cpu().numpy() ->
// numpy.bincount per image).  All of it is a function of  conf[t][p] = #pixels with label t predicted as p :
//   pixel_labeled = sum(conf), pixel_correct = trace, area_inter = diag, area_pred = column sums, area_lab = row sums,
//   ForegroundDetect's hist = conf itself.
// Labels: t in [0, C) -> row t; t == ignore_index or t < 0 -> not counted; any other value -> row C ("labelled but out
// of range": MIoU counts such pixels as labelled and their predictions in area_pred, ForegroundDetect drops them).
// HBM-bound: the materialised path reads B*C*HW*4 bytes of logits once; the fused path reads only the low-resolution
// NHWC logits (the 16 full-resolution pixels of a 4x4 patch share their four source pixels through L1).
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/cavp_b200.h"
#include "common.cuh"

namespace cavp {

constexpr int MET_THREADS = 256;

__device__ __forceinline__ void count_pixel(long long t, int pred, int C, int ignore_index, int* sh_conf,
                                            unsigned long long* conf, bool use_smem) {
  if (t == ignore_index || t < 0) return;
  const int row = t < C ? static_cast<int>(t) : C;
  if (use_smem)
    atomicAdd(&sh_conf[row * C + pred], 1);
  else
    atomicAdd(&conf[row * C + pred], 1ull);
}

__device__ __forceinline__ void flush_counts(const int* sh_conf, unsigned long long* conf, int n) {
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int v = sh_conf[i];
    if (v) atomicAdd(&conf[i], static_cast<unsigned long long>(v));
  }
}

// logits NCHW [B][C][HW]; one thread per pixel, channel loop with coalesced reads across the warp
__global__ void __launch_bounds__(MET_THREADS)
argmax_confusion_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C,
                        long long HW, int ignore_index, long long* __restrict__ pred_out,
                        unsigned long long* __restrict__ conf, int use_smem) {
  extern __shared__ int sh_conf[];
  const int ncell = (C + 1) * C;
  if (use_smem) {
    for (int i = threadIdx.x; i < ncell; i += blockDim.x) sh_conf[i] = 0;
    __syncthreads();
  }
  const long long total = static_cast<long long>(B) * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / HW, px = i - b * HW;
    const float* lp = logits + b * C * HW + px;
    float best = lp[0];
    int arg = 0;
    for (int ch = 1; ch < C; ++ch) {
      const float v = lp[static_cast<long long>(ch) * HW];
      if (v > best) {  // strict: the first maximum wins, as torch.max / torch.argmax on equal values
        best = v;
        arg = ch;
      }
    }
    if (pred_out) pred_out[i] = arg;
    if (labels && conf) count_pixel(labels[i], arg, C, ignore_index, sh_conf, conf, use_smem != 0);
  }
  if (use_smem && conf) flush_counts(sh_conf, conf, ncell);
}

// x: low-resolution NHWC logits [n][hin][win][ldx]; output pixel (oy, ox) takes the bilinear sample (align_corners =
// False, cavp_model.py:140) of every class with the SAME arithmetic as cavp_bilinear_fwd, then the argmax.
__global__ void __launch_bounds__(MET_THREADS)
upsample_argmax_confusion_kernel(const float* __restrict__ x, int ldx, int hin, int win, int hout, int wout, int n,
                                 int C, float sh, float sw, const long long* __restrict__ labels, int ignore_index,
                                 long long* __restrict__ pred_out, unsigned long long* __restrict__ conf,
                                 int use_smem) {
  extern __shared__ int sh_conf[];
  const int ncell = (C + 1) * C;
  if (use_smem) {
    for (int i = threadIdx.x; i < ncell; i += blockDim.x) sh_conf[i] = 0;
    __syncthreads();
  }
  const long long plane = static_cast<long long>(hout) * wout;
  const long long total = static_cast<long long>(n) * plane;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % wout);
    const int oy = static_cast<int>((i / wout) % hout);
    const long long img = i / plane;
    const Lerp ly = lerp_index(oy, hin, sh, 0), lx = lerp_index(ox, win, sw, 0);
    const float* base = x + img * hin * win * ldx;
    const float* pa = base + (static_cast<long long>(ly.i0) * win + lx.i0) * ldx;
    const float* pb = base + (static_cast<long long>(ly.i0) * win + lx.i1) * ldx;
    const float* pc = base + (static_cast<long long>(ly.i1) * win + lx.i0) * ldx;
    const float* pd = base + (static_cast<long long>(ly.i1) * win + lx.i1) * ldx;
    float best = bilerp(ly, lx, __ldg(pa), __ldg(pb), __ldg(pc), __ldg(pd));
    int arg = 0;
    for (int ch = 1; ch < C; ++ch) {
      const float v = bilerp(ly, lx, __ldg(pa + ch), __ldg(pb + ch), __ldg(pc + ch), __ldg(pd + ch));
      if (v > best) {
        best = v;
        arg = ch;
      }
    }
    if (pred_out) pred_out[i] = arg;
    if (labels && conf) count_pixel(labels[i], arg, C, ignore_index, sh_conf, conf, use_smem != 0);
  }
  if (use_smem && conf) flush_counts(sh_conf, conf, ncell);
}

static int smem_for(int C, int* use_smem) {
  const long long bytes = static_cast<long long>(C + 1) * C * 4;
  *use_smem = bytes <= 96 * 1024;
  return *use_smem ? static_cast<int>(bytes) : 0;
}

}  // namespace cavp

using namespace cavp;

extern "C" int cavp_argmax_confusion(const float* logits, const long long* labels, int B, int C, long long HW,
                                     int ignore_index, long long* pred, unsigned long long* conf, void* stream) {
  if (!logits) return CAVP_ERR_NULL;
  if (B <= 0 || C <= 0 || HW <= 0) return CAVP_ERR_ARG;
  int use_smem = 0;
  const int smem = smem_for(C, &use_smem);
  auto kern = argmax_confusion_kernel;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  kern<<<grid_for(static_cast<long long>(B) * HW, MET_THREADS, 4), MET_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      logits, labels, B, C, HW, ignore_index, pred, conf, use_smem);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int cavp_upsample_argmax_confusion(const float* x, int ldx, int hin, int win, int hout, int wout, int n,
                                              int C, const long long* labels, int ignore_index, long long* pred,
                                              unsigned long long* conf, void* stream) {
  if (!x) return CAVP_ERR_NULL;
  if (n <= 0 || C <= 0 || hin <= 0 || win <= 0 || hout <= 0 || wout <= 0 || ldx < C) return CAVP_ERR_ARG;
  int use_smem = 0;
  const int smem = smem_for(C, &use_smem);
  auto kern = upsample_argmax_confusion_kernel;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const float sh = lerp_scale(hin, hout, 0), sw = lerp_scale(win, wout, 0);
  kern<<<grid_for(static_cast<long long>(n) * hout * wout, MET_THREADS, 4), MET_THREADS, smem,
         static_cast<cudaStream_t>(stream)>>>(x, ldx, hin, win, hout, wout, n, C, sh, sw, labels, ignore_index, pred,
                                              conf, use_smem);
  return static_cast<int>(cudaGetLastError());
}
