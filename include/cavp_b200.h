#pragma once
#ifdef __cplusplus
extern "C" {
#endif
#define CAVP_OK 0
#define CAVP_ERR_NULL (-1)
#define CAVP_ERR_ALIGN (-2)
#define CAVP_ERR_ARG (-3)
int cavp_igemm(const float* x, const float* w, float* y, float* y_pre, const float* scale, const float* shift,
               const float* res, float* stats, int nimg, int hs, int ws, int c, int ldx, int ho, int wo, int r, int s,
               int stride, int pad, int dil, int dgrad, int ncols, int ldw, int ldy, int ldr, int res_mod, int ldstat,
               int act, float slope, int splits, int prec, void* stream);
int cavp_igemm_wgrad(const float* dy, const float* x, float* dw, int nimg, int hs, int ws, int c, int ldx, int ho,
                     int wo, int r, int s, int stride, int pad, int dil, int cout, int lddy, int splits, int prec,
                     void* stream);
#ifdef __cplusplus
}
#endif
