/* cavp_b200 - C ABI of the B200 (sm_100a) kernels behind the CAVP hot path.
 *
 * The reference (cyh-0/CAVP) is pure Python/PyTorch: it has no FFI of its own.  Each entry point below replaces the
 * torch op(s) that the reference calls at the cited file:line (paths relative to the reference repo), and is what a
 * ctypes binding on the reference side would bind.  Conventions:
 *   - raw device pointers, explicit shapes / leading dimensions, `void* stream` = cudaStream_t, int status return
 *     (0 = ok, <0 = argument error below, >0 = cudaError_t); nothing throws across the ABI;
 *   - the caller owns every buffer (outputs, workspaces); kernels keep no global state;
 *   - activations are NHWC fp32 with a pixel stride `ld` (>= channels, multiple of 4), so a channel slice of a concat
 *     buffer is an ordinary operand; weights are [Cout][R][S][Cin] (torch channels_last storage of an OIHW tensor);
 *   - `prec`: 1 = TF32 products, 2 = 3xTF32 products + fp32 register promotion (fp32-parity mode).
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

#define CAVP_OK 0
#define CAVP_ERR_NULL (-1)
#define CAVP_ERR_ALIGN (-2)
#define CAVP_ERR_ARG (-3)

/* activation codes for `act` */
#define CAVP_ACT_NONE 0
#define CAVP_ACT_RELU 1
#define CAVP_ACT_LEAKY 2
#define CAVP_ACT_GELU 3
#define CAVP_ACT_SIGMOID 4

/* ---- tensor-core tiles (csrc/igemm.cu) ---------------------------------------------------------------------------
 * cavp_igemm: y[M][ldy] = epilogue( im2col(x)[M][K] * w[ncols][K]^T ),  M = nimg*ho*wo, K = r*s*c.
 *   Replaces F.conv2d / nn.Conv2d.forward (models/visual/backbones/resnet.py:75-98,107-121;
 *   models/visual/deeplabv3/encoder_decoder.py:62-75,84-88,137-156; models/audio/backbones/vgg.py:26-36) and
 *   F.linear / nn.Linear (models/attn.py:30-39,64-71,100-104; timm Mlp; vgg.py:11-16) with r=s=1.
 *   dgrad=1 runs the transposed-stride gather (x = dY, w = weights transposed to [Cin][R][S][Cout], rows = input pixels):
 *   the data gradient of the same convolution (autograd's convolution_backward, trainer_cavp_vpo_mono.py:191).
 *   epilogue: v = acc*scale[col] + shift[col] + res[(row/res_div)%res_mod][col]; y_pre = v; y = act(v);
 *   stats (optional) receives per-(row-tile, warp) column sums of y and y^2: [ceil(M/128)*4][2][ldstat] - the
 *   BatchNorm batch statistics (train mode) without a second pass over y.
 *   splits > 1 splits K over CTAs and accumulates raw products into a PRE-ZEROED y with red.global (no epilogue).
 *   splits < -1 is the deterministic form: y holds |splits| slabs of M*ldy floats, split i stores its raw partial
 *   product in slab i (nothing to pre-zero); the caller reduces the slabs in a fixed order (cavp_partials_sum). */
int cavp_igemm(const float* x, const float* w, float* y, float* y_pre, const float* scale, const float* shift,
               const float* res, float* stats, int nimg, int hs, int ws, int c, int ldx, int ho, int wo, int r, int s,
               int stride, int pad, int dil, int dgrad, int ncols, int ldw, int ldy, int ldr, int res_mod, int res_div,
               int ldstat, int act, float slope, int splits, int prec, long long b_lo_off, void* stream);
/* cavp_split_tf32: hi = rn_tf32(w), lo = rn_tf32(w - hi), once per step per weight matrix.  Passing w = hi and
 * b_lo_off = (lo - hi) > 0 to cavp_igemm makes the kernel fetch the weight operand with TMA (cp.async.bulk.tensor, 128B
 * swizzle) instead of through the producer warps; b_lo_off = 0 keeps the in-kernel split (operand = activations). */
int cavp_split_tf32(const float* w, float* hi, float* lo, long long n, void* stream);
/* the same split for every weight operand of a model in ONE launch: table = device array of 32-byte rows {const float*
 * src; float* hi; float* lo; long long n}, work = (row, chunk) int pairs with chunk = cavp_opt_chunk_elems() elements */
int cavp_split_tf32_multi(const void* table, const int* work, int nwork, void* stream);
/* cavp_igemm with bf16 operands (BASELINE.json configs[2] "bf16 training loop"; the reference has no bf16 path):
 * tcgen05.mma.kind::f16, bf16 A/B, fp32 accumulation in TMEM, fp32 activations / epilogue / statistics.  w_bf16 =
 * bf16 copy of the K-major weight operand [ncols][ldw]; the fp32 activations are converted by the producer warps.
 * Shapes outside the kernel's envelope (c % 8, ldw % 8) run the TF32 kernels on `w` / `b_lo_off` instead. */
int cavp_igemm_bf16(const float* x, const float* w, const void* w_bf16, float* y, float* y_pre, const float* scale,
                    const float* shift, const float* res, float* stats, int nimg, int hs, int ws, int c, int ldx, int ho,
                    int wo, int r, int s, int stride, int pad, int dil, int dgrad, int ncols, int ldw, int ldy, int ldr,
                    int res_mod, int res_div, int ldstat, int act, float slope, int splits, long long b_lo_off,
                    void* stream);
/* bf16 weight gradient of the bf16 row (cavp_prec = 3): dw[cout][r*s*c] (+)= dy^T * im2col(x) with bf16 operands
 * (tcgen05.mma.kind::f16, both MN-major) and fp32 accumulation.  dy_bf16 = dense bf16 [P][cout] (cavp_cvt_bf16_2d, or
 * written by cavp_bn_bwd_apply with split_mode 1); x = fp32 NHWC.  Needs c % 8 == 0, cout % 8 == 0, cout > 128;
 * splits > 1 accumulates into a PRE-ZEROED dw. */
int cavp_igemm_wgrad_bf16(const void* dy_bf16, const float* x, float* dw, int nimg, int hs, int ws, int c, int ldx, int ho,
                          int wo, int r, int s, int stride, int pad, int dil, int cout, int splits, void* stream);
int cavp_cvt_bf16_2d(const float* src, int ld, long long rows, int cols, void* dst, void* stream);
/* cavp_igemm_wgrad: dw[cout][r*s*c] (+)= dy[P][cout]^T * im2col(x)[P][r*s*c]   (weight gradient; P = nimg*ho*wo).
 *   splits > 1 accumulates into a PRE-ZEROED dw. */
int cavp_igemm_wgrad(const float* dy, const float* x, float* dw, int nimg, int hs, int ws, int c, int ldx, int ho,
                     int wo, int r, int s, int stride, int pad, int dil, int cout, int lddy, int splits, int prec,
                     void* stream);
/* Same product with the dY operand pre-split (cavp_split_tf32_2d: dense hi | lo [P][cout], lo = hi + dy_lo_off elements)
 * and fetched by TMA (128B swizzle with 32-byte atoms = the MN-major tf32 operand layout); the kernel's producer warps
 * then gather only im2col(x). */
int cavp_igemm_wgrad_tma(const float* dy_hi, long long dy_lo_off, const float* x, float* dw, int nimg, int hs, int ws,
                         int c, int ldx, int ho, int wo, int r, int s, int stride, int pad, int dil, int cout,
                         int splits, int prec, void* stream);
int cavp_split_tf32_2d(const float* src, int ld, long long rows, int cols, float* hi, float* lo, void* stream);

/* ---- layout / plumbing (csrc/elementwise.cu) -------------------------------------------------------------------- */
int cavp_zero(void* ptr, long long bytes, void* stream);
/* dst[r][0..c) = value for a [rows][ld] window (zeroing channel slices / gradient pads) */
int cavp_fill_strided(float* dst, int ld, long long rows, int c, float value, void* stream);
/* NCHW [n][c][hw] -> NHWC [n][hw][cpad] (channels >= c zero-filled); the model boundary (image, log-mel, OIHW stem
 * weights).  cavp_nhwc_to_nchw is the inverse for the first c channels. */
int cavp_nchw_to_nhwc(const float* src, float* dst, int n, int c, int hw, int cpad, void* stream);
int cavp_nhwc_to_nchw(const float* src, float* dst, int n, int c, int hw, int ld, void* stream);
/* dst[b][j][i] = src[b][i][j] (weight transposes for dgrad / the InfoNCE gradient GEMM) */
int cavp_transpose(const float* src, float* dst, int rows, int cols, long long src_ld, long long dst_ld, int batch,
                   long long src_bs, long long dst_bs, void* stream);
/* the same transpose writing the TF32 split of the result (hi, lo): the dgrad weight operand in one pass */
int cavp_transpose_split(const float* src, float* hi, float* lo, int rows, int cols, long long src_ld, long long dst_ld,
                         int batch, long long src_bs, long long dst_bs, void* stream);
/* cavp_transpose_split for every dgrad weight operand of a model in ONE launch: table = device array of 72-byte rows
 * {const float* src; float* hi; float* lo; int rows, cols; long long src_ld, dst_ld, src_bs, dst_bs; int tiles_c,
 * tiles_r}; work = (row, tile) int pairs, tile enumerating (batch, 32-row tile, 32-column tile). */
int cavp_transpose_split_multi(const void* table, const int* work, int nwork, int bf16, void* stream);
/* ^ bf16 != 0: `hi` points to a bf16 destination and receives the rounded transpose (lo unused) */
int cavp_add_inplace(float* dst, const float* src, long long n, float alpha, void* stream);
/* dst[i] = src[idx[i]] (fea_a[shuffle_idx], models/cavp_model.py:171) or, accumulate_scatter=1, dst[idx[i]] += src[i] */
int cavp_gather_rows(const float* src, const long long* idx, float* dst, int nrows, int c, int accumulate_scatter,
                     void* stream);

/* ---- BatchNorm (nn.BatchNorm2d, eps 1e-5, momentum 0.1: encoder_decoder.py:10-11, resnet.py:64-72) --------------
 * finalize: reduce the igemm `stats` partials -> mean, invstd, scale = gamma*invstd, shift = beta - mean*scale, and the
 * running-stat update (unbiased variance; num_batches_tracked += 1 in the same launch when the pointer is given).
 * sums_mode: 0 = local statistics; 1 = only export the fp64 sums [2][C] and the local count to sums_io [2*C + 1]
 * (SyncBatchNorm, main_vpo_mono.py:130: the caller all-reduces that buffer); 2 = take the sums AND the count from
 * sums_io (the global count never visits the host). */
int cavp_bn_finalize(const float* partials, int nparts, int ldstat, int C, double count, const float* gamma,
                     const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                     float* mean_out, float* invstd_out, float* scale_out, float* shift_out, double* sums_io,
                     int sums_mode, long long* num_batches_tracked, void* stream);
int cavp_bn_eval_coeffs(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C,
                        float* scale, float* shift, void* stream);
/* out = act(y*scale + shift (+ res))   (BN apply + ReLU / LeakyReLU(0.01) + Bottleneck residual, resnet.py:86-96) */
int cavp_bn_apply(const float* y, int ldy, const float* scale, const float* shift, const float* res, int ldr, float* out,
                  int ldo, long long rows, int C, int act, float slope, void* stream);
/* zscale / zshift (colreduce, bn_bwd_apply): with z == NULL and an activation, act'(.) is taken from the recomputed
 * activation input fma(y, zscale, zshift) - the BN-apply expression - instead of reading the output z back.
 * column partial sums of g = dz*act'(z) and g*xhat (xhat from y, mean, invstd; y may be NULL -> plain column sums =
 * bias gradient); optionally writes g.  partials: [nblk][2][ldp]. */
int cavp_colreduce(const float* dz, int lddz, const float* z, int ldz, const float* y, int ldy, const float* mean,
                   const float* invstd, long long rows, int C, int act, float slope, float* gout, int ldg,
                   float* partials, int ldp, int nblk, const float* zscale, const float* zshift, void* stream);
int cavp_partials_sum(const float* partials, int nparts, int ldp, int C, int nk, float* out, void* stream);
/* dy = gamma*invstd*(g - sum_g/count - xhat*sum_gxhat/count); dres = g (gradient of the residual input) */
int cavp_bn_bwd_apply(const float* dz, int lddz, const float* z, int ldz, const float* y, int ldy, const float* mean,
                      const float* invstd, const float* gamma, const float* sums, float inv_count, long long rows, int C,
                      int act, float slope, float* dy, int lddy, float* dres, int lddres, const float* zscale,
                      const float* zshift, const double* count_dev, float* dy_hi, float* dy_lo, int split_bf16,
                      void* stream);
/* ^ count_dev != NULL: 1/count is taken from device memory (the all-reduced count of SyncBatchNorm) instead of inv_count.
 *   dy_hi / dy_lo != NULL: also write the dense [rows][C] TF32 split of dy (the operand of cavp_igemm_wgrad_tma);
 *   split_bf16 != 0: dy_hi receives a dense bf16 [rows][C] copy instead (the operand of cavp_igemm_wgrad_bf16). */

/* ---- pooling (F.max_pool2d resnet.py:189 / vgg.py:30; ASPP global pooling encoder_decoder.py:158-164;
 *      AdaptiveMaxPool2d audio_network.py:24) ------------------------------------------------------------------- */
int cavp_maxpool_fwd(const float* x, int ldx, float* y, int ldy, int* idx, int n, int h, int w, int c, int k,
                     int stride, int pad, int ho, int wo, void* stream);
int cavp_maxpool_bwd(const float* dy, int lddy, const int* idx, float* dx, int lddx, long long opix, int c,
                     void* stream);
int cavp_pixel_sum(const float* x, int ldx, float* out, int n, int hw, int c, float scale, void* stream);
int cavp_pixel_bcast(const float* dout, float* dx, int lddx, int n, int hw, int c, float scale, int accumulate,
                     void* stream);
int cavp_pixel_max(const float* x, int ldx, float* out, int* arg, int n, int hw, int c, void* stream);
int cavp_pixel_max_bwd(const float* dout, const int* arg, float* dx, int lddx, int n, int hw, int c, void* stream);

/* ---- bilinear resampling (F.interpolate mode="bilinear": encoder_decoder.py:101 align_corners=True,
 *      cavp_model.py:140 align_corners=False).  nchw_out=1 writes the full-resolution prediction [n][c][hout][wout].
 *      The backward is a gather (deterministic); images >= n_valid are known to have zero gradient. */
int cavp_bilinear_fwd(const float* x, int ldx, int hin, int win, float* y, int ldy, int hout, int wout, int n, int c,
                      int align_corners, int nchw_out, void* stream);
int cavp_bilinear_bwd(const float* dy, int lddy, int hout, int wout, float* dx, int lddx, int hin, int win, int n,
                      int c, int align_corners, int nchw_in, int n_valid, void* stream);

/* ---- fusion (csrc/fusion.cu): nn.LayerNorm (attn.py:133,137,229), the cross-attention core (attn.py:73-106 with a
 *      single audio key token: sigmoid gate), GELU' (timm Mlp) ------------------------------------------------------ */
int cavp_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                       long long T, int C, float eps, void* stream);
int cavp_layernorm_bwd_nparts(long long T);
int cavp_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       const float* add, float* dx, float* partials, long long T, int C, void* stream);
/* q [Bq][N][C]; k, v [Bq*rep][C]; x [Bq*rep][N][C]; attn [Bq*rep][heads][N].  Row r uses q row r % Bq. */
int cavp_gate_fwd(const float* q, const float* k, const float* v, float* x, float* attn, int Bq, int rep, int N, int C,
                  int heads, void* stream);
/* dk, dv must be pre-zeroed [Bq*rep][C] */
int cavp_gate_bwd(const float* dx, const float* q, const float* k, const float* v, const float* attn, float* dq,
                  float* dk, float* dv, int Bq, int rep, int N, int C, int heads, void* stream);
/* y = gelu(pre) (exact erf form), in place allowed; dx = dy * gelu'(pre) */
int cavp_gelu_fwd(const float* pre, float* y, long long n, void* stream);
int cavp_gelu_bwd(const float* dy, const float* pre, float* dx, long long n, void* stream);

/* ---- losses (csrc/loss.cu): CrossEntropyLoss(ignore_index) loss/losser.py:60-62; ContrastLoss
 *      loss/contrastive_aud.py:17-74 -------------------------------------------------------------------------------- */
int cavp_ce_nblocks(int B, long long HW);
/* logits NCHW [>=B][C][HW]; labels int64 [B][HW]; loss_and_count = {mean loss, #valid pixels} */
int cavp_ce_fwd(const float* logits, const long long* labels, int B, int C, long long HW, int ignore_index,
                float* partials, float* loss_and_count, void* stream);
int cavp_ce_bwd(const float* logits, const long long* labels, int B, int C, long long HW, int ignore_index,
                const float* loss_and_count, const float* gscale, float* dlogits, void* stream);
/* Fused forward_cls upsample + CrossEntropyLoss for training (models/cavp_model.py:138-141 -> trainer_cavp_vpo_mono.py
 * :171,187 -> loss/losser.py:60-62): x = low-resolution NHWC logits [n][hin][win][ldx]; the bilinear (align_corners =
 * False) full-resolution logits are evaluated on the fly with the arithmetic of cavp_bilinear_fwd and never written.
 * fwd: the first B images; labels int64 [B][hout][wout]; lse [B*hout*wout] (log-sum-exp per output pixel, kept for the
 * backward); partials [cavp_ce_nblocks(B, hout*wout)][2]; loss_and_count = {mean loss, #valid pixels}.
 * bwd: dx [n][hin][win][lddx] receives d loss / d x for channels < C, exact zeros for pad channels [C, Cpad) and for
 * images >= n_valid (no memset needed); gather form, deterministic.  Labels outside [0, C) other than ignore_index are
 * treated as ignored (they never index the logits). */
int cavp_upsample_ce_fwd(const float* x, int ldx, int hin, int win, int hout, int wout, int B, int C,
                         const long long* labels, int ignore_index, float* lse, float* partials, float* loss_and_count,
                         void* stream);
int cavp_upsample_ce_bwd(const float* x, int ldx, int hin, int win, int hout, int wout, int n, int n_valid, int C,
                         int Cpad, const long long* labels, int ignore_index, const float* lse,
                         const float* loss_and_count, const float* gscale, float* dx, int lddx, void* stream);
int cavp_l2norm_gather(const float* f, int ld, const long long* pix, int A, int C, float* anchors, int lda,
                       float* inv_norm, void* stream);
int cavp_l2norm_scatter_bwd(const float* danchors, const float* anchors, int lda, const float* inv_norm,
                            const long long* pix, int A, int C, float* df, int ld, void* stream);
int cavp_infonce_fwd(const float* S, int lds, const long long* labels, int A, float temperature, float* rowmax,
                     float* rowneg, float* rowmean, float* loss, void* stream);
int cavp_infonce_bwd(const float* S, int lds, const long long* labels, int A, float temperature, const float* rowmax,
                     const float* rowneg, const float* gscale, float* G, int ldg, void* stream);

/* ---- fused optimiser steps (csrc/optim.cu) -----------------------------------------------------
 * Replace torch.optim.SGD(momentum, weight_decay).step() over the 12 visual parameter groups and
 * torch.optim.Adam.step() over the audio backbone (main_vpo_mono.py:118-125, trainer/trainer_cavp_vpo_mono.py:192-193)
 * with ONE launch each.  table: device array of 48-byte rows {float* p; const float* g; float* s0; float* s1;
 * long long n; float lr; float wd} (g == NULL: skipped, like a parameter whose .grad is None); work: device array of
 * int pairs (table row, chunk index), chunk = cavp_opt_chunk_elems() elements.
 * SGD: g' = g + wd*p; s0 = momentum*s0 + g'; p -= lr*s0            (dampening 0, no nesterov; s0 starts at zero)
 * Adam: m = lerp(m, g', 1-beta1); v = beta2*v + (1-beta2)*g'^2; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps) */
int cavp_opt_chunk_elems(void);
int cavp_sgd_multi(const void* table, const int* work, int nwork, float momentum, void* stream);
int cavp_adam_multi(const void* table, const int* work, int nwork, double beta1, double beta2, double eps,
                    double bias_correction1, double bias_correction2, void* stream);
/* dst[i] = src[i] for a table of tensors in ONE launch: 24-byte rows {const float* src; float* dst; long long n}, same
 * (row, chunk) work list.  Packs the gradients that were not produced in place into the flat all-reduce buffer
 * (cavp_b200/parallel.py; the reference's DDP reducer copies into its buckets the same way, main_vpo_mono.py:131-141). */
int cavp_copy_multi(const void* table, const int* work, int nwork, void* stream);
/* same table format, dst = bf16: the bf16 copies of every weight operand of the model, one launch per step */
int cavp_cvt_bf16_multi(const void* table, const int* work, int nwork, void* stream);

/* ---- eval epilogue (csrc/metrics.cu) -----------------------------------------------------------
 * Replace torch.max(logits, 1) + 3x torch.histc (utils/eval_utils.py:73-97, MIoU) and argmax -> .cpu().numpy() ->
 * numpy.bincount (utils/eval_utils.py:107-117,151-155, ForegroundDetect) with one pass: pred[b][px] = first argmax over
 * the C classes and conf[(C+1)][C] += 1 at (label, pred); label == ignore_index or < 0 is skipped, label >= C counts in
 * row C.  conf accumulates (zero it once per evaluation).  pred / labels / conf may each be NULL (labels and conf go
 * together).  cavp_upsample_argmax_confusion takes the LOW-resolution NHWC logits of forward_cls and applies the bilinear
 * upsample of models/cavp_model.py:140 (align_corners=False) on the fly with the arithmetic of cavp_bilinear_fwd, so the
 * full-resolution logits are never written. */
int cavp_argmax_confusion(const float* logits, const long long* labels, int B, int C, long long HW, int ignore_index,
                          long long* pred, unsigned long long* conf, void* stream);
int cavp_upsample_argmax_confusion(const float* x, int ldx, int hin, int win, int hout, int wout, int n, int C,
                                   const long long* labels, int ignore_index, long long* pred, unsigned long long* conf,
                                   void* stream);

/* ---- audio front-end (csrc/audio.cu) -------------------------------------------------------------
 * Replace CAVP_TRAINER.preprocess_audio (trainer/trainer_cavp_vpo_mono.py:43-53,61-71): torchaudio MelSpectrogram
 * (n_fft 512, win 400, hop 160, 64 mels, 125-3800 Hz, power 2, center / reflect) -> first T frames -> transpose ->
 * 20*log10(max(1e-5, x)) -> 2*(x - spec_min)/(spec_max - spec_min) - 1 (utils/sourcesep.py:23-47).
 * cavp_mel_frames: frames[(r*T + t)][j] = window[j] * reflect_pad(wave_r, n_fft/2)[t*hop + j]  (window already padded
 *   to n_fft).  The real DFT is then ONE cavp_igemm against the constant [2*(n_fft/2+1)][n_fft] cos | -sin basis.
 * cavp_mel_power_db: spec rows hold Re[0..nf) | Im[nf..2nf); out[frame][m] = norm(db_scale*log10(max(amin, sum_k
 *   (Re_k^2 + Im_k^2) * fb[k][m]))), written as [rows][T][n_mels]. */
int cavp_mel_frames(const float* wave, long long ldw, int A, int rows, int T, int n_fft, int hop, const float* window,
                    float* frames, void* stream);
int cavp_mel_power_db(const float* spec, int lds, int nf, const float* fb, int n_mels, long long frames, float amin,
                      float db_scale, float spec_min, float spec_max, float* out, void* stream);

/* ---- AVContrast (csrc/avcontrast.cu; loss/av_contrast.py:20-112) --------------------------------
 * f_v [b][hw][c] normalised over hw, masked-average-pooled over the foreground pixels (mask: uint8 [b][hw] from the
 * 128x128 nearest-resized labels), contrasted with the normalised audio vectors f_a [b][c]; target[i] = the image's
 * foreground class or -1.  cavp_avc_colstats: partials [b][nchunk][2][c] (sum f^2, sum mask*f).  cavp_avc_loss: loss[0]
 * and, for a unit upstream gradient, d_fa [b][c] plus the coefficients dms / dnn [b][c] of
 * cavp_avc_bwd: dfv = g * (mask * dms + f_v * dnn)  (gscale: device scalar or NULL = 1). */
int cavp_avc_colstats(const float* fv, const unsigned char* mask, int b, int hw, int c, int nchunk, float* partials,
                      void* stream);
int cavp_avc_loss(const float* partials, int nchunk, int b, int c, const float* fa, const float* cnt, const int* target,
                  float temperature, float eps, float* feats, float* dfeat, float* nrm, float* msum, float* loss,
                  float* d_fa, float* dms, float* dnn, void* stream);
int cavp_avc_bwd(const float* fv, const unsigned char* mask, const float* dms, const float* dnn, const float* gscale,
                 int b, int hw, int c, float* dfv, void* stream);

/* ---- SyncBatchNorm statistics over NVLink peer memory (csrc/peer.cu) --------------------------------
 * Replace the two small NCCL collectives torch.nn.SyncBatchNorm issues per layer and step (main_vpo_mono.py:130).
 * Every rank allocates one exchange buffer (cavp_peer_alloc: cudaMalloc + CUDA IPC handle, 64 bytes), opens the buffers
 * of the other ranks of the node (cavp_peer_open) and then calls cavp_peer_allreduce with the same, increasing `seq`
 * (1, 2, ...) on every rank: io[0..n) (double if is_f64 else float) becomes the sum over ranks, added in rank order
 * (bit-identical on every rank).  bases[r] = rank r's buffer as mapped in THIS process; n * sizeof(element) <=
 * slot_bytes; world <= 8.  A peer that never arrives turns the result into NaN after ~30 s instead of hanging. */
int cavp_peer_alloc(long long slot_bytes, void** buf, unsigned char* handle64);
int cavp_peer_open(const unsigned char* handle64, void** buf);
int cavp_peer_close(void* buf, int own);
int cavp_peer_allreduce(void* io, int n, int is_f64, void* const* bases, int rank, int world, long long slot_bytes,
                        int seq, void* stream);

#ifdef __cplusplus
}
#endif
