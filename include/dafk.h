/*
 * dafk.h -- C ABI of the B200 (sm_100a) kernel library behind the MMSDNet / DAFNet
 * training and inference step.
 *
 * This is the drop-in boundary.  The reference
 * (agis85/multimodal_segmentation) has no FFI: its hot path is a Keras/TensorFlow
 * graph, so every entry point below replaces one TF/Keras *op call site* of the
 * reference.  Each declaration cites the reference file:line whose arithmetic it
 * reproduces.  A maintainer binds these with ctypes.
 *
 * Conventions
 *   - plain C, no torch types: device pointers + sizes + an explicit cudaStream_t
 *     passed as void* (NULL = legacy default stream);
 *   - all tensors are NHWC, contiguous; "M" is the number of pixels N*H*W;
 *   - return 0 on success, a negative DAFK_ERR_* otherwise; never throws/aborts;
 *     dafk_last_error_string() describes the last failure of the calling thread;
 *   - kernels allocate nothing: the caller owns every buffer (workspaces included);
 *   - no implicit synchronisation: everything is enqueued on the given stream and
 *     is CUDA-graph capturable;
 *   - 16-byte alignment is required for every tensor base pointer (checked).
 */
#ifndef DAFK_H_
#define DAFK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAFK_OK 0
#define DAFK_ERR_BAD_ARG (-1)
#define DAFK_ERR_ALIGN (-2)
#define DAFK_ERR_UNSUPPORTED (-3)
#define DAFK_ERR_CUDA (-4)

/* storage dtypes of feature maps */
#define DAFK_F32 0
#define DAFK_BF16 1

/* activation codes (Keras semantics: LeakyReLU'(0) = 0, ReLU'(0) = 0) */
#define DAFK_ACT_NONE 0
#define DAFK_ACT_RELU 1
#define DAFK_ACT_LRELU 2
#define DAFK_ACT_TANH 3

const char* dafk_last_error_string(void);
int dafk_version(void);
/* number of kernel launches issued through this library by the calling process */
int64_t dafk_launch_count(void);
/* cudaMemsetAsync(p, 0, bytes) on the given stream (accumulators, gradient buckets) */
int dafk_memset_zero(void* p, int64_t bytes, void* stream);

/* ------------------------------------------------------------------ rounding
 * layers/rounding.py:33-42  roundWithGrad: y = np.round(x) (round-half-to-even),
 * gradient = identity (straight-through, so there is no backward kernel). */
int dafk_round_fwd(const float* x, float* y, int64_t n, void* stream);

/* model_components/anatomy_encoder.py:23-25,58-66  softmax over the last axis
 * followed by Rounding.  p = softmax(x[M,C]); r = rint(p).  r may be NULL. */
int dafk_softmax_fwd(const float* x, float* p, float* r, int64_t M, int C, void* stream);
/* dx = p * (dp - sum_c(dp*p))  (Keras softmax gradient) */
int dafk_softmax_bwd(const float* p, const float* dp, float* dx, int64_t M, int C, void* stream);

/* ------------------------------------------------------------------ pointwise
 * Activation('relu') models/unet.py:97; LeakyReLU() alpha=.3 decoder.py:38,
 * LeakyReLU(0.2) spade.py:12, discriminator.py:25; tanh decoder.py:28.
 * y may alias x.  bwd uses the OUTPUT y (sign(y)==sign(x) for these). */
int dafk_act_fwd(const float* x, float* y, int64_t n, int act, float alpha, void* stream);
int dafk_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int act, float alpha,
                 void* stream);
/* same, with dx written as bf16 (operand dtype of the tensor-core gradient kernels); n % 4 == 0 */
/* dx = (dy1 + dy2) * act'(y): the two gradients of a map with two consumers (model_components/decoder.py:44-54: l1 feeds
 * conv2 and the residual Add) summed inside the activation backward; dx may alias dy1 or dy2; n % 4 == 0. */
int dafk_add_act_bwd(const float* dy1, const float* dy2, const float* y, float* dx, int64_t n, int act, float alpha,
                     void* stream);
int dafk_act_bwd_bf16(const float* dy, const float* y, void* dx, int64_t n, int act, float alpha, void* stream);
/* out = a + b (keras Add, decoder.py:53, spade.py:23); out may alias a or b */
int dafk_add(const float* a, const float* b, float* out, int64_t n, void* stream);
/* same for either storage dtype (gradient accumulation of bf16 feature maps) */
int dafk_add_dt(const void* a, const void* b, void* out, int dt, int64_t n, void* stream);
/* y = a*x + b*y */
int dafk_axpby(float a, const float* x, float b, float* y, int64_t n, void* stream);
int dafk_fill(float* x, float v, int64_t n, void* stream);
/* dtype conversion between DAFK_F32 and DAFK_BF16 */
int dafk_cast(const void* x, int x_dt, void* y, int y_dt, int64_t n, void* stream);
/* dst[:, dst_off:dst_off+c] (+)= src[:, src_off:src_off+c] over M pixels: Concatenate
 * (modality_encoder.py:35, stn_spline.py:104), channel slices (dafnet.py:187) and
 * their gradients (accumulate=1). */
int dafk_copy_channels(const float* src, int src_c, int src_off, float* dst, int dst_c,
                       int dst_off, int c, int64_t M, int accumulate, void* stream);
/* gather rows: dst[i,:] = src[idx[i],:]   (utils/data_utils.py:125-129 sample) */
int dafk_gather_rows(const float* src, const int32_t* idx, float* dst, int64_t rows,
                     int64_t row_elems, void* stream);

/* ------------------------------------------------------------------ FiLM
 * layers/film.py:26-36  y = x*gamma[b,c] + beta[b,c], x:[B,HW,C], gamma/beta:[B,C] */
int dafk_film_fwd(const float* x, const float* gamma, const float* beta, float* y, int B,
                  int64_t HW, int C, void* stream);
/* dx = dy*gamma ; dgamma[b,c] = sum_hw dy*x ; dbeta[b,c] = sum_hw dy.
 * dgamma/dbeta are OVERWRITTEN.  ws: B*C*2 doubles of scratch. */
int dafk_film_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma,
                  float* dbeta, double* ws, int B, int64_t HW, int C, void* stream);
/* the tail of the decoder's FiLM layer in one pass (model_components/decoder.py:50-54):
 *   y = res + act(x*gamma + beta)   (res may be NULL; act: NONE / RELU / LRELU with slope alpha)
 * backward: dy is first multiplied by act'(x*gamma + beta) (recomputed), then as dafk_film_bwd; the gradient towards
 * `res` is dy itself. */
int dafk_film_act_add_fwd(const float* x, const float* gamma, const float* beta, const float* res, float* y, int B,
                          int64_t HW, int C, int act, float alpha, void* stream);
int dafk_film_act_add_bwd(const float* dy, const float* x, const float* gamma, const float* beta, float* dx,
                          float* dgamma, float* dbeta, double* ws, int B, int64_t HW, int C, int act, float alpha,
                          void* stream);

/* bf16-storage variants (x, res, y, dy, dx are bf16; gamma, beta and their gradients fp32; C % 8 == 0, for the
 * backward C a power of two): used when the FiLM decoder keeps its activations in bf16 (engine.DEC_BF16). */
int dafk_film_act_add_fwd_bf16(const void* x, const float* gamma, const float* beta, const void* res, void* y, int B,
                               int64_t HW, int C, int act, float alpha, void* stream);
int dafk_film_act_add_bwd_bf16(const void* dy, const void* x, const float* gamma, const float* beta, void* dx,
                               float* dgamma, float* dbeta, double* ws, int B, int64_t HW, int C, int act, float alpha,
                               void* stream);
/* dx = dy * act'(y), y = the activation's OUTPUT; dy, y, dx bf16; n % 8 == 0 */
int dafk_act_bwd_bf16io(const void* dy, const void* y, void* dx, int64_t n, int act, float alpha, void* stream);

/* ------------------------------------------------------------------ Maximum
 * model_components/anatomy_fuser.py:33  keras Maximum = tf.maximum; gradient goes
 * entirely to the first input where a >= b (tie rule of tf.maximum). */
int dafk_max_fwd(const float* a, const float* b, float* out, int64_t n, void* stream);
int dafk_max_bwd(const float* a, const float* b, const float* dout, float* da, float* db,
                 int64_t n, void* stream);

/* ------------------------------------------------------------------ BatchNormalization
 * utils/model_utils.py:10, model_components/segmentor.py:17,20  (Keras defaults:
 * eps 1e-3, momentum .99, biased batch variance in training).
 * stats: acc[0:C] += sum_x, acc[C:2C] += sum_x^2 (double; caller zeroes acc).
 * x (the convolution output) is f32 or bf16 (x_dt): the tensor-core convolutions store it as bf16. */
int dafk_bn_stats(const void* x, int x_dt, double* acc, int64_t M, int C, void* stream);
/* mean/rstd from acc; moving <- moving*momentum + batch*(1-momentum) when moving_* != NULL */
int dafk_bn_finalize(const double* acc, int64_t M, int C, float eps, float momentum, float* mean,
                     float* rstd, float* moving_mean, float* moving_var, void* stream);
/* rstd = 1/sqrt(var+eps) for inference with the moving statistics */
int dafk_bn_rstd_from_var(const float* var, float* rstd, int C, float eps, void* stream);
/* out = act((x-mean)*rstd*gamma+beta); out dtype f32 or bf16; act in {NONE, RELU} */
int dafk_bn_apply(const void* x, int x_dt, const float* mean, const float* rstd, const float* gamma,
                  const float* beta, void* out, int out_dt, int64_t M, int C, int act,
                  void* stream);
/* backward, pass 1: acc[0:C] += sum dz, acc[C:2C] += sum dz*xhat, dz = dout*act'(z) */
int dafk_bn_bwd_reduce(const void* dout, int dout_dt, const void* x, int x_dt, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, double* acc,
                       int64_t M, int C, int act, void* stream);
/* backward, pass 2: dx = gamma*rstd*(dz - acc0/M - xhat*acc1/M); dgamma += acc1; dbeta += acc0
 * (dgamma/dbeta may be NULL for frozen layers).  dx dtype f32 or bf16.
 * dbias_prev (may be NULL): += sum_pixels dx, the bias gradient of the convolution feeding this BN. */
int dafk_bn_bwd_apply(const void* dout, int dout_dt, const void* x, int x_dt, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, const double* acc,
                      void* dx, int dx_dt, float* dgamma, float* dbeta, float* dbias_prev, int64_t M,
                      int C, int act, void* stream);
/* Wide BatchNorm passes (csrc/norm_wide.cuh): C a power of two in [8,1024].  One launch does the batch statistics AND
 * the finalize step of keras BatchNormalization (utils/model_utils.py:10: mean, 1/sqrt(var+eps), moving averages) /
 * the two backward sums; `ws` is a caller-owned persistent workspace of dafk_bn_wide_ws_bytes(C) bytes that must be
 * zero before the first call and is left zero by every call (kernels on one stream may share it). */
int dafk_bn_wide_supported(int C);
int64_t dafk_bn_wide_ws_bytes(int C);
int dafk_bn_stats_fused(const void* x, int x_dt, void* ws, int64_t ws_bytes, int64_t M, int C, float eps, float momentum,
                        float* mean, float* rstd, float* moving_mean, float* moving_var, void* stream);
int dafk_bn_bwd_reduce_fused(const void* dout, int dout_dt, const void* x, int x_dt, const float* mean, const float* rstd,
                             const float* gamma, const float* beta, double* acc, void* ws, int64_t ws_bytes, int64_t M,
                             int C, int act, void* stream);
/* inference-mode backward (frozen statistics): dx = dz*gamma*rstd */
int dafk_bn_bwd_frozen(const float* dout, const float* x, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, float* dx, int64_t M, int C, int act,
                       void* stream);

/* ------------------------------------------------------------------ pooling / resampling
 * MaxPooling2D(2,2) models/unet.py:39-51, stn_spline.py:108,111 (floor(H/2) outputs);
 * backward routes to the first maximum in row-major window order. */
int dafk_maxpool2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, void* stream);
int dafk_maxpool2_bwd(const void* x, const void* dy, void* dx, int dt, int N, int H, int W, int C,
                      void* stream);
/* UpSampling2D(2) utils/model_utils.py:16 (nearest repeat); bwd sums each 2x2 block */
int dafk_upsample2_fwd(const void* x, void* y, int dt, int N, int H, int W, int C, void* stream);
int dafk_upsample2_bwd(const void* dy, void* dx, int dt, int N, int H, int W, int C, void* stream);
/* tf.image.resize_nearest_neighbor (layers/spade.py:36-38): src = floor(dst*in/out) */
int dafk_resize_nn_fwd(const float* x, float* y, int N, int H, int W, int C, int Ho, int Wo,
                       void* stream);
int dafk_resize_nn_bwd(const float* dy, float* dx, int N, int H, int W, int C, int Ho, int Wo,
                       void* stream);

/* ------------------------------------------------------------------ convolution
 * keras Conv2D call sites: models/unet.py:95,99; utils/model_utils.py:17;
 * model_components/{anatomy_encoder.py:23,102-154, segmentor.py:15-24,
 * modality_encoder.py:36-42, decoder.py:28,45-48}; layers/stn_spline.py:106-112;
 * layers/spade.py:14-31; models/discriminator.py:24,39.
 * x:[N,H,W,Cin], w: HWIO [KH,KW,Cin,Cout] (Keras layout), y:[N,Ho,Wo,Cout],
 * Ho = (H + 2*pad - KH)/stride + 1.  Cross-correlation. */
typedef struct dafk_conv_desc {
  int32_t N, H, W, Cin;
  int32_t Cout, KH, KW;
  int32_t stride, pad;
  int32_t Ho, Wo;
} dafk_conv_desc;

/* general CUDA-core path (any shape, fp32): y = act(conv(x,w)+bias) */
int dafk_conv2d_fwd(const dafk_conv_desc* d, const float* x, const float* w, const float* bias,
                    float* y, int act, float alpha, void* stream);
/* dx (overwritten) = conv_transpose(dy, w) */
int dafk_conv2d_dgrad(const dafk_conv_desc* d, const float* dy, const float* w, float* dx,
                      void* stream);
/* dw += x (*) dy  (accumulates: caller zeroes);  db += sum_pixels dy (db may be NULL) */
int dafk_conv2d_wgrad(const dafk_conv_desc* d, const float* x, const float* dy, float* dw,
                      float* db, void* stream);
/* direct kernels for the narrow layers (few input and/or output channels: FiLM decoder 8->8,
 * first UNet/segmentor/discriminator layers, 1x1 heads, locnet 5x5, modality encoder); same
 * semantics as dafk_conv2d_{fwd,dgrad,wgrad}.  dafk_conv_small_supported() tells whether the filter
 * bank fits in shared memory; the strided data gradient additionally needs Cin <= 16. */
int dafk_conv_small_supported(int Cin, int Cout, int KH, int KW);
int dafk_conv_small_fwd(const dafk_conv_desc* d, const float* x, const float* w, const float* bias,
                        float* y, int act, float alpha, void* stream);
int dafk_conv_small_dgrad(const dafk_conv_desc* d, const float* dy, const float* w, float* dx,
                          void* stream);
int dafk_conv_small_wgrad(const dafk_conv_desc* d, const float* x, const float* dy, float* dw,
                          float* db, void* stream);
/* Narrow-channel convolutions on tcgen05 (csrc/conv_nc.cu): stride 1, any KH x KW (KW <= 8 for the weight
 * gradient), few input channels (FiLM decoder 8->8 model_components/decoder.py:44-54, segmentor / UNet /
 * discriminator first layers, locnet 5x5 layers/stn_spline.py:106-112).  x is f32 or bf16 NHWC and is
 * converted to bf16 while it is staged; accumulation is fp32 in tensor memory.
 * kind: 0 forward, 1 data gradient, 2 weight gradient.  Returns 1 if the geometry fits in shared memory. */
int dafk_conv_nc_supported(int Cin, int Cout, int KH, int KW, int W, int pad, int kind);
/* 1 if dafk_conv_nc_wgrad would bring x / dy rows of these dtypes (DAFK_F32 / DAFK_BF16) into shared memory with bulk
 * copies (large maps with 16-byte-multiple rows): a bf16 output gradient is then read at half the bytes, while the
 * register-staged kernel is slower on bf16 than on fp32.  The host engine asks before it lets the BatchNormalization
 * backward of a first layer (models/unet.py:95, model_components/segmentor.py:15) write its gradient in bf16. */
int dafk_conv_nc_wgrad_stages_raw(int N, int H, int W, int Cin, int Cout, int KH, int KW, int pad, int x_dt, int dy_dt);
/* The launch plan dafk_conv_nc_fwd (kind 0; dy_dt = the output dtype) or dafk_conv_nc_wgrad (kind 2) would use for this
 * shape, without launching anything (host only): plan[0..9] = {bulk-copy staging 0/1, rows per strip, raster stages,
 * ring slots, bytes per slot, pixels per segment and segments per row of x, the same of dy (kind 2), dynamic shared
 * memory bytes}.  DAFK_ERR_UNSUPPORTED if the geometry does not fit.  Tests assert the plan's invariants on the CPU. */
int dafk_conv_nc_plan(int kind, int N, int H, int W, int Cin, int Cout, int KH, int KW, int pad, int x_dt, int dy_dt,
                      int64_t* plan);
/* number of bf16 elements of the packed weight buffer for a kernel that reduces over Cin_k channels
 * and produces Cout_k channels */
int64_t dafk_conv_nc_packed_elems(int Cin_k, int Cout_k, int KH, int KW);
/* HWIO f32 -> packed bf16.  mode 0: forward operand.  mode 1: operand of the stride-1 data gradient
 * (taps mirrored, channels transposed): run dafk_conv_nc_fwd on dy with Cin=Cout_layer, Cout=Cin_layer,
 * pad = K-1-pad_layer. */
int dafk_pack_conv_nc(const float* w_hwio, void* wp, int KH, int KW, int Cin, int Cout, int mode,
                      void* stream);
/* forward operand of w * scale[Cout] (inference-mode BatchNorm folded into a narrow layer, see dafk_bn_fold) */
int dafk_pack_conv_nc_scaled(const float* w_hwio, const float* scale, void* wp, int KH, int KW, int Cin, int Cout,
                             void* stream);
/* y[N,Ho,Wo,Cout] = act(conv(x, w) + bias), Ho = H + 2*pad - KH + 1; y is f32 or bf16 */
int dafk_conv_nc_fwd(const void* x, int x_dt, const void* wp, const float* bias, void* y, int y_dt, int N,
                     int H, int W, int Cin, int Cout, int KH, int KW, int pad, int act, float alpha,
                     void* stream);
/* dw[KH,KW,Cin,Cout] (HWIO f32) += x (*) dy ; db[Cout] += sum_pixels dy (db may be NULL) */
int dafk_conv_nc_wgrad(const void* x, int x_dt, const void* dy, int dy_dt, float* dw, float* db, int N,
                       int H, int W, int Cin, int Cout, int KH, int KW, int pad, void* stream);
/* The same two kernels on Concatenate([xa, xb]) without materialising it (the locnet's input, layers/stn_spline.py:104:
 * two 8-channel anatomies): channels [0,Ca) of the kernel's Cin = Ca + Cb come from xa [N,H,W,Ca], the rest from xb
 * [N,H,W,Cb]; both sources have dtype x_dt; Ca must be a multiple of 8 (a channel group never straddles the sources). */
int dafk_conv_nc_fwd_cat(const void* xa, int Ca, const void* xb, int Cb, int x_dt, const void* wp, const float* bias,
                         void* y, int y_dt, int N, int H, int W, int Cout, int KH, int KW, int pad, int act, float alpha,
                         void* stream);
int dafk_conv_nc_wgrad_cat(const void* xa, int Ca, const void* xb, int Cb, int x_dt, const void* dy, int dy_dt, float* dw,
                           float* db, int N, int H, int W, int Cout, int KH, int KW, int pad, void* stream);
/* Pointwise heads on a 64-channel bf16 feature map (csrc/conv_1x1.cu): `conv_anatomy` 64 -> 8
 * (model_components/anatomy_encoder.py:26) and the segmentor's 64 -> num_masks+1 (model_components/segmentor.py:24).
 * x / dx: bf16 [M,64]; w: f32 [64,Cout] (HWIO with KH=KW=1); y / dy: f32 [M,Cout]; M = N*H*W pixels.
 * round_bf16 != 0: w and dy are rounded to bf16 as they are loaded (operand precision of the tensor-core mode).
 * wgrad accumulates into dw / db (db may be NULL). */
int dafk_conv1x1_supported(int Cin, int Cout);
int dafk_conv1x1_fwd(const void* x, const float* w, const float* bias, float* y, int64_t M, int Cin, int Cout,
                     int round_bf16, void* stream);
int dafk_conv1x1_dgrad(const float* dy, const float* w, void* dx, int64_t M, int Cin, int Cout, int round_bf16,
                       void* stream);
int dafk_conv1x1_wgrad(const void* x, const float* dy, float* dw, float* db, int64_t M, int Cin, int Cout,
                       int round_bf16, void* stream);
/* Space-to-depth (2x2 pixel blocks -> 4C channels, bf16 out, zero beyond odd sizes), its inverse, and the matching
 * rearrangement of an HWIO kernel [KH,KW,C,Cout] -> [ceil(KH/2),ceil(KW/2),4C,Cout] (backward != 0: accumulate the
 * rearranged gradient w2 back into w).  conv(x, w, stride 2, valid) == conv(s2d(x), s2d(w), stride 1, valid):
 * the stride-2 narrow layers (model_components/modality_encoder.py:36-42, models/discriminator.py:24) run on
 * dafk_conv_nc_* this way. */
int dafk_space_to_depth2(const void* x, int x_dt, void* y_bf16, int N, int H, int W, int C, void* stream);
/* The same rearrangement of Concatenate([xa, xb]) (model_components/modality_encoder.py:34) without materialising the
 * concatenation: channels [0,Ca) of every pixel come from xa [N,H,W,Ca], [Ca,Ca+Cb) from xb [N,H,W,Cb]; and the backward,
 * which writes the fp32 gradients of the two sources straight from the rearranged gradient (ga / gb may be NULL). */
int dafk_space_to_depth2_cat(const void* xa, int xa_dt, int Ca, const void* xb, int xb_dt, int Cb, void* y_bf16, int N,
                             int H, int W, void* stream);
int dafk_depth_to_space2_split(const void* y, int y_dt, float* ga, int Ca, float* gb, int Cb, int N, int H, int W,
                               void* stream);
int dafk_depth_to_space2(const void* y, int y_dt, void* x, int x_dt, int N, int H, int W, int C,
                         void* stream);
int dafk_conv_s2d_weights(float* w, float* w2, int KH, int KW, int C, int Cout, int backward,
                          void* stream);
/* out[c] += sum_m x[m,c]   (bias gradients); x is f32 or bf16 */
int dafk_colsum(const void* x, int x_dt, float* out, int64_t M, int C, void* stream);

/* tcgen05 / TMEM / TMA implicit-GEMM path: 3x3, stride 1, pad 1, bf16 operands,
 * fp32 accumulation in tensor memory.  Cin % 64 == 0 and Cout % 64 == 0.
 * x0:[N,H,W,C0] (+ optional second K source x1:[N,H,W,C1] = Concatenate([x0,x1]),
 * models/unet.py:68-69), wp: packed bf16 weights [9][Cout][C0+C1] (dafk_pack_conv3x3).
 * y: [N,H,W,Cout] f32 or bf16.  Returns DAFK_ERR_UNSUPPORTED for other shapes.
 * The packed weight matrix has w_rows_per_tap rows per tap; this call produces the Cout output
 * channels whose rows start at w_row_off (w_rows_per_tap = Cout, w_row_off = 0 for a plain
 * forward; a data-gradient towards one source of a Concatenate uses a row window). */
/* Data gradient of a VALID stride-2 convolution with an even kernel (models/discriminator.py:24,39: 4x4 stride 2) as
 * ONE launch over the four output-parity classes dx[:, pa::2, pb::2, :] (each is a stride-1 convolution of dy with
 * KH/2 x KW/2 taps).  dy: bf16 [N,Ho,Wo,Cout]; wp4: the four dafk_pack_conv(mode 2, pa, pb) matrices stored back to back
 * in class order c = 2*pa + pb; w_rows_per_tap / w_row_off select the Cin rows of this source as in dafk_conv_tc_fwd;
 * dx: [N,H,W,Cin] f32 or bf16, every element written (rows / columns the convolution never read get 0). */
int dafk_conv_tc_dgrad_s2(const void* dy, int Cout, const void* wp4, int w_rows_per_tap, int w_row_off, void* dx,
                          int dx_dt, int N, int Ho, int Wo, int Cin, int KH, int KW, int H, int W, void* stream);
/* dafk_conv_tc_fwd with a fused epilogue activation (NONE or RELU), and the folding of an inference-mode
 * BatchNormalization (utils/model_utils.py:10, Keras learning phase 0) into the convolution that feeds it:
 *   dafk_bn_fold:          scale[c] = gamma / sqrt(moving_var + eps),  bias_out[c] = (conv_bias - moving_mean) * scale + beta
 *   dafk_pack_conv_scaled: forward operand of w[KH,KW,Cin,Cout] * scale[Cout] (as dafk_pack_conv mode 0)
 * conv -> BN -> ReLU of a predict pass then is ONE kernel: dafk_conv_tc_fwd_act(..., bias_out, ..., DAFK_ACT_RELU, 0).
 * DAFK_ACT_LRELU with `alpha`: the discriminator's Conv2D -> LeakyReLU(0.2) pairs (models/discriminator.py:24-25,39-40),
 * whose convolution output is only ever used through the activation. */
int dafk_conv_tc_fwd_act(const void* x0, int C0, const void* x1, int C1, const void* wp, int w_rows_per_tap,
                         int w_row_off, const float* bias, void* y, int y_dt, int N, int H, int W, int Cout, int KH,
                         int KW, int stride, int pad, int Ho, int Wo, int64_t y_sn, int64_t y_sy, int64_t y_sx, int act,
                         float alpha, void* stream);
/* Conv2D -> BatchNormalization in the training phase (models/unet.py:95-96, utils/model_utils.py:10): the convolution
 * writes its output y [N,Ho,Wo,Cout] in bf16 and ALSO accumulates the batch statistics of exactly those stored values,
 * bn_acc[c] += sum y[..,c], bn_acc[Cout + c] += sum y[..,c]^2 (fp64, caller zeroes it), from its epilogue; dafk_bn_finalize
 * turns them into mean / rstd and updates the moving statistics -- the separate statistics pass over y disappears. */
int dafk_conv_tc_fwd_bn(const void* x0, int C0, const void* x1, int C1, const void* wp, int w_rows_per_tap, int w_row_off,
                        const float* bias, void* y_bf16, double* bn_acc, int N, int H, int W, int Cout, int KH, int KW,
                        int stride, int pad, void* stream);
int dafk_bn_fold(const float* gamma, const float* beta, const float* moving_mean, const float* moving_var,
                 const float* conv_bias, float eps, float* scale, float* bias_out, int C, void* stream);
int dafk_pack_conv_scaled(const float* w_hwio, const float* scale, void* wp, int KH, int KW, int Cin, int Cout, void* stream);
int dafk_conv3x3_tc_fwd(const void* x0, int C0, const void* x1, int C1, const void* wp,
                        int w_rows_per_tap, int w_row_off, const float* bias, void* y, int y_dt,
                        int N, int H, int W, int Cout, void* stream);
/* General form: KH x KW kernel, stride 1 or 2 (TMA traversal stride), symmetric padding `pad`,
 * logical output Ho x Wo written through explicit element strides (y_sn, y_sy, y_sx) so that the
 * four parity classes of a stride-2 data gradient can be scattered into one dx tensor.
 * Serves models/discriminator.py:24,39 (4x4, stride 2 / 1, valid) on the tensor cores. */
int dafk_conv_tc_fwd(const void* x0, int C0, const void* x1, int C1, const void* wp,
                     int w_rows_per_tap, int w_row_off, const float* bias, void* y, int y_dt, int N,
                     int H, int W, int Cout, int KH, int KW, int stride, int pad, int Ho, int Wo,
                     int64_t y_sn, int64_t y_sy, int64_t y_sx, void* stream);
/* weight packing for the tensor-core path.  mode 0: [tap][Cout][Cin] (forward); mode 1: mirrored
 * taps, [tap][Cin][Cout] (data gradient of a stride-1 conv); mode 2: the (pa,pb) parity class of
 * the data gradient of a stride-2 conv, [(KH/2)*(KW/2)][Cin][Cout]. */
int dafk_pack_conv(const float* w_hwio, void* wp, int KH, int KW, int Cin, int Cout, int mode, int pa,
                   int pb, void* stream);
int dafk_conv_tc_wgrad(const void* x, int Cin, int cin_off, int cin_total, const void* dy, int Cout,
                       float* dw, int N, int H, int W, int KH, int KW, int stride, int pad, int Ho,
                       int Wo, void* stream);
/* weights HWIO f32 [3,3,Cin,Cout] -> bf16 [9][Cout][Cin] (fwd) or flipped/transposed
 * [9][Cin][Cout] with tap index mirrored (dgrad) */
int dafk_pack_conv3x3(const float* w_hwio, void* wp, int Cin, int Cout, int for_dgrad,
                      void* stream);
/* dw[3,3,Cin,Cout] (HWIO f32) += sum_pixels x (*) dy, tensor-core path */
int dafk_conv3x3_tc_wgrad(const void* x, int Cin, int cin_off, int cin_total, const void* dy,
                          int Cout, float* dw, int N, int H, int W, void* stream);
/* Same result as dafk_conv3x3_tc_wgrad (dw[3,3,cin_total,Cout] += x (*) dy for the channel block at cin_off) for
 * Cin and Cout multiples of 64 (csrc/conv_tc_wgrad_halo.cu): one CTA owns all nine taps of a 64 x 64 block, loads a
 * haloed X tile once per 16 x 8 pixel tile and pairs two taps per M=128 MMA.  Preferred for the 64/128-channel
 * full-resolution layers (models/unet.py:95,99), where the per-tap kernel is bound by L2 -> SM bandwidth. */
int dafk_conv3x3_tc_wgrad_halo_supported(int Cin, int Cout);
int dafk_conv3x3_tc_wgrad_halo(const void* x, int Cin, int cin_off, int cin_total, const void* dy, int Cout, float* dw,
                               int N, int H, int W, void* stream);

/* ------------------------------------------------------------------ Dense
 * keras Dense: modality_encoder.py:46-50, stn_spline.py:115-116, discriminator.py:33,
 * decoder.py:37-40,68, balancer.py:24-25.  y[B,Nout] = x[B,K] w[K,Nout] + b */
int dafk_dense_fwd(const float* x, const float* w, const float* bias, float* y, int B, int64_t K,
                   int Nout, void* stream);
/* dx[B,K] = dy[B,Nout] w^T */
int dafk_dense_bwd_data(const float* dy, const float* w, float* dx, int B, int64_t K, int Nout,
                        void* stream);
/* dw[K,Nout] += x^T dy ; db[Nout] += sum_b dy */
int dafk_dense_bwd_weight(const float* x, const float* dy, float* dw, float* db, int B, int64_t K,
                          int Nout, void* stream);

/* ------------------------------------------------------------------ thin-plate spline STN
 * layers/stn_spline.py:14-91 + layers/interpolate_spline.py:30-278 +
 * tf.contrib.resampler (call site stn_spline.py:65).
 *
 * General batched polyharmonic solve (interpolate_spline.py:76-147): for each b builds the
 * (n+d+1)^2 system with phi(r^2) (order 1,2,4 supported), LU with partial pivoting
 * (tf.matrix_solve) and returns w[b,n,k], v[b,d+1,k].  n+d+1 <= 32, d == 2. */
int dafk_tps_solve_batched(const float* train_points, const float* train_values, float* w_out,
                           float* v_out, int B, int n, int k, int order, float reg, void* stream);
/* evaluate (interpolate_spline.py:150-179): out[b,m,k] = phi(|q-c|^2) w + [q,1] v */
int dafk_tps_apply(const float* query, const float* train_points, const float* w, const float* v,
                   float* out, int B, int64_t m, int n, int k, int order, int query_batched,
                   void* stream);
/* Fast path for ThinPlateSpline2D(inverse=False) (the only mode the reference uses,
 * anatomy_fuser.py:30): the LHS of the spline system depends only on the constant control-point
 * grid, so w_b = Winv.theta_b and v_b = v_identity + Vinv.theta_b.  The HOST builds, in fp64, the
 * constant block  [ c(n,2) | Winv(n,n) | Vinv(3,n) ]  (dafk_tps_consts_floats(n) floats) into a
 * caller-provided host buffer; the caller uploads it once per control-point grid. */
int dafk_tps_consts_floats(int n_cp);
int dafk_tps_build_constants(int cp_h, int cp_w, float* consts_host);
/* fused spline evaluation + bilinear gather (stn_spline.py:55-67):
 * out[b,y,x,:] = resample(vol[b], (X,Y)(b,y,x)); also writes locs[b,m,2] = (x,y) in pixels when
 * locs != NULL.  theta:[B,n_cp,2] control-point offsets in (row,col) normalised units. */
int dafk_tps_warp_fwd(const float* vol, const float* theta, const float* consts, float* out,
                      float* locs, int B, int H, int W, int C, int n_cp, void* stream);
/* Same warp with phi(|q - c|^2) read from a per-geometry table instead of evaluated per pixel (25 logf):
 * dafk_tps_phi_table fills table[n_cp + 2][H*W] (dafk_tps_phi_table_floats floats, caller-owned, reusable for every call
 * with the same H, W and control grid) with exactly the values the in-kernel evaluation produces; rows n_cp and n_cp + 1
 * hold the normalised pixel coordinates row/(H-1), col/(W-1) that the affine part of the spline reads.  The spline
 * coefficients of the batch are computed once per call into coef_ws (B*(n_cp+3)*2 floats, caller-owned) instead of
 * once per CTA. */
int64_t dafk_tps_phi_table_floats(int H, int W, int n_cp);
int dafk_tps_phi_table(const float* consts, float* table, int H, int W, int n_cp, void* stream);
int dafk_tps_warp_fwd_tab(const float* vol, const float* theta, const float* consts, const float* phi_table,
                          float* coef_ws, float* out, float* locs, int B, int H, int W, int C, int n_cp, void* stream);
/* backward: dvol += scatter(dout) (caller zeroes dvol; NULL skips it);
 * dtheta[b,n,2] (overwritten) via the resampler's analytic coordinate gradient.
 * ws: B*(n_cp+3)*2 doubles (zeroed by the callee). */
int dafk_tps_warp_bwd(const float* vol, const float* theta, const float* consts, const float* dout,
                      float* dvol, float* dtheta, double* ws, int B, int H, int W, int C, int n_cp,
                      void* stream);
/* the same with phi and the pixel coordinates read from the geometry's table (dafk_tps_phi_table: (n_cp + 2) rows) */
int dafk_tps_warp_bwd_tab(const float* vol, const float* theta, const float* consts, const float* phi_table,
                          const float* dout, float* dvol, float* dtheta, double* ws, int B, int H, int W, int C, int n_cp,
                          void* stream);
/* plain tf.contrib.resampler forward for arbitrary warp[b,m,2] */
int dafk_resampler_fwd(const float* vol, const float* warp, float* out, int B, int H, int W, int C,
                       int64_t m, void* stream);

/* ------------------------------------------------------------------ losses (costs.py)
 * All losses write a scalar (Keras mean-reduced) into loss[0] as  loss[0] += weight*value
 * and produce the gradient of  weight*value  w.r.t. the prediction.
 *
 * costs.py:43-67 dice over the first `nch` channels, per-sample, smooth 1e-12, mean over B.
 * costs.py:70-85 + :129-136 weighted cross entropy with swapped arguments, lambda_bce .01.
 * pred:[B,HW,Cp], target:[B,HW,Ct]; ws: doubles, size dafk_segloss_ws_doubles(B,Cp). */
int64_t dafk_segloss_ws_doubles(int B, int C);
int dafk_segloss_fwd(const float* pred, int Cp, const float* target, int Ct, int nch, int use_bce,
                     float lambda_bce, double* ws, int B, int64_t HW, void* stream);
int dafk_segloss_finish(const double* ws, float weight, float* loss, int B, int Cp, int nch,
                        int use_bce, float lambda_bce, int64_t HW, void* stream);
int dafk_segloss_bwd(const float* pred, int Cp, const float* target, int Ct, int nch, int use_bce,
                     float lambda_bce, const double* ws, float weight, float* dpred, int B,
                     int64_t HW, void* stream);
/* keras 'mae' / 'mse' (mean over all elements); kind 0 = mae, 1 = mse.  target == NULL means a
 * constant target `cval`.  loss[0] += weight*value; dpred (may be NULL) is overwritten with the
 * gradient of weight*value (tf.abs gradient = sign, 0 at 0). */
int dafk_l1l2_loss(const float* pred, const float* target, float cval, int kind, float weight,
                   float* loss, float* dpred, int64_t n, void* stream);
/* utils/sdnet_utils.py:9-21 sampling + costs.py:186-189 kl + costs.py:194 ypred (mean):
 * z = mu + exp(.5*lv)*eps ; kl[b] = -.5*sum(1+lv-mu^2-exp(lv)) ; loss += weight*mean(kl) */
int dafk_vae_fwd(const float* mu, const float* logvar, const float* eps, float* z, float* kl,
                 float weight, float* loss, int B, int Z, void* stream);
/* dmu = dz + weight/B * mu ; dlv = dz*.5*exp(.5 lv)*eps + weight/B*.5*(exp(lv)-1) */
int dafk_vae_bwd(const float* mu, const float* logvar, const float* eps, const float* dz,
                 float weight, float* dmu, float* dlogvar, int B, int Z, void* stream);

/* ------------------------------------------------------------------ Spectral regulariser
 * layers/spectralnorm.py:199-246.  W:[dim,cout]; 3 power iterations from u0; loss +=
 * alpha*mean|W/sigma - W| ; dW += -alpha*sign(W/sigma - W)/numel (sigma is stop_gradient).
 * ws: (2*dim + cout + 4) floats. */
int dafk_spectral_reg(const float* W, const float* u0, float alpha, float* loss, float* dW,
                      float* ws, int dim, int cout, void* stream);

/* ------------------------------------------------------------------ Adam (Keras 2.1.6)
 * call sites models/dafnet.py:93,114,155,161,349.  One launch over a flat parameter bucket.
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed on the host; g is multiplied by grad_scale first
 * (1/world_size after the NCCL all-reduce).  bf16_shadow (may be NULL) receives p as bf16. */
int dafk_adam_step(float* p, const float* g, float* m, float* v, void* bf16_shadow, int64_t n,
                   float lr_t, float beta1, float beta2, float eps, float grad_scale,
                   void* stream);
/* Same update with the schedule kept on the device (CUDA-graph capturable optimizer step):
 * state[0] = t, state[1] = lr_t.  dafk_adam_tick advances t and recomputes lr_t (in double);
 * dafk_adam_step_dev reads lr_t = state[1]. */
int dafk_adam_tick(float* state, float lr, float beta1, float beta2, void* stream);
int dafk_adam_step_dev(float* p, const float* g, float* m, float* v, void* bf16_shadow, int64_t n,
                       const float* state, float beta1, float beta2, float eps, float grad_scale,
                       void* stream);

/* ------------------------------------------------------------------ instance norm + SPADE
 * keras_contrib InstanceNormalization(axis=None, scale=False, center=False) layers/spade.py:27:
 * per-sample statistics over H,W,C jointly, (x-mean)/(std+1e-3).
 * stats: acc[b*2+{0,1}] += sum, sum_sq (double, caller zeroes). */
int dafk_in_stats(const float* x, double* acc, int B, int64_t HWC, void* stream);
/* layers/spade.py:41-55 SPADE_COND fused with the normalisation and LeakyReLU(0.2):
 * xn = (x-mean_b)/(std_b+eps); y = act(xn*(1+gamma)+beta) */
int dafk_spade_fwd(const float* x, const double* acc, const float* gamma, const float* beta,
                   float* y, int B, int64_t HWC, float eps, int act, float alpha, void* stream);
/* backward of the fused op: produces dgamma, dbeta, and dx (through the instance norm).
 * ws: B*2 doubles zeroed by the callee. */
int dafk_spade_bwd(const float* dy, const float* x, const double* acc, const float* gamma,
                   const float* beta, float* dx, float* dgamma, float* dbeta, double* ws, int B,
                   int64_t HWC, float eps, int act, float alpha, void* stream);

/* layers/spade.py:41-58 SPADE_COND as a stand-alone layer (SPADE_COND()([x, gamma, beta]) on an already-normalised x):
 * y = x*(1+gamma)+beta, all four tensors f32 with n elements.  Backward: dx = dy*(1+gamma), dgamma = dy*x; dbeta is dy
 * itself (the caller aliases it). */
int dafk_spade_cond_fwd(const float* x, const float* gamma, const float* beta, float* y, int64_t n, void* stream);
int dafk_spade_cond_bwd(const float* dy, const float* x, const float* gamma, float* dx, float* dgamma, int64_t n,
                        void* stream);
/* utils/model_utils.py:6-12 normalise('instance') = keras_contrib InstanceNormalization() with its defaults (axis=None,
 * epsilon 1e-3, center and scale: gamma, beta of shape (1,)): y = act(gamma[0]*(x-mean_b)/(std_b+eps) + beta[0]) with the
 * per-sample statistics of dafk_in_stats.  Backward: dx through the normalisation, dgamma[0] / dbeta[0] ACCUMULATED
 * (either may be NULL).  ws: 2*B+2 doubles, zeroed by the callee. */
int dafk_in_affine_fwd(const float* x, const double* acc, const float* gamma, const float* beta, float* y, int B,
                       int64_t HWC, float eps, int act, float alpha, void* stream);
int dafk_in_affine_bwd(const float* dy, const float* x, const double* acc, const float* gamma, const float* beta,
                       float* dx, float* dgamma, float* dbeta, double* ws, int B, int64_t HWC, float eps, int act,
                       float alpha, void* stream);

/* ------------------------------------------------------------------ balancer
 * model_components/balancer.py:33-38 soft dice between two anatomies per sample:
 * out[b] = (2*sum(a*b)+1e-12)/(sum(a)+sum(b)+1e-12).  ws: B*3 doubles. */
int dafk_pair_dice(const float* a, const float* b, float* out, double* ws, int B, int64_t HWC,
                   void* stream);

/* ---- automated-pairing trainers (models/dafnet.py:224-334; csrc/pairing.cu) --------------------------------------
 * Per-sample losses L[b] whose Balancer-weighted sum over the candidate pairs is the `SegmentorDef` / `DecoderDef`
 * output (loss costs.ypred = mean over the batch).  The backward passes take per-sample coefficients coef[b]
 * (= loss_weight / B * w[b,j], written by dafk_pair_combine).
 * dafk_segloss_pb_*: costs.make_combined_dice_bce_perbatch (costs.py:138-143) = soft Dice over the first `nch`
 *   channels + lambda_bce * weighted_cross_entropy_perbatch with the reference's swapped arguments (costs.py:88-108):
 *   class weights from the prediction summed over the batch, log of softmax(mask); pred/target f32 [B,HW,C], C <= 8.
 *   ws: dafk_segloss_pb_ws_doubles(B,C) doubles, written by fwd, read by bwd.
 * dafk_mae_pb_*: costs.mae_single_input (costs.py:24-26), mean |pred - target| over the n = H*W*C values of a sample.
 * dafk_pair_dice_bwd: backward of dafk_pair_dice (model_components/balancer.py:33-38); ws as left by the forward;
 *   g[B] = d loss / d dice; da or db may be NULL.
 * dafk_pair_combine: loss[0] += weight/B * sum_b sum_j w[b,j] L[j,b]; dw[b,j] = weight/B * L[j,b];
 *   coef[j,b] = weight/B * w[b,j]   (w == NULL: all ones; dw may be NULL).  w [B,P], L and coef [P,B]. */
int64_t dafk_segloss_pb_ws_doubles(int B, int C);
int dafk_segloss_pb_fwd(const float* pred, const float* target, int C, int nch, float lambda_bce, double* ws, float* L,
                        int B, int64_t HW, void* stream);
int dafk_segloss_pb_bwd(const float* target, int C, int nch, float lambda_bce, const double* ws, const float* coef,
                        float* dpred, int B, int64_t HW, void* stream);
int dafk_mae_pb_fwd(const float* pred, const float* target, double* ws, float* L, int B, int64_t n, void* stream);
int dafk_mae_pb_bwd(const float* pred, const float* target, const float* coef, float* dpred, int B, int64_t n,
                    void* stream);
int dafk_pair_dice_bwd(const float* a, const float* b, const double* ws, const float* g, float* da, float* db, int B,
                       int64_t HWC, void* stream);
int dafk_pair_combine(const float* w, const float* L, float weight, float* loss, float* dw, float* coef, int B, int P,
                      void* stream);

/* Augmentation of a staged batch (model_executors/base_executor.py:37-78,103-110: keras ImageDataGenerator with
 * rotation_range=20 -> scipy.ndimage.affine_transform(order=1, mode='nearest') about the image centre).
 * x, y: f32 [B,H,W,C] (C in 1..5 or 8), distinct buffers; theta[B]: rotation angle per sample in radians. */
int dafk_rotate_bilinear(const float* x, const float* theta, float* y, int B, int H, int W, int C, void* stream);
/* Executor.add_residual (model_executors/base_executor.py:83-87) applied to the AUGMENTED batch as the reference does
 * (dafnet_executor.py:493-494): m[p, C-1] = 0 if any m[p, c < C-1] == 1 else 1, in place; m: f32 [pixels, C]. */
int dafk_mask_residual(float* m, int64_t pixels, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAFK_H_ */
