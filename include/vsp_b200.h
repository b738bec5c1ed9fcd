/*
 * vsp_b200.h — C ABI of the B200-native (sm_100a) synthesis hot path of VSPBFR.
 *
 * Every entry point is `extern "C"`, takes raw device pointers + int64 sizes +
 * an explicit `cudaStream_t` (as void*), never allocates or frees user-visible
 * memory, never synchronises, and returns 0 on success (non-zero: call
 * vsp_last_error()).  Entry points are re-entrant and CUDA-graph capturable.
 *
 * Each function cites the reference interface (file:line under the VSPBFR
 * tree) it stands in for.  The Python host side (vspbfr_b200/op/*) binds this
 * header with ctypes; INTEGRATION.md shows the stub a reference maintainer
 * would add.
 */
#ifndef VSP_B200_H_
#define VSP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSP_ABI_VERSION 1

/* ---- bookkeeping ------------------------------------------------------- */

/* ABI version of the loaded library (== VSP_ABI_VERSION it was built with). */
int vsp_version(void);

/* Thread-local, NUL-terminated message for the last non-zero return on the
 * calling thread.  Replaces TORCH_CHECK -> RuntimeError of
 * op/upfirdn2d.cpp:9-15 and op/fused_bias_act.cpp:10-16. */
const char *vsp_last_error(void);

/* Number of kernels this library has launched since load (all threads).
 * bench.py reads it to report `gpu_launches`. */
int64_t vsp_launch_count(void);

/* ---- upfirdn2d --------------------------------------------------------- */

/*
 * Up-sample (zero-stuff) / FIR-filter / down-sample of `major` fp32 planes.
 * Replaces the pybind entry `upfirdn2d.upfirdn2d(input, kernel, up_x, up_y,
 * down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)` of op/upfirdn2d.cpp:17-30
 * (launcher op/upfirdn2d_kernel.cu:209-368) with minor_dim == 1, i.e. the
 * `[N*C, H, W, 1]` view op/upfirdn2d.py:297 always passes.
 *
 *   x     [major, in_h, in_w]   fp32, contiguous
 *   filt  [kh, kw]              fp32, contiguous (NOT flipped; the op is a true
 *                               convolution, op/upfirdn2d.py:388)
 *   y     [major, out_h, out_w] fp32, out = (in*up + pad0 + pad1 - k + down)/down
 *
 * Optional fused epilogue (north_star "fused upfirdn+bias+lrelu"): when
 * `act` != 0 the kernel applies y = lrelu(y + bias[plane % channels], alpha)
 * * scale before the store (bias may be NULL = no bias).  act == 0: plain op.
 * Negative pads crop, as in op/upfirdn2d.py:378-386.
 */
int vsp_upfirdn2d_f32(const float *x, const float *filt, float *y,
                      int64_t major, int64_t in_h, int64_t in_w,
                      int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                      int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                      const float *bias, int64_t channels, int act,
                      float alpha, float scale, void *stream);

/* Output extent helper: (in*up + pad0 + pad1 - k + down) / down
 * (op/upfirdn2d.py:301-302). Returns <= 0 for an empty output. */
int64_t vsp_upfirdn2d_out_size(int64_t in, int k, int up, int down, int pad0, int pad1);

/*
 * Channels-last bf16 variant used inside fused layer chains:
 *   x [n, in_h, in_w, c] bf16 -> y [n, out_h, out_w, c] bf16 (c % 8 == 0).
 * Same arithmetic (fp32 accumulate), same parameter meaning.  Optional epilogue `epi`
 * (struct declared below; noise/bias/act/residual fields only): the StyledConv(upsample)
 * tail "blur -> NoiseInjection -> FusedLeakyReLU -> + feat + sty_de_feat"
 * (models/RestoreNet.py:599-603, :1031-1035) in the same pass.
 */
struct vsp_conv_epilogue;
int vsp_upfirdn2d_nhwc_bf16(const void *x, const float *filt, void *y,
                            int64_t n, int64_t in_h, int64_t in_w, int64_t c,
                            int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                            int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                            const struct vsp_conv_epilogue *epi, void *stream);

/*
 * Same blur for a SEPARABLE filter filt[j][i] = fy[j] * fx[i] (the model's outer([1,3,3,1]) kernels,
 * models/RestoreNet.py:32-40): up = down = 1, filter up to 4x4, 1-D taps passed from HOST memory.  For c % 64 == 0 and
 * in_w >= 35 this is the TMA strip kernel (32 output columns x 64 channels per CTA walking down the image, input rows as
 * TMA boxes in a 4-stage ring with out-of-bounds zero fill as the padding, rolling 4-row window, packed fp32 FMA:
 * 5.3-6.0 TB/s); other shapes take the register-tiled form (one horizontal pass per input row shared by two output
 * columns, then a vertical scatter).
 */
int vsp_blur_sep_nhwc_bf16(const void *x, const float *fy_host, const float *fx_host, void *y,
                           int64_t n, int64_t in_h, int64_t in_w, int64_t c, int kh, int kw,
                           int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                           const struct vsp_conv_epilogue *epi, void *stream);

/* ---- bias + activation ------------------------------------------------- */

/*
 * y[i] = act(x[i] + b[(i / step_b) % size_b]) * scale
 * Replaces `fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)`
 * of op/fused_bias_act.cpp:18-32 / op/fused_bias_act_kernel.cu:18-105.
 * `b` / `ref` may be NULL (the reference passes 0-element tensors,
 * op/fused_bias_act_kernel.cu:79-80).  act*10+grad selects the branch exactly
 * as op/fused_bias_act_kernel.cu:40-61: 30 lrelu fwd, 31 lrelu grad gated on
 * ref > 0, 32 / 12 zero, 10 / 11 linear.
 */
int vsp_bias_act_f32(const float *x, const float *b, const float *ref, float *y,
                     int64_t n, int64_t step_b, int64_t size_b,
                     int act, int grad, float alpha, float scale, void *stream);

/*
 * First-order backward of the leaky-ReLU bias-act with the bias-gradient
 * reduction fused in (the reference runs a separate ATen `sum`,
 * op/fused_act.py:139-145):
 *   dx[i]  = scale * (ref[i] > 0 ? dy[i] : alpha * dy[i])
 *   dbias[c] = sum over i with channel(i) == c of dx[i]      (may be NULL)
 * `dbias` [size_b] is overwritten (zeroed on `stream` inside the call).
 */
int vsp_bias_act_bwd_f32(const float *dy, const float *ref, float *dx, float *dbias,
                         int64_t n, int64_t step_b, int64_t size_b,
                         float alpha, float scale, void *stream);

/* ---- layout / modulation prologues ------------------------------------- */

/*
 * Output side of the inference loop (restoration_test.py:138-157: every restored / low / sample image goes through
 * torchvision.utils.save_image(..., normalize=True, range=(-1, 1))): quantise on the DEVICE, so the device->host copy
 * moves 3 bytes per pixel instead of 12 and the host only encodes.
 *   x [batch, channels, hw] fp32 (NCHW)  ->  y [batch, hw, channels] uint8 (HWC), channels <= 4
 *   y = trunc(clamp(((clamp(x, lo, hi) - lo) / max(hi - lo, 1e-5)) * 255 + 0.5, 0, 255))  — torchvision's arithmetic,
 *   operation by operation in round-to-nearest fp32, so the bytes equal the reference's.
 */
int vsp_quantize_nchw_f32_to_hwc_u8(const float *x, void *y, int64_t batch, int64_t channels, int64_t hw,
                                    float lo, float hi, void *stream);

/* NCHW fp32 -> NHWC bf16 (optionally scaled per (n, c): `scale_nc` may be NULL).
 * x [n, c, hw] -> y [n, hw, c_pad]; channels c..c_pad-1 are zero filled. */
int vsp_nchw_f32_to_nhwc_bf16(const float *x, const float *scale_nc, void *y,
                              int64_t n, int64_t c, int64_t hw, int64_t c_pad, void *stream);

/* NCHW fp32 -> NHWC bf16 TWO-TERM operand: y [n, hw, 3*c] = [hi | lo | hi] with v = x * scale_nc (may be NULL),
 * hi = bf16(v), lo = bf16(v - hi).  Convolved with weights laid out [w_hi | w_hi | w_lo] over 3*c input channels it gives
 * conv(x_hi,w_hi) + conv(x_lo,w_hi) + conv(x_hi,w_lo): near-fp32 products on the bf16 tensor pipe, used for the <= 32x32
 * encoder layers of Restoration_net whose output feeds every decoder style (models/RestoreNet.py:937-940).  c even. */
int vsp_nchw_f32_to_nhwc_split3_bf16(const float *x, const float *scale_nc, void *y,
                                     int64_t n, int64_t c, int64_t hw, void *stream);

/* NHWC bf16 -> NCHW fp32: x [n, hw, c_pad] -> y [n, c, hw]. */
int vsp_nhwc_bf16_to_nchw_f32(const void *x, float *y,
                              int64_t n, int64_t c, int64_t hw, int64_t c_pad, void *stream);

/*
 * The two conversions with the reductions of the modulated convolution's backward pass fused in (they ride on data the
 * conversion already holds; hw % 4 == 0 and 16-byte aligned tensors when `other` or, for the second, `scale_nc` is given).
 * The modulated convolution runs as y = d[b,o] * conv(x * s[b,i], wscale * W) (models/RestoreNet.py:481-508, the
 * reference's un-fused branch, algebraically its fused one :510-553), so no per-sample weights or weight gradients exist:
 *
 *   vsp_nchw_f32_to_nhwc_bf16_dot:  y[n,p,c] = bf16(x[n,c,p] * scale_nc[n,c]);  dot_nc[n,c] = sum_p x[n,c,p] * other[n,c,p]
 *       forward:  x = activation, scale_nc = s           (style modulation inside the layout conversion)
 *       backward: x = dy, scale_nc = d, other = y        (dz = d * dy; dot = sum_p dy * y gives dL/dd = dot / d)
 *   vsp_nhwc_bf16_to_nchw_f32_dot:  y[n,c,p] = x[n,p,c] * scale_nc[n,c];        dot_nc[n,c] = sum_p x[n,p,c] * other[n,c,p]
 *       backward: x = d(x*s) from the adjoint convolution, scale_nc = s, other = the forward activation
 *                 (dx = s * dxs; dot = ds, the style gradient, with no per-sample weight gradient)
 * `scale_nc`, and `other` together with `dot_nc`, may be NULL.  dot_nc [n, c] is overwritten (zeroed on `stream` first).
 */
int vsp_nchw_f32_to_nhwc_bf16_dot(const float *x, const float *scale_nc, const float *other, void *y, float *dot_nc,
                                  int64_t n, int64_t c, int64_t hw, int64_t c_pad, void *stream);
int vsp_nhwc_bf16_to_nchw_f32_dot(const void *x, const float *scale_nc, const float *other, float *y, float *dot_nc,
                                  int64_t n, int64_t c, int64_t hw, int64_t c_pad, void *stream);

/*
 * Weight gradient of a 1x1, stride-1, unpadded convolution with 1..8 input channels (NCHW fp32):
 *   gw[o, i] = sum_{b, p} dy[b, o, p] * x[b, i, p]        dy [batch, cout, hw], x [batch, cin, hw], hw % 4 == 0
 * Replaces aten::cudnn_convolution_backward_weight (op/conv2d_gradfix.py:180-199) for the RGB-side layers
 * (LargeConvLayer 3->16, models/RestoreNet.py:725-787; Discriminator stem 3->64, :1218), where the library's
 * tall-skinny GEMM takes milliseconds.  gw is zeroed by the call (stream-ordered) and accumulated with atomics.
 */
int vsp_conv1x1_wgrad_small_f32(const float *dy, const float *x, float *gw, int64_t batch, int64_t cout,
                                int64_t cin, int64_t hw, void *stream);

/*
 * Tail of an IR-SE residual unit of the e4e encoder (e4e/models/encoders/helpers.py:97-123), channels-last bf16, one pass:
 *   y[n,p,c] = res[n,p,c] * gate[n,c] + shortcut[n,p,c]          (SE scale + residual add)
 *   z[n,p,c] = bf16(y) * bn_a[c] + bn_b[c]                       (the next unit's leading eval BatchNorm; z may be NULL)
 * res / y / z [batch, h, w, c] dense NHWC bf16 (c % 8 == 0), gate [batch, c] fp32, bn_a / bn_b [c] fp32.  `shortcut` is
 * addressed as n*sc_n_stride + y*sc_h_stride + x*sc_w_stride + c (elements; multiples of 8), so the strided view that
 * MaxPool2d(1, 2) amounts to needs no copy.  First piece of SURVEY.md §8 f-2 (front end) on own kernels.
 */
int vsp_se_tail_nhwc_bf16(const void *res, const float *gate, const void *shortcut, void *y, void *z,
                          const float *bn_a, const float *bn_b, int64_t batch, int64_t h, int64_t w, int64_t c,
                          int64_t sc_n_stride, int64_t sc_h_stride, int64_t sc_w_stride, void *stream);

/* y[b,p,c] = bf16(x[b,p,c] * s[b,c]) on an NHWC bf16 activation [batch, hw, c] (c % 8 == 0, s [batch, c] fp32):
 * the input-modulated form of ModulatedConv2d (models/RestoreNet.py:481-508, `fused=False`), used where the
 * activation is smaller than the per-sample weights so the convolution can run on shared, cached weights. */
int vsp_scale_nhwc_bf16(const void *x, const float *s, void *y, int64_t batch, int64_t hw, int64_t c, void *stream);

/* NCHW fp32 -> NCHW bf16, optionally scaled per (n, c) plane. */
int vsp_nchw_f32_to_bf16(const float *x, const float *scale_nc, void *y,
                         int64_t planes, int64_t hw, void *stream);

/*
 * Weight prologue of the modulated convolution (models/RestoreNet.py:510-520,
 * e4e/models/stylegan2/model.py:237-254):
 *   m[b,o,i,t]  = wscale * W[o,i,t] * s[b,i]
 *   demod[b,o]  = rsqrt(sum_{i,t} m^2 + eps)             (if demod != NULL)
 *   wq[b,t',o',i'] = bf16(m * (fold_demod ? demod[b,o] : 1))
 * with the packed layout the tcgen05 kernel consumes ([group][tap][n][k],
 * k contiguous):
 *   transpose == 0 : n = o (Cout), k = i (Cin)   -> fprop
 *   transpose == 1 : n = i (Cin),  k = o (Cout)  -> dgrad (tap mirroring is expressed by the
 *                    tap offsets of vsp_conv2d_gather_bf16, the tap index t' = t is kept)
 * `s` may be NULL (plain EqualConv2d: s == 1, batch == 1 group shared by all
 * samples).  n is padded to n_pad rows and k to k_pad columns with zeros.
 * `wsq` (optional, [cout, cin] from vsp_weight_sumsq_f32) turns the demodulation sum into
 * sum_i s^2 * wsq — the host caches it while the weights are static (inference).
 */
int vsp_modulate_weights_bf16(const float *w, const float *s, float *demod, void *wq,
                              int64_t batch, int64_t cout, int64_t cin, int taps,
                              float wscale, float eps, int transpose, int fold_demod,
                              int64_t n_pad, int64_t k_pad, const float *wsq, void *stream);

/* wsq[o,i] = sum_t w[o,i,t]^2 (style-independent part of the demodulation sum). */
int vsp_weight_sumsq_f32(const float *w, float *wsq, int64_t cout, int64_t cin, int taps, void *stream);

/*
 * Grouped EqualLinear (models/RestoreNet.py:142-176 without activation): every modulation linear of a network
 * pass in one launch.  Problem j:  y[y_off_j + b*out_dim_j + o] = wscale_j * sum_i w_j[o,i] * x[x_off_j + b*x_bstride + i]
 *                                                               + bscale_j * bias_j[o]
 * `descs_dev` / `row_start_dev` live in device memory; row_start[j] = sum over problems < j of out_dim rounded up to
 * a multiple of 8 (a block of 8 warps works on 8 rows of one problem); total_rows = that sum over all problems.
 * w_j must be 16-byte aligned when in_dim % 4 == 0, and x_off_j a multiple of 4 floats.
 */
typedef struct vsp_linear_desc {
  const float *w;     /* [out_dim, in_dim] row-major */
  const float *bias;  /* [out_dim] or NULL */
  int64_t x_off;      /* element offset of this problem's style row inside sample 0 of x */
  int64_t y_off;      /* element offset of this problem's [batch, out_dim] block in y */
  int32_t in_dim, out_dim;
  float wscale, bscale;
  int64_t x_bstride;  /* elements between samples of this problem's input rows; 0 = the call's x_bstride */
  int32_t act;        /* 0 = none; 3 = leaky relu: y = lrelu(w.x * wscale + bias * bscale, alpha) * gain
                         (EqualLinear(activation="fused_lrelu"), models/RestoreNet.py:167-170: the style MLP) */
  float alpha, gain;
  int32_t pad_;
} vsp_linear_desc;
int vsp_grouped_linear_f32(const vsp_linear_desc *descs_dev, const int *row_start_dev, int n_problems,
                           int total_rows, const float *x, int64_t x_bstride, float *y, int batch, void *stream);

/* ---- tcgen05 implicit-GEMM convolution --------------------------------- */

/* Epilogue description shared by the conv entry points. All pointers optional.
 * Order of application:
 *   v = acc * row_scale[b,n]
 *   if pre_act: v = act(v + pre_bias[n]) * scale        (SMART_layer fusion conv + its FusedLeakyReLU)
 *   v = act(v + noise_weight * noise[b,pixel] + bias[n]) * scale   (NoiseInjection + FusedLeakyReLU; act==0: linear)
 *   v += residual + residual2 */
typedef struct vsp_conv_epilogue {
  const float *row_scale; /* [batch, cout]  demodulation coefficient d[b,o]            */
  const float *noise;     /* [batch or 1, full_h, full_w] noise image (NoiseInjection) */
  int64_t noise_bstride;  /* elements between samples of `noise` (0 = shared)          */
  float noise_weight;     /* NoiseInjection.weight (host scalar) ...                   */
  const float *noise_weight_dev; /* ... or, if non-NULL, a device scalar read by the kernel (no host sync) */
  const float *bias;      /* [cout] FusedLeakyReLU / ToRGB bias                        */
  const float *pre_bias;  /* [cout] bias of the first (pre-noise) activation stage     */
  int pre_act;            /* 0 = no first stage, 3 = leaky relu                        */
  int act;                /* 0 = none, 3 = leaky relu (as fused_bias_act)              */
  float alpha;            /* negative slope                                            */
  float scale;            /* output gain (sqrt 2)                                      */
  const void *residual;   /* same layout/dtype as the output; added after activation   */
  const void *residual2;  /* second residual (decoder skip fusion out + feat + feat2)  */
  const float *alpha_vec; /* [cout] per-channel negative slope (PReLU, e4e/models/encoders/helpers.py:76-123): when non-NULL
                             and act == 3 the activation is v > 0 ? v : alpha_vec[n] * v (times scale), `alpha` is ignored */
} vsp_conv_epilogue;

/*
 * Forward 2-D convolution as a bf16 implicit GEMM on tcgen05/TMEM fed by TMA:
 *   M = pixels of one sample tile, N = Cout, K = taps * Cin.
 * Replaces cuDNN via conv2d_gradfix.conv2d (op/conv2d_gradfix.py:22-42) for the
 * grouped form ModulatedConv2d issues (models/RestoreNet.py:547-553: weights
 * [B*Cout, Cin, k, k], groups = B) and for the plain EqualConv2d form
 * (groups = 1, one weight group shared by every sample).
 *
 *   x   [batch, in_h, in_w, cin]   bf16 NHWC (cin % 16 == 0)
 *   wq  [groups, kh*kw, cout_pad, cin] bf16 from vsp_modulate_weights_bf16;
 *       groups == batch (per-sample weights) or 1 (shared)
 *   out [batch, cout, out_h, out_w] fp32 NCHW           (out_nhwc_bf16 == 0)
 *       [batch, out_h, out_w, ldo] bf16 NHWC at channel offset `co_off`
 *                                                       (out_nhwc_bf16 == 1)
 * out_h = (in_h + 2*pad - dil*(kh-1) - 1)/stride + 1 (same for w).
 */
int vsp_conv2d_fprop_bf16(const void *x, const void *wq, void *out,
                          int64_t batch, int64_t groups, int64_t in_h, int64_t in_w,
                          int64_t cin, int64_t cout, int64_t cout_pad,
                          int kh, int kw, int stride, int pad, int dil,
                          int out_nhwc_bf16, int64_t ldo, int64_t co_off,
                          const vsp_conv_epilogue *epi, void *stream);

/*
 * General "gather convolution" on the same tcgen05 kernel:
 *   out[b, oh*os+oo_h, ow*os+oo_w, n] = sum_{t<ntaps} sum_c
 *        x[b, oh*stride + tap_dy[t], ow*stride + tap_dx[t], c] * wq[g, tap_w[t], n, c]
 * for oh < out_h, ow < out_w, written into an output whose spatial extent is
 * [full_h, full_w].  Expresses input gradients (transposed weights + mirrored tap
 * offsets: op/conv2d_gradfix.py:158-167) and the parity classes of stride-2 transposed
 * convolutions.  wq is [groups, taps_total, cout_pad, cin].
 */
int vsp_conv2d_gather_bf16(const void *x, const void *wq, void *out,
                           int64_t batch, int64_t groups, int64_t in_h, int64_t in_w,
                           int64_t cin, int64_t cout, int64_t cout_pad, int taps_total,
                           int ntaps, const int *tap_w, const int *tap_dy, const int *tap_dx,
                           int stride, int64_t out_h, int64_t out_w,
                           int out_nhwc_bf16, int64_t full_h, int64_t full_w,
                           int os, int oo_h, int oo_w, int64_t ldo, int64_t co_off,
                           const vsp_conv_epilogue *epi, void *stream);

/*
 * Stride-2 transposed convolution (padding 0) without zero-stuffing: four parity-class
 * launches of the gather kernel. Replaces conv2d_gradfix.conv_transpose2d(stride=2,
 * padding=0, groups=B) of models/RestoreNet.py:522-535.
 *   x [batch, in_h, in_w, cin] bf16 NHWC -> out extent ((in_h-1)*2 + kh, (in_w-1)*2 + kw)
 *   wq [groups, kh*kw, cout_pad, cin] (n = Cout, k = Cin, tap = kh_i*kw + kw_i, not flipped)
 */
int vsp_conv_transpose2d_s2_bf16(const void *x, const void *wq, void *out,
                                 int64_t batch, int64_t groups, int64_t in_h, int64_t in_w,
                                 int64_t cin, int64_t cout, int64_t cout_pad, int kh, int kw,
                                 int out_nhwc_bf16, int64_t ldo, int64_t co_off,
                                 const vsp_conv_epilogue *epi, void *stream);

/*
 * Fused up-convolution: stride-2 transposed 3x3 convolution followed by the 4x4 FIR blur (pad 1,1), i.e.
 * ModulatedConv2d(upsample=True) of models/RestoreNet.py:522-535 (conv_transpose2d -> Blur), as ONE dense 3x3
 * convolution on the low-resolution grid with 4*cout GEMM columns (the composite 6x6 stride-2 kernel split into
 * its four output-parity classes) and a pixel-shuffle epilogue: column n = (pa*2+pb)*cout + o is written to
 * out[b, 2*oh+pa, 2*ow+pb, co_off+o].  No (2H+1)^2 intermediate, no separate blur pass.
 *   x   [batch, in_h, in_w, cin] bf16 NHWC
 *   wq  [groups, 9, 4*cout, cin] bf16: vsp_modulate_weights_bf16 applied to the class-split composite weights
 *   out [batch, 2*in_h, 2*in_w, ldo] bf16 NHWC (ldo % 8 == 0); epilogue vectors are indexed by the real channel o
 */
int vsp_conv2d_up2_fused_bf16(const void *x, const void *wq, void *out,
                              int64_t batch, int64_t groups, int64_t in_h, int64_t in_w,
                              int64_t cin, int64_t cout, int64_t ldo, int64_t co_off,
                              const vsp_conv_epilogue *epi, void *stream);

/*
 * The same up-sampling layer (conv_transpose2d stride 2, 3x3 -> Blur 4x4 pad 1, models/RestoreNet.py:522-535) at HALF the
 * dense form's tensor work, for the wide levels (in_w >= 128; cin = 64 or 128; cout % 32 == 0; in_w % 32 == 0): only the
 * HORIZONTAL half of the separable blur is composed into the weights, the vertical 4-tap half is applied in the epilogue
 * to fp32 accumulators while the CTA walks down a 128-column strip (conv_up2h_sm100.cu).
 *   wq  [groups, 9, 2*cout, cin] bf16: vsp_modulate_weights_bf16 applied to the horizontally composed weights
 *       Wc[q*cout + o, i, kh, dx] = sum_{v, kw : q + v - 1 - kw = 2 (dx - 1)} fx[3 - v] * W[o, i, kh, kw]
 *       (W = the transposed convolution's [cout, cin, 3, 3] weights, fx = horizontal factor of the blur filter)
 *   ky_host [4] (HOST memory): vertical factor of the blur as applied, out[y] = sum_u ky[u] * hz[y + u - 1]
 *       (= the flipped vertical factor; the model's filters are symmetric)
 *   out [batch, 2*in_h, 2*in_w, ldo] bf16 NHWC; epilogue as vsp_conv2d_up2_fused_bf16 (no first activation stage)
 */
int vsp_conv2d_up2h_bf16(const void *x, const void *wq, void *out,
                         int64_t batch, int64_t groups, int64_t in_h, int64_t in_w,
                         int64_t cin, int64_t cout, int64_t ldo, int64_t co_off,
                         const float *ky_host, const vsp_conv_epilogue *epi, void *stream);

/*
 * The four dilated branches of a SMART_layer (models/RestoreNet.py:196-209,229-233: Dilated_ModulatedConv2d x4 ->
 * torch.cat) as ONE launch: branch j is a 3x3 stride-1 convolution with dilation = padding = dils[j] producing
 * channels [j*cout/n, (j+1)*cout/n) of the output (channel tile = branch; no torch.cat, one wave of tiles).
 *   wq [groups, 9, cout, cin]: the branches' weights concatenated along the output-channel axis.
 */
int vsp_conv2d_branches_bf16(const void *x, const void *wq, void *out,
                             int64_t batch, int64_t groups, int64_t in_h, int64_t in_w,
                             int64_t cin, int64_t cout, int n_branches, const int *dils,
                             int out_nhwc_bf16, int64_t ldo, int64_t co_off,
                             const vsp_conv_epilogue *epi, void *stream);

/*
 * ToRGB (models/RestoreNet.py:647-666): 1x1 modulated convolution to 3 channels, no demodulation,
 *   out[b,o,p] = sum_c x[b,p,c] * wscale * w[o,c] * s[b,c] + bias[o] + skip[b,o,p]
 * x [batch, hw, c] bf16 NHWC (c % 8 == 0), w [3, c], s [batch, c] (NULL = 1), bias [3] / skip
 * [batch, 3, hw] fp32 optional, out [batch, 3, hw] fp32 NCHW.  Memory-bound SIMT kernel.
 */
int vsp_torgb_nhwc_bf16(const void *x, const float *w, const float *s, const float *bias,
                        const float *skip, float *out, int64_t batch, int64_t hw, int64_t c,
                        float wscale, void *stream);

/*
 * Last ToRGB of the style decoder fused with the face_pool that follows it (e4e/models/psp.py:245-246):
 *   out = AvgPool2x2( conv1x1_mod(x) + bias + Upsample(skip) )      out [batch, 3, out_h, out_w], x [batch, 2*out_h, 2*out_w, c]
 * skip [batch, 3, out_h, out_w] is the previous level's RGB image (same resolution as `out`); k3_host[9] (HOST memory) is
 * the 3x3 composite of the 2x FIR upsample followed by the 2x2 mean (outer([1/8, 3/4, 1/8]) for the model's filter).
 * k3_host == NULL: `skip` has already been passed through that composite (e.g. vsp_upfirdn2d_f32 with the flipped 3x3 taps,
 * pad 1) and is added as is — the form the lane-split kernel (C = 16/32/64/128, 512 contiguous bytes per warp load) takes.
 */
int vsp_torgb_pool2_nhwc_bf16(const void *x, const float *w, const float *s, const float *bias,
                              const float *skip, const float *k3_host, float *out,
                              int64_t batch, int64_t out_h, int64_t out_w, int64_t c, float wscale, void *stream);

/*
 * Weight gradient as a bf16 GEMM on tcgen05 (K = pixels):
 *   gw[g, t, o, i] = sum_{p in sample(s) of group g} dy[b,p,o] * x[b, p*stride + t*dil - pad, i]   (pad / dil per axis)
 * Replaces aten::cudnn_convolution_backward_weight, op/conv2d_gradfix.py:177-199.
 *   dy [batch, out_h, out_w, cout] bf16 NHWC, x [batch, in_h, in_w, cin] bf16 NHWC
 *   (cin, cout multiples of 8); both operands are consumed channel-contiguous (MN-major UMMA).
 *   gw [groups, kh*kw, cout, cin] fp32 (TAP-major); groups == batch (per-sample) or 1
 *   (summed over the batch).
 */
int vsp_conv2d_wgrad_bf16(const void *dy, const void *x, float *gw,
                          int64_t batch, int64_t groups, int64_t in_h, int64_t in_w,
                          int64_t cin, int64_t cout, int64_t out_h, int64_t out_w,
                          int kh, int kw, int stride, int pad_h, int pad_w, int dil_h, int dil_w,
                          void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VSP_B200_H_ */
