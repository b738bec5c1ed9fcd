/*
 * Plain-C restatement of the two memory-bound operators of the VSPBFR hot path — TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use the oracle).
 *
 *  oracle_upfirdn2d_f32 : semantics of the reference's generic CUDA kernel `upfirdn2d_kernel_large`
 *                         (/root/reference/op/upfirdn2d_kernel.cu:49-105) == its CPU path `upfirdn2d_native`
 *                         (/root/reference/op/upfirdn2d.py:365-406): zero-stuff by `up`, pad (negative = crop), TRUE
 *                         convolution with the filter (correlation with the flipped filter), keep every `down`-th sample.
 *                         Taps are visited y-outer, x-inner with sequential fp32 accumulation, as :83-98 does.
 *  oracle_bias_act_f32  : the `act * 10 + grad` switch of /root/reference/op/fused_bias_act_kernel.cu:40-61
 *                         (act 1 = linear, 3 = leaky ReLU; grad 0 = forward, 1 = first derivative gated on `ref`,
 *                         2 = second derivative = 0), bias broadcast as (i / step_b) % size_b (:33-36).
 *
 * Built by oracle/build_c.py (gcc -O2 -shared -fPIC) into oracle/_build/, pinned against the reference-generated
 * goldens in tests/test_oracle_golden.py next to the numpy restatement.
 */
#include <stdint.h>
#include <stddef.h>

static int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

int64_t oracle_upfirdn2d_out_size(int64_t n_in, int k, int up, int down, int pad0, int pad1) {
  /* op/upfirdn2d.py:301-302 (python floor division) */
  return floor_div(n_in * up + pad0 + pad1 - k + down, down);
}

/* x [major, in_h, in_w], filt [kh, kw] -> y [major, out_h, out_w]; returns 0, or 1 on a bad argument. */
int oracle_upfirdn2d_f32(const float *x, const float *filt, float *y, int64_t major, int64_t in_h, int64_t in_w,
                         int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                         int pad_y0, int pad_y1) {
  if (up_x < 1 || up_y < 1 || down_x < 1 || down_y < 1 || kh < 1 || kw < 1) return 1;
  const int64_t out_h = oracle_upfirdn2d_out_size(in_h, kh, up_y, down_y, pad_y0, pad_y1);
  const int64_t out_w = oracle_upfirdn2d_out_size(in_w, kw, up_x, down_x, pad_x0, pad_x1);
  if (out_h <= 0 || out_w <= 0) return 0;
  for (int64_t m = 0; m < major; ++m) {
    const float *xp = x + m * in_h * in_w;
    float *yp = y + m * out_h * out_w;
    for (int64_t oy = 0; oy < out_h; ++oy)
      for (int64_t ox = 0; ox < out_w; ++ox) {
        float acc = 0.f;
        for (int jy = 0; jy < kh; ++jy) {
          const int64_t uy = oy * down_y - pad_y0 + jy;          /* row in the zero-stuffed, padded image */
          if (uy < 0 || uy % up_y != 0) continue;
          const int64_t iy = uy / up_y;
          if (iy >= in_h) continue;
          for (int jx = 0; jx < kw; ++jx) {
            const int64_t ux = ox * down_x - pad_x0 + jx;
            if (ux < 0 || ux % up_x != 0) continue;
            const int64_t ix = ux / up_x;
            if (ix >= in_w) continue;
            /* true convolution: window tap (jy, jx) meets the flipped filter entry */
            acc += xp[iy * in_w + ix] * filt[(kh - 1 - jy) * kw + (kw - 1 - jx)];
          }
        }
        yp[oy * out_w + ox] = acc;
      }
  }
  return 0;
}

/* y[i] = f(x[i] + bias[(i / step_b) % size_b]) * scale with f chosen by act * 10 + grad; bias / ref may be NULL. */
int oracle_bias_act_f32(const float *x, const float *bias, const float *ref, float *y, int64_t n, int64_t step_b,
                        int64_t size_b, int act, int grad, float alpha, float scale) {
  if (n < 0 || (bias != NULL && (step_b < 1 || size_b < 1))) return 1;
  for (int64_t i = 0; i < n; ++i) {
    float v = x[i];
    if (bias != NULL) v += bias[(i / step_b) % size_b];
    const float r = ref != NULL ? ref[i] : 0.f;
    float o;
    switch (act * 10 + grad) {
      default:
      case 10: o = v; break;                                  /* linear */
      case 11: o = v; break;
      case 12: o = 0.f; break;
      case 30: o = (v > 0.f) ? v : v * alpha; break;          /* leaky ReLU */
      case 31: o = (r > 0.f) ? v : v * alpha; break;          /* d/dx, gated on the saved output */
      case 32: o = 0.f; break;
    }
    y[i] = o * scale;
  }
  return 0;
}
