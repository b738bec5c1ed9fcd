// layout_sm100.cu — layout conversions at the reference-visible boundary and the
// weight prologue / weight+style gradient epilogue of the modulated convolution.
//
// The reference is NCHW fp32 everywhere; the tcgen05 kernels want
// NHWC bf16 activations (K = channels contiguous) and [group][tap][n][k] bf16
// weights.  Modulation maths follows models/RestoreNet.py:510-520.
#include "common.cuh"

namespace vsp {
namespace {

constexpr int kThreads = 256;

// ---- NCHW fp32 -> NHWC bf16 (64 channels x 64 pixels per block, smem transpose)
__global__ void __launch_bounds__(kThreads)
nchw_to_nhwc_kernel(const float *__restrict__ x, const float *__restrict__ scale_nc,
                    __nv_bfloat16 *__restrict__ y, long long c, long long hw, long long c_pad) {
  __shared__ float tile[64][65];
  const long long n = blockIdx.z;
  const long long c0 = (long long)blockIdx.y * 64, p0 = (long long)blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int cl = ty + 4 * i;
    const long long cc = c0 + cl, pp = p0 + tx;
    float v = 0.f;
    if (cc < c && pp < hw) {
      v = ld_stream_f1(x + (n * c + cc) * hw + pp);
      if (scale_nc != nullptr) v *= __ldg(scale_nc + n * c + cc);
    }
    tile[cl][tx] = v;
  }
  __syncthreads();
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;  // 32 channel pairs x 8 pixels
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pl = py + 8 * i;
    const long long pp = p0 + pl, cc = c0 + 2 * cx;
    if (pp < hw && cc < c_pad) {
      __nv_bfloat162 o = __floats2bfloat162_rn(tile[2 * cx][pl], tile[2 * cx + 1][pl]);
      *reinterpret_cast<__nv_bfloat162 *>(y + (n * hw + pp) * c_pad + cc) = o;
    }
  }
}

// ---- NCHW fp32 -> NHWC bf16 two-term operand [hi | lo | hi] over 3*C channels: hi = bf16(v), lo = bf16(v - hi),
// v = x * scale.  A convolution of it with the weights [w_hi | w_hi | w_lo] is conv(x_hi, w_hi) + conv(x_lo, w_hi) +
// conv(x_hi, w_lo) ~ the fp32 product to 2^-17 — used for the few low-resolution encoder layers whose rounding error
// reaches every decoder modulation through x_global (fastpath.smart_layer_split).
__global__ void __launch_bounds__(kThreads)
nchw_to_nhwc_split3_kernel(const float *__restrict__ x, const float *__restrict__ scale_nc,
                           __nv_bfloat16 *__restrict__ y, long long c, long long hw) {
  __shared__ float tile[64][65];
  const long long n = blockIdx.z;
  const long long c0 = (long long)blockIdx.y * 64, p0 = (long long)blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int cl = ty + 4 * i;
    const long long cc = c0 + cl, pp = p0 + tx;
    float v = 0.f;
    if (cc < c && pp < hw) {
      v = ld_stream_f1(x + (n * c + cc) * hw + pp);
      if (scale_nc != nullptr) v *= __ldg(scale_nc + n * c + cc);
    }
    tile[cl][tx] = v;
  }
  __syncthreads();
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pl = py + 8 * i;
    const long long pp = p0 + pl, cc = c0 + 2 * cx;
    if (pp < hw && cc < c) {
      const float a = tile[2 * cx][pl], b = tile[2 * cx + 1][pl];
      const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
      const float2 hf = __bfloat1622float2(hi);
      const __nv_bfloat162 lo = __floats2bfloat162_rn(a - hf.x, b - hf.y);
      __nv_bfloat16 *row = y + (n * hw + pp) * 3 * c + cc;
      *reinterpret_cast<__nv_bfloat162 *>(row) = hi;
      *reinterpret_cast<__nv_bfloat162 *>(row + c) = lo;
      *reinterpret_cast<__nv_bfloat162 *>(row + 2 * c) = hi;
    }
  }
}

// ---- NCHW fp32 -> NHWC bf16 for c_pad == 8 (the RGB network input): one thread per pixel, coalesced plane reads,
// one 128-bit store; the 64 x 64 transpose tile above would be 95 % padding.
__global__ void __launch_bounds__(kThreads)
nchw_to_nhwc8_kernel(const float *__restrict__ x, const float *__restrict__ scale_nc, uint4 *__restrict__ y,
                     int c, long long hw, long long total) {
  for (long long idx = blockIdx.x * (long long)kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
    const long long n = idx / hw, pp = idx % hw;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = 0.f;
      if (i < c) {
        v[i] = ld_stream_f1(x + (n * c + i) * hw + pp);
        if (scale_nc != nullptr) v[i] *= __ldg(scale_nc + n * c + i);
      }
    }
    uint4 o;
    __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    y[idx] = o;
  }
}

// ---- NHWC bf16 -> NCHW fp32
__global__ void __launch_bounds__(kThreads)
nhwc_to_nchw_kernel(const __nv_bfloat16 *__restrict__ x, float *__restrict__ y, long long c,
                    long long hw, long long c_pad) {
  __shared__ float tile[64][65];  // [pixel][channel]
  const long long n = blockIdx.z;
  const long long c0 = (long long)blockIdx.y * 64, p0 = (long long)blockIdx.x * 64;
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int pl = py + 8 * i;
    const long long pp = p0 + pl, cc = c0 + 2 * cx;
    float2 v = make_float2(0.f, 0.f);
    if (pp < hw && cc < c_pad)
      v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(x + (n * hw + pp) * c_pad + cc));
    tile[pl][2 * cx] = v.x;
    tile[pl][2 * cx + 1] = v.y;
  }
  __syncthreads();
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int cl = ty + 4 * i;
    const long long cc = c0 + cl, pp = p0 + tx;
    if (cc < c && pp < hw) st_stream_f1(y + (n * c + cc) * hw + pp, tile[tx][cl]);
  }
}

// ---- vectorised forms of the two conversions above (hw % 4 == 0, 16-byte aligned tensors): a block moves a tile of
// 64 channels x 128 pixels.  The NCHW side is touched with 128-bit accesses (a warp covers 512 contiguous bytes of one
// plane, 8 loads per thread in flight), the NHWC side with 32 contiguous bytes per lane (two 8-channel vectors = one full
// sector); the shared tile [channel][pixel] has a pitch of 132 floats, which makes the 128-bit accesses of the NCHW side and
// the pixel-per-lane accesses of the NHWC side both conflict free.  Optional fused reductions for the backward passes of
// the modulated convolution (they ride on data the conversion already has in registers):
//   DOT: dot_nc[n,c] += sum_p x[n,c,p] * other[n,c,p]     (`other` in the NCHW fp32 layout; x UNSCALED)
constexpr int kTileC = 64, kTileP = 128, kPitch = 132;

template <bool DOT>
__global__ void __launch_bounds__(kThreads)
nchw_to_nhwc_v4_kernel(const float *__restrict__ x, const float *__restrict__ scale_nc, const float *__restrict__ other,
                       __nv_bfloat16 *__restrict__ y, float *__restrict__ dot_nc, long long c, long long hw,
                       long long c_pad) {
  __shared__ __align__(16) float tile[kTileC * kPitch];
  const long long n = blockIdx.z;
  const long long c0 = (long long)blockIdx.y * kTileC, p0 = (long long)blockIdx.x * kTileP;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  {
    const long long pp = p0 + 4 * lane;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long cc = c0 + wrp + 8 * i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cc < c && pp < hw) v[i] = ld_stream_f4(reinterpret_cast<const float4 *>(x + (n * c + cc) * hw + pp));
    }
    if (DOT) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long cc = c0 + wrp + 8 * i;
        float acc = 0.f;
        if (cc < c && pp < hw) {
          const float4 o = ld_stream_f4(reinterpret_cast<const float4 *>(other + (n * c + cc) * hw + pp));
          acc = fmaf(v[i].x, o.x, fmaf(v[i].y, o.y, fmaf(v[i].z, o.z, v[i].w * o.w)));
        }
        acc = warp_sum(acc);
        if (lane == 0 && cc < c) atomicAdd(dot_nc + n * c + cc, acc);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int cl = wrp + 8 * i;
      float4 o = v[i];
      if (scale_nc != nullptr && c0 + cl < c) {
        const float sc = __ldg(scale_nc + n * c + c0 + cl);
        o.x *= sc; o.y *= sc; o.z *= sc; o.w *= sc;
      }
      *reinterpret_cast<float4 *>(&tile[cl * kPitch + 4 * lane]) = o;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int pl = threadIdx.x & (kTileP - 1), q = (threadIdx.x >> 7) + 2 * i;   // pixel, 16-channel group
    const long long pp = p0 + pl, cc = c0 + 16 * q;
    if (pp >= hw || cc >= c_pad) continue;
    uint4 o[2];
    __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(o);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      oh[j] = __floats2bfloat162_rn(tile[(16 * q + 2 * j) * kPitch + pl], tile[(16 * q + 2 * j + 1) * kPitch + pl]);
    uint4 *dst = reinterpret_cast<uint4 *>(y + (n * hw + pp) * c_pad + cc);
    dst[0] = o[0];
    if (cc + 8 < c_pad) dst[1] = o[1];
  }
}

//   y[n,c,p] = x[n,p,c] * scale_nc[n,c]   and, when DOT,   dot_nc[n,c] += sum_p x[n,p,c] * other[n,c,p]   (x UNSCALED)
template <bool DOT>
__global__ void __launch_bounds__(kThreads)
nhwc_to_nchw_v4_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ scale_nc,
                       const float *__restrict__ other, float *__restrict__ y, float *__restrict__ dot_nc, long long c,
                       long long hw, long long c_pad) {
  __shared__ __align__(16) float tile[kTileC * kPitch];
  const long long n = blockIdx.z;
  const long long c0 = (long long)blockIdx.y * kTileC, p0 = (long long)blockIdx.x * kTileP;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int pl = threadIdx.x & (kTileP - 1), q = (threadIdx.x >> 7) + 2 * i;
    const long long pp = p0 + pl, cc = c0 + 16 * q;
    uint4 v[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (pp < hw && cc < c_pad) {
      const uint4 *src = reinterpret_cast<const uint4 *>(x + (n * hw + pp) * c_pad + cc);
      v[0] = __ldg(src);
      if (cc + 8 < c_pad) v[1] = __ldg(src + 1);
    }
    const __nv_bfloat162 *vh = reinterpret_cast<const __nv_bfloat162 *>(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = __bfloat1622float2(vh[j]);
      tile[(16 * q + 2 * j) * kPitch + pl] = f.x;
      tile[(16 * q + 2 * j + 1) * kPitch + pl] = f.y;
    }
  }
  __syncthreads();
  const long long pp = p0 + 4 * lane;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int cl = wrp + 8 * i;
    const long long cc = c0 + cl;
    if (cc >= c) continue;                                // warp-uniform
    float acc = 0.f;
    if (pp < hw) {
      float4 v = *reinterpret_cast<const float4 *>(&tile[cl * kPitch + 4 * lane]);
      if (DOT) {
        const float4 o = ld_stream_f4(reinterpret_cast<const float4 *>(other + (n * c + cc) * hw + pp));
        acc = fmaf(v.x, o.x, fmaf(v.y, o.y, fmaf(v.z, o.z, v.w * o.w)));
      }
      if (scale_nc != nullptr) {
        const float sc = __ldg(scale_nc + n * c + cc);
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      }
      st_stream_f4(reinterpret_cast<float4 *>(y + (n * c + cc) * hw + pp), v);
    }
    if (DOT) {
      acc = warp_sum(acc);
      if (lane == 0) atomicAdd(dot_nc + n * c + cc, acc);
    }
  }
}

// ---- NCHW fp32 -> NCHW bf16 with a per-plane scale
__global__ void __launch_bounds__(kThreads)
nchw_cast_kernel(const float *__restrict__ x, const float *__restrict__ scale_nc,
                 __nv_bfloat16 *__restrict__ y, long long hw, int vec) {
  const long long plane = blockIdx.y;
  const float s = scale_nc ? __ldg(scale_nc + plane) : 1.f;
  const float *xp = x + plane * hw;
  __nv_bfloat16 *yp = y + plane * hw;
  const long long stride = (long long)gridDim.x * kThreads;
  if (vec) {
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < hw / 4; i += stride) {
      const float4 v = ld_stream_f4(reinterpret_cast<const float4 *>(xp) + i);
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x * s, v.y * s), b = __floats2bfloat162_rn(v.z * s, v.w * s);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t *>(&a);
      o.y = *reinterpret_cast<uint32_t *>(&b);
      *(reinterpret_cast<uint2 *>(yp) + i) = o;
    }
  } else {
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < hw; i += stride)
      yp[i] = __float2bfloat16_rn(xp[i] * s);
  }
}

// ---- demod[b,o] = rsqrt(sum_{i,t} (wscale*W[o,i,t]*s[b,i])^2 + eps): one warp per (b,o)
__global__ void __launch_bounds__(kThreads)
demod_kernel(const float *__restrict__ w, const float *__restrict__ s, float *__restrict__ demod,
             long long batch, long long cout, long long cin, int taps, float wscale, float eps) {
  const long long warp = (blockIdx.x * (long long)kThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch * cout) return;
  const long long b = warp / cout, o = warp % cout;
  const float *wr = w + o * cin * taps;
  float acc = 0.f;
  for (long long e = lane; e < cin * taps; e += 32) {
    const long long i = e / taps;
    const float m = wscale * wr[e] * (s ? __ldg(s + b * cin + i) : 1.f);
    acc = fmaf(m, m, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) demod[warp] = rsqrtf(acc + eps);
}

// ---- wsq[o,i] = sum_t W[o,i,t]^2 (style independent: cached by the host while weights are static)
__global__ void __launch_bounds__(kThreads)
weight_sumsq_kernel(const float *__restrict__ w, float *__restrict__ wsq, long long n, int taps) {
  const long long e = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (e >= n) return;
  float acc = 0.f;
  for (int t = 0; t < taps; ++t) {
    const float v = w[e * taps + t];
    acc = fmaf(v, v, acc);
  }
  wsq[e] = acc;
}

// ---- y[b,p,c] = bf16(x[b,p,c] * s[b,c]): per-sample channel scaling of an NHWC bf16 activation (the
// "modulate the input instead of the weights" form of ModulatedConv2d, models/RestoreNet.py:481-508)
__global__ void __launch_bounds__(kThreads)
scale_nhwc_kernel(const uint4 *__restrict__ x, const float *__restrict__ s, uint4 *__restrict__ y, long long hw,
                  int cg, long long s_bstride, long long total) {
  for (long long idx = blockIdx.x * (long long)kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
    const int g = (int)(idx % cg);
    const long long b = idx / ((long long)cg * hw);
    const uint4 v = __ldg(x + idx);
    const float4 s0 = __ldg(reinterpret_cast<const float4 *>(s + b * s_bstride + g * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4 *>(s + b * s_bstride + g * 8 + 4));
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
    uint4 o;
    __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      oh[i] = __floats2bfloat162_rn(f.x * sv[2 * i], f.y * sv[2 * i + 1]);
    }
    y[idx] = o;
  }
}

// ---- demod[b,o] = rsqrt(wscale^2 * sum_i s[b,i]^2 * wsq[o,i] + eps): one warp per (b,o)
__global__ void __launch_bounds__(kThreads)
demod_from_wsq_kernel(const float *__restrict__ wsq, const float *__restrict__ s, float *__restrict__ demod,
                      long long batch, long long cout, long long cin, float wscale, float eps) {
  const long long warp = (blockIdx.x * (long long)kThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch * cout) return;
  const long long b = warp / cout, o = warp % cout;
  float acc = 0.f;
  for (long long i = lane; i < cin; i += 32) {
    const float sv = s ? __ldg(s + b * cin + i) : 1.f;
    acc = fmaf(sv * sv, __ldg(wsq + o * cin + i), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) demod[warp] = rsqrtf(wscale * wscale * acc + eps);
}

// ---- pack modulated weights to bf16 [b][tap'][n_pad][k_pad]; one thread per (n', k')
__global__ void __launch_bounds__(kThreads)
pack_weights_kernel(const float *__restrict__ w, const float *__restrict__ s,
                    const float *__restrict__ demod, __nv_bfloat16 *__restrict__ wq, long long cout,
                    long long cin, int taps, float wscale, int transpose, long long n_pad,
                    long long k_pad) {
  const long long b = blockIdx.y;
  const long long idx = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (idx >= n_pad * k_pad) return;
  const long long nn = idx / k_pad, kk = idx % k_pad;
  const long long o = transpose ? kk : nn, i = transpose ? nn : kk;
  const bool valid = o < cout && i < cin;
  float f = 0.f;
  if (valid) {
    f = wscale * (s ? __ldg(s + b * cin + i) : 1.f);
    if (demod) f *= __ldg(demod + b * cout + o);
  }
  const float *wr = w + (o * cin + i) * taps;
  __nv_bfloat16 *dst = wq + b * taps * n_pad * k_pad + idx;
  for (int t = 0; t < taps; ++t) {
    dst[(long long)t * n_pad * k_pad] = __float2bfloat16_rn(valid ? wr[t] * f : 0.f);
  }
}

// Vectorised form for the inference prologue (no transpose, cin % 8 == 0): one thread owns 8 consecutive input channels
// of one output channel for ALL taps — 8*TAPS contiguous fp32 weights, read once with 128-bit loads and reused for kBG
// samples — and writes one 16-byte bf16 vector per tap and sample (a warp stores 512 contiguous bytes; the scalar kernel
// above issues 2-byte stores: 1.4 TB/s).
constexpr int kBG = 4;
template <int TAPS>
__global__ void __launch_bounds__(kThreads)
pack_weights_vec_kernel(const float *__restrict__ w, const float *__restrict__ s, const float *__restrict__ demod,
                        uint4 *__restrict__ wq, long long cout, long long cin, float wscale, long long n_pad,
                        long long batch) {
  const long long kg = cin / 8;
  const long long idx = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (idx >= n_pad * kg) return;
  const long long o = idx / kg, i0 = (idx % kg) * 8;
  const bool valid = o < cout;
  float wv[8 * TAPS];
  if (valid) {
    const float4 *src = reinterpret_cast<const float4 *>(w + (o * cin + i0) * TAPS);
#pragma unroll
    for (int q = 0; q < 2 * TAPS; ++q) {
      const float4 v = __ldg(src + q);
      wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8 * TAPS; ++q) wv[q] = 0.f;
  }
  const long long b0 = (long long)blockIdx.y * kBG;
#pragma unroll
  for (int bb = 0; bb < kBG; ++bb) {
    const long long b = b0 + bb;
    if (b >= batch) break;
    float f[8];
    const float d = (valid && demod) ? __ldg(demod + b * cout + o) : 1.f;
    if (s) {
      const float4 s0 = __ldg(reinterpret_cast<const float4 *>(s + b * cin + i0));
      const float4 s1 = __ldg(reinterpret_cast<const float4 *>(s + b * cin + i0) + 1);
      f[0] = s0.x; f[1] = s0.y; f[2] = s0.z; f[3] = s0.w; f[4] = s1.x; f[5] = s1.y; f[6] = s1.z; f[7] = s1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = 1.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= wscale * d;
    uint4 *dst = wq + (b * TAPS * n_pad + o) * kg + i0 / 8;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
      uint4 ov;
      __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&ov);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        oh[e] = __floats2bfloat162_rn(wv[(2 * e) * TAPS + t] * f[2 * e], wv[(2 * e + 1) * TAPS + t] * f[2 * e + 1]);
      dst[(long long)t * n_pad * kg] = ov;
    }
  }
}

}  // namespace
}  // namespace vsp

using namespace vsp;

namespace vsp {
namespace {
inline bool vec4_ok(const void *a, const void *b, const void *c, long long hw, long long c_pad) {
  return hw % 4 == 0 && c_pad % 8 == 0 &&
         ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}
}  // namespace
}  // namespace vsp

extern "C" int vsp_nchw_f32_to_nhwc_bf16_dot(const float *x, const float *scale_nc, const float *other, void *y,
                                             float *dot_nc, int64_t n, int64_t c, int64_t hw, int64_t c_pad,
                                             void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(n >= 0 && c >= 0 && hw >= 0 && c_pad >= c && c_pad % 2 == 0, "nchw->nhwc: bad geometry");
  VSP_REQUIRE((other == nullptr) == (dot_nc == nullptr), "nchw->nhwc: `other` and `dot_nc` go together");
  if (dot_nc != nullptr && n * c > 0) VSP_CUDA(cudaMemsetAsync(dot_nc, 0, sizeof(float) * n * c, stream));
  if (n == 0 || c_pad == 0 || hw == 0) return 0;
  VSP_REQUIRE(x && y, "nchw->nhwc: null pointer");
  VSP_REQUIRE(n <= 65535 && ceil_div64(c_pad, 64) <= 65535, "nchw->nhwc: batch/channel extent too large");
  if (dot_nc == nullptr && c_pad == 8 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    const long long total = n * hw;
    long long nb = ceil_div64(total, kThreads);
    if (nb > (long long)num_sms() * 32) nb = (long long)num_sms() * 32;
    nchw_to_nhwc8_kernel<<<(unsigned)nb, kThreads, 0, stream>>>(x, scale_nc, static_cast<uint4 *>(y), (int)c, hw, total);
    return check_launch("nchw_to_nhwc8_kernel");
  }
  if (vec4_ok(x, y, other, hw, c_pad)) {
    dim3 grid((unsigned)ceil_div64(hw, kTileP), (unsigned)ceil_div64(c_pad, kTileC), (unsigned)n);
    if (dot_nc != nullptr)
      nchw_to_nhwc_v4_kernel<true><<<grid, kThreads, 0, stream>>>(x, scale_nc, other, static_cast<__nv_bfloat16 *>(y),
                                                                   dot_nc, c, hw, c_pad);
    else
      nchw_to_nhwc_v4_kernel<false><<<grid, kThreads, 0, stream>>>(x, scale_nc, nullptr, static_cast<__nv_bfloat16 *>(y),
                                                                    nullptr, c, hw, c_pad);
    return check_launch("nchw_to_nhwc_v4_kernel");
  }
  VSP_REQUIRE(dot_nc == nullptr, "nchw->nhwc: the fused dot needs hw %% 4 == 0 and 16-byte aligned tensors");
  dim3 grid((unsigned)ceil_div64(hw, 64), (unsigned)ceil_div64(c_pad, 64), (unsigned)n);
  nchw_to_nhwc_kernel<<<grid, kThreads, 0, stream>>>(x, scale_nc, static_cast<__nv_bfloat16 *>(y), c, hw, c_pad);
  return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int vsp_nchw_f32_to_nhwc_split3_bf16(const float *x, const float *scale_nc, void *y, int64_t n, int64_t c,
                                                int64_t hw, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(n >= 0 && c >= 0 && hw >= 0 && c % 2 == 0, "nchw->nhwc split3: channels must be even");
  if (n == 0 || c == 0 || hw == 0) return 0;
  VSP_REQUIRE(x && y, "nchw->nhwc split3: null pointer");
  VSP_REQUIRE(n <= 65535 && ceil_div64(c, 64) <= 65535, "nchw->nhwc split3: batch/channel extent too large");
  dim3 grid((unsigned)ceil_div64(hw, 64), (unsigned)ceil_div64(c, 64), (unsigned)n);
  nchw_to_nhwc_split3_kernel<<<grid, kThreads, 0, stream>>>(x, scale_nc, static_cast<__nv_bfloat16 *>(y), c, hw);
  return check_launch("nchw_to_nhwc_split3_kernel");
}

extern "C" int vsp_nchw_f32_to_nhwc_bf16(const float *x, const float *scale_nc, void *y, int64_t n,
                                         int64_t c, int64_t hw, int64_t c_pad, void *stream_) {
  return vsp_nchw_f32_to_nhwc_bf16_dot(x, scale_nc, nullptr, y, nullptr, n, c, hw, c_pad, stream_);
}

extern "C" int vsp_nhwc_bf16_to_nchw_f32_dot(const void *x, const float *scale_nc, const float *other, float *y,
                                             float *dot_nc, int64_t n, int64_t c, int64_t hw, int64_t c_pad,
                                             void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(n >= 0 && c >= 0 && hw >= 0 && c_pad >= c && c_pad % 2 == 0, "nhwc->nchw: bad geometry");
  VSP_REQUIRE((other == nullptr) == (dot_nc == nullptr), "nhwc->nchw: `other` and `dot_nc` go together");
  if (dot_nc != nullptr && n * c > 0) VSP_CUDA(cudaMemsetAsync(dot_nc, 0, sizeof(float) * n * c, stream));
  if (n == 0 || c == 0 || hw == 0) return 0;
  VSP_REQUIRE(x && y, "nhwc->nchw: null pointer");
  VSP_REQUIRE(n <= 65535 && ceil_div64(c_pad, 64) <= 65535, "nhwc->nchw: batch/channel extent too large");
  if (vec4_ok(x, y, other, hw, c_pad)) {
    dim3 grid((unsigned)ceil_div64(hw, kTileP), (unsigned)ceil_div64(c_pad, kTileC), (unsigned)n);
    if (dot_nc != nullptr)
      nhwc_to_nchw_v4_kernel<true><<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16 *>(x), scale_nc, other,
                                                                   y, dot_nc, c, hw, c_pad);
    else
      nhwc_to_nchw_v4_kernel<false><<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16 *>(x), scale_nc,
                                                                    nullptr, y, nullptr, c, hw, c_pad);
    return check_launch("nhwc_to_nchw_v4_kernel");
  }
  VSP_REQUIRE(dot_nc == nullptr && scale_nc == nullptr,
              "nhwc->nchw: the fused scale / dot needs hw %% 4 == 0 and 16-byte aligned tensors");
  dim3 grid((unsigned)ceil_div64(hw, 64), (unsigned)ceil_div64(c_pad, 64), (unsigned)n);
  nhwc_to_nchw_kernel<<<grid, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16 *>(x), y, c, hw, c_pad);
  return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int vsp_nhwc_bf16_to_nchw_f32(const void *x, float *y, int64_t n, int64_t c, int64_t hw,
                                         int64_t c_pad, void *stream_) {
  return vsp_nhwc_bf16_to_nchw_f32_dot(x, nullptr, nullptr, y, nullptr, n, c, hw, c_pad, stream_);
}

extern "C" int vsp_nchw_f32_to_bf16(const float *x, const float *scale_nc, void *y, int64_t planes,
                                    int64_t hw, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(planes >= 0 && hw >= 0, "nchw cast: bad geometry");
  if (planes == 0 || hw == 0) return 0;
  VSP_REQUIRE(x && y, "nchw cast: null pointer");
  VSP_REQUIRE(planes <= 65535LL * 65535LL, "nchw cast: too many planes");
  const int vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(y) & 7) == 0);
  long long bx = ceil_div64(vec ? hw / 4 : hw, kThreads);
  if (bx > 64) bx = 64;
  for (int64_t p0 = 0; p0 < planes; p0 += 65535) {
    const int64_t np = planes - p0 < 65535 ? planes - p0 : 65535;
    dim3 grid((unsigned)bx, (unsigned)np);
    nchw_cast_kernel<<<grid, kThreads, 0, stream>>>(x + p0 * hw, scale_nc ? scale_nc + p0 : nullptr,
                                                    static_cast<__nv_bfloat16 *>(y) + p0 * hw, hw, vec);
    if (int rc = check_launch("nchw_cast_kernel")) return rc;
  }
  return 0;
}

// Image quantisation for the output side (restoration_test.py:138-157 saves every image through
// torchvision.utils.save_image(normalize=True, range=(-1, 1))): NCHW fp32 -> HWC uint8 on the device, so the device->host
// copy moves 3 bytes per pixel instead of 12 and the host only has to encode.  Same arithmetic, operation by operation, as
// torchvision (clamp to [lo, hi]; (x - lo) / max(hi - lo, 1e-5); * 255; + 0.5; clamp to [0, 255]; truncate) in
// round-to-nearest fp32 without contraction, so the bytes are identical.
namespace vsp {
namespace {
__device__ __forceinline__ unsigned quant_u8(float v, float lo, float hi, float inv_is_div, float span) {
  v = fminf(fmaxf(v, lo), hi);
  v = __fsub_rn(v, lo);
  v = __fdiv_rn(v, span);
  v = __fmul_rn(v, 255.f);
  v = __fadd_rn(v, 0.5f);
  v = fminf(fmaxf(v, 0.f), 255.f);
  (void)inv_is_div;
  return (unsigned)(int)v;          // truncation, as .to(torch.uint8)
}

__global__ void __launch_bounds__(256)
quantize_nchw_to_hwc_u8_kernel(const float *__restrict__ x, unsigned char *__restrict__ y, long long hw, int c, float lo,
                               float hi) {
  const long long b = blockIdx.y;
  const float *xb = x + b * c * hw;
  unsigned char *yb = y + b * c * hw;
  const float span = fmaxf(hi - lo, 1e-5f);
  if (c == 3 && (hw & 3) == 0) {
    // four pixels per thread: three 128-bit plane loads -> twelve bytes = three 32-bit stores
    for (long long i = (blockIdx.x * 256LL + threadIdx.x) * 4; i < hw; i += (long long)gridDim.x * 1024) {
      const float4 r = ld_stream_f4(reinterpret_cast<const float4 *>(xb + i));
      const float4 g = ld_stream_f4(reinterpret_cast<const float4 *>(xb + hw + i));
      const float4 bl = ld_stream_f4(reinterpret_cast<const float4 *>(xb + 2 * hw + i));
      const float rr[4] = {r.x, r.y, r.z, r.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {bl.x, bl.y, bl.z, bl.w};
      unsigned q[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        q[3 * k] = quant_u8(rr[k], lo, hi, 0.f, span);
        q[3 * k + 1] = quant_u8(gg[k], lo, hi, 0.f, span);
        q[3 * k + 2] = quant_u8(bb[k], lo, hi, 0.f, span);
      }
      uint32_t *dst = reinterpret_cast<uint32_t *>(yb + i * 3);
#pragma unroll
      for (int w = 0; w < 3; ++w)
        dst[w] = q[4 * w] | (q[4 * w + 1] << 8) | (q[4 * w + 2] << 16) | (q[4 * w + 3] << 24);
    }
  } else {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < hw; i += (long long)gridDim.x * 256)
      for (int ch = 0; ch < c; ++ch) yb[i * c + ch] = (unsigned char)quant_u8(xb[ch * hw + i], lo, hi, 0.f, span);
  }
}
}  // namespace
}  // namespace vsp

extern "C" int vsp_quantize_nchw_f32_to_hwc_u8(const float *x, void *y, int64_t batch, int64_t channels, int64_t hw,
                                               float lo, float hi, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(batch >= 0 && channels >= 1 && channels <= 4 && hw >= 0, "quantize: bad geometry");
  if (batch == 0 || hw == 0) return 0;
  VSP_REQUIRE(x && y, "quantize: null pointer");
  VSP_REQUIRE(batch <= 65535, "quantize: batch too large");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 3) == 0,
              "quantize: operands must be 16 / 4-byte aligned");
  long long bx = ceil_div64(hw, 1024);
  if (bx > 2 * num_sms()) bx = 2 * num_sms();
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)batch);
  quantize_nchw_to_hwc_u8_kernel<<<grid, 256, 0, stream>>>(x, static_cast<unsigned char *>(y), hw, (int)channels, lo, hi);
  return check_launch("quantize_nchw_to_hwc_u8_kernel");
}

extern "C" int vsp_modulate_weights_bf16(const float *w, const float *s, float *demod, void *wq,
                                         int64_t batch, int64_t cout, int64_t cin, int taps,
                                         float wscale, float eps, int transpose, int fold_demod,
                                         int64_t n_pad, int64_t k_pad, const float *wsq, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(batch >= 1 && cout >= 1 && cin >= 1 && taps >= 1, "modulate_weights: bad geometry");
  VSP_REQUIRE(w != nullptr, "modulate_weights: null weight");
  VSP_REQUIRE(!(fold_demod && !demod), "modulate_weights: fold_demod needs a demod buffer");
  VSP_REQUIRE(batch <= 65535, "modulate_weights: batch too large");
  if (demod) {
    const long long warps = batch * cout;
    if (wsq) {
      demod_from_wsq_kernel<<<(unsigned)ceil_div64(warps * 32, kThreads), kThreads, 0, stream>>>(wsq, s, demod, batch,
                                                                                                 cout, cin, wscale, eps);
      if (int rc = check_launch("demod_from_wsq_kernel")) return rc;
    } else {
      demod_kernel<<<(unsigned)ceil_div64(warps * 32, kThreads), kThreads, 0, stream>>>(w, s, demod, batch, cout, cin,
                                                                                        taps, wscale, eps);
      if (int rc = check_launch("demod_kernel")) return rc;
    }
  }
  if (wq) {
    const int64_t n_real = transpose ? cin : cout, k_real = transpose ? cout : cin;
    VSP_REQUIRE(n_pad >= n_real && k_pad >= k_real, "modulate_weights: padded extents smaller than the real ones");
    const bool vec_ok = !transpose && cin % 8 == 0 && k_pad == cin && (taps == 9 || taps == 1) &&
                        (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(wq) & 15) == 0 &&
                        (s == nullptr || (reinterpret_cast<uintptr_t>(s) & 15) == 0);
    if (vec_ok) {
      dim3 vgrid((unsigned)ceil_div64(n_pad * (cin / 8), kThreads), (unsigned)ceil_div64(batch, kBG));
      if (taps == 9)
        pack_weights_vec_kernel<9><<<vgrid, kThreads, 0, stream>>>(w, s, fold_demod ? demod : nullptr,
                                                                    static_cast<uint4 *>(wq), cout, cin, wscale, n_pad, batch);
      else
        pack_weights_vec_kernel<1><<<vgrid, kThreads, 0, stream>>>(w, s, fold_demod ? demod : nullptr,
                                                                    static_cast<uint4 *>(wq), cout, cin, wscale, n_pad, batch);
      return check_launch("pack_weights_vec_kernel");
    }
    dim3 grid((unsigned)ceil_div64(n_pad * k_pad, kThreads), (unsigned)batch);
    pack_weights_kernel<<<grid, kThreads, 0, stream>>>(w, s, fold_demod ? demod : nullptr,
                                                       static_cast<__nv_bfloat16 *>(wq), cout, cin, taps, wscale,
                                                       transpose, n_pad, k_pad);
    if (int rc = check_launch("pack_weights_kernel")) return rc;
  }
  return 0;
}

extern "C" int vsp_weight_sumsq_f32(const float *w, float *wsq, int64_t cout, int64_t cin, int taps, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(w && wsq && cout >= 1 && cin >= 1 && taps >= 1, "weight_sumsq: bad arguments");
  weight_sumsq_kernel<<<(unsigned)ceil_div64(cout * cin, kThreads), kThreads, 0, stream>>>(w, wsq, cout * cin, taps);
  return check_launch("weight_sumsq_kernel");
}

// ---- weight gradient of a 1x1 convolution with very few input channels (the RGB-side layers: LargeConvLayer 3->16,
// the Discriminator's 3->64 stem; models/RestoreNet.py:725-787, :1218).  gw[o,i] = sum_{b,p} dy[b,o,p] * x[b,i,p] is a
// [Cout x Cin] reduction over ~1e6 pixels: ATen hands it to a tall-skinny GEMM that takes 3.5-3.8 ms per call at 512^2
// (29 ms of a 160 ms training step); as a streaming reduction it is a read of dy (17-67 MB): one block owns 8 output
// channels x a chunk of pixels of one sample, accumulates 8 x Cin partial sums per thread over 128-bit loads, reduces them
// with shuffles + shared memory and issues one atomic per (o, i).
constexpr int kSmallCin = 8;
__global__ void __launch_bounds__(kThreads)
conv1x1_wgrad_small_kernel(const float4 *__restrict__ dy, const float4 *__restrict__ x, float *__restrict__ gw,
                           long long p4, int cout, int cin, int iters) {
  __shared__ float red[kThreads / 32][8 * kSmallCin];
  const int og = blockIdx.y * 8;
  const long long b = blockIdx.z;
  float acc[8][kSmallCin];
#pragma unroll
  for (int o = 0; o < 8; ++o)
#pragma unroll
    for (int i = 0; i < kSmallCin; ++i) acc[o][i] = 0.f;
  const float4 *dyb = dy + (b * cout + og) * p4;
  const float4 *xb = x + b * cin * p4;
  for (int it = 0; it < iters; ++it) {
    const long long p = ((long long)blockIdx.x * iters + it) * kThreads + threadIdx.x;
    if (p >= p4) break;
    float4 xv[kSmallCin];
#pragma unroll
    for (int i = 0; i < kSmallCin; ++i) xv[i] = i < cin ? __ldg(xb + i * p4 + p) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (og + o >= cout) break;
      const float4 g = ld_stream_f4(dyb + o * p4 + p);
#pragma unroll
      for (int i = 0; i < kSmallCin; ++i)
        acc[o][i] = fmaf(g.x, xv[i].x, fmaf(g.y, xv[i].y, fmaf(g.z, xv[i].z, fmaf(g.w, xv[i].w, acc[o][i]))));
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 0; o < 8; ++o)
#pragma unroll
    for (int i = 0; i < kSmallCin; ++i) {
      const float v = warp_sum(acc[o][i]);
      if (lane == 0) red[warp][o * kSmallCin + i] = v;
    }
  __syncthreads();
  if (threadIdx.x < 8 * kSmallCin) {
    const int o = threadIdx.x / kSmallCin, i = threadIdx.x % kSmallCin;
    if (og + o < cout && i < cin) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) v += red[w][threadIdx.x];
      atomicAdd(gw + (long long)(og + o) * cin + i, v);
    }
  }
}

extern "C" int vsp_conv1x1_wgrad_small_f32(const float *dy, const float *x, float *gw, int64_t batch, int64_t cout,
                                           int64_t cin, int64_t hw, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(batch >= 1 && cout >= 1 && cin >= 1 && cin <= kSmallCin && hw >= 4 && hw % 4 == 0,
              "conv1x1_wgrad_small: needs 1 <= cin <= 8 and a pixel count that is a multiple of 4");
  VSP_REQUIRE(dy && x && gw, "conv1x1_wgrad_small: null pointer");
  VSP_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 15) == 0,
              "conv1x1_wgrad_small: tensors must be 16-byte aligned");
  VSP_REQUIRE(batch <= 65535 && cout <= 8 * 65535, "conv1x1_wgrad_small: extent too large");
  VSP_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * cout * cin, stream));
  const long long p4 = hw / 4;
  const int iters = 8;
  dim3 grid((unsigned)ceil_div64(p4, (long long)kThreads * iters), (unsigned)ceil_div64(cout, 8), (unsigned)batch);
  conv1x1_wgrad_small_kernel<<<grid, kThreads, 0, stream>>>(reinterpret_cast<const float4 *>(dy),
                                                            reinterpret_cast<const float4 *>(x), gw, p4, (int)cout,
                                                            (int)cin, iters);
  return check_launch("conv1x1_wgrad_small_kernel");
}

// ---- tail of an IR-SE residual unit (e4e encoder, helpers.py:97-123), channels-last bf16:
//   y = res * gate[n,c] + shortcut            (SE scale + residual add)
//   z = bf16(y) * bn_a[c] + bn_b[c]           (the NEXT unit's leading eval-mode BatchNorm; optional)
// One pass instead of three (mul, add, batch_norm) over the activation: reads res + shortcut, writes y (+ z).  The shortcut
// may be a strided view (MaxPool2d(1, 2) == x[:, :, ::2, ::2]): its pixel strides are passed in elements.
__global__ void __launch_bounds__(kThreads)
se_tail_nhwc_kernel(const uint4 *__restrict__ res, const float *__restrict__ gate, const __nv_bfloat16 *__restrict__ sc,
                    uint4 *__restrict__ y, uint4 *__restrict__ z, const float *__restrict__ bn_a,
                    const float *__restrict__ bn_b, long long total, int h, int w, int cg, long long sc_n, long long sc_h,
                    long long sc_w) {
  for (long long idx = blockIdx.x * (long long)kThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kThreads) {
    const int g = (int)(idx % cg);
    long long t = idx / cg;
    const int xx = (int)(t % w); t /= w;
    const int yy = (int)(t % h);
    const long long n = t / h;
    const uint4 rv = __ldg(res + idx);
    const uint4 sv = __ldg(reinterpret_cast<const uint4 *>(sc + n * sc_n + yy * sc_h + xx * sc_w + g * 8));
    const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gate + n * cg * 8 + g * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4 *>(gate + n * cg * 8 + g * 8) + 1);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const __nv_bfloat162 *rh = reinterpret_cast<const __nv_bfloat162 *>(&rv);
    const __nv_bfloat162 *sh = reinterpret_cast<const __nv_bfloat162 *>(&sv);
    uint4 yo;
    __nv_bfloat162 *yh = reinterpret_cast<__nv_bfloat162 *>(&yo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 r2 = __bfloat1622float2(rh[i]), s2 = __bfloat1622float2(sh[i]);
      yh[i] = __floats2bfloat162_rn(fmaf(r2.x, gv[2 * i], s2.x), fmaf(r2.y, gv[2 * i + 1], s2.y));
    }
    y[idx] = yo;
    if (z != nullptr) {
      const float4 a0 = __ldg(reinterpret_cast<const float4 *>(bn_a + g * 8)), a1 = __ldg(reinterpret_cast<const float4 *>(bn_a + g * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bn_b + g * 8)), b1 = __ldg(reinterpret_cast<const float4 *>(bn_b + g * 8) + 1);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint4 zo;
      __nv_bfloat162 *zh = reinterpret_cast<__nv_bfloat162 *>(&zo);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 y2 = __bfloat1622float2(yh[i]);       // the rounded y, as the separate BatchNorm pass would read it
        zh[i] = __floats2bfloat162_rn(fmaf(y2.x, av[2 * i], bv[2 * i]), fmaf(y2.y, av[2 * i + 1], bv[2 * i + 1]));
      }
      z[idx] = zo;
    }
  }
}

extern "C" int vsp_se_tail_nhwc_bf16(const void *res, const float *gate, const void *shortcut, void *y, void *z,
                                     const float *bn_a, const float *bn_b, int64_t batch, int64_t h, int64_t w, int64_t c,
                                     int64_t sc_n_stride, int64_t sc_h_stride, int64_t sc_w_stride, void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(batch >= 0 && h >= 0 && w >= 0 && c >= 8 && c % 8 == 0, "se_tail: channels must be a positive multiple of 8");
  if (batch == 0 || h == 0 || w == 0) return 0;
  VSP_REQUIRE(res && gate && shortcut && y && (z == nullptr || (bn_a && bn_b)), "se_tail: null pointer");
  VSP_REQUIRE(((reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(shortcut) | reinterpret_cast<uintptr_t>(y) |
                reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gate) | reinterpret_cast<uintptr_t>(bn_a) |
                reinterpret_cast<uintptr_t>(bn_b)) & 15) == 0 &&
                  ((sc_n_stride | sc_h_stride | sc_w_stride) & 7) == 0,
              "se_tail: pointers must be 16-byte aligned and shortcut strides multiples of 8 elements");
  const long long total = batch * h * w * (c / 8);
  long long nb = (total + kThreads - 1) / kThreads;
  if (nb > (long long)num_sms() * 32) nb = (long long)num_sms() * 32;
  se_tail_nhwc_kernel<<<(unsigned)nb, kThreads, 0, stream>>>(static_cast<const uint4 *>(res), gate,
                                                            static_cast<const __nv_bfloat16 *>(shortcut),
                                                            static_cast<uint4 *>(y), static_cast<uint4 *>(z), bn_a, bn_b,
                                                            total, (int)h, (int)w, (int)(c / 8), sc_n_stride, sc_h_stride,
                                                            sc_w_stride);
  return check_launch("se_tail_nhwc_kernel");
}

extern "C" int vsp_scale_nhwc_bf16(const void *x, const float *s, void *y, int64_t batch, int64_t hw, int64_t c,
                                   void *stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(batch >= 0 && hw >= 0 && c >= 8 && c % 8 == 0, "scale_nhwc: channels must be a positive multiple of 8");
  if (batch == 0 || hw == 0) return 0;
  VSP_REQUIRE(x && s && y, "scale_nhwc: null pointer");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(s) & 15) == 0,
              "scale_nhwc: pointers must be 16-byte aligned");
  const long long total = batch * hw * (c / 8);
  long long nb = ceil_div64(total, kThreads);
  if (nb > (long long)num_sms() * 32) nb = (long long)num_sms() * 32;
  scale_nhwc_kernel<<<(unsigned)nb, kThreads, 0, stream>>>(static_cast<const uint4 *>(x), s, static_cast<uint4 *>(y), hw,
                                                          (int)(c / 8), c, total);
  return check_launch("scale_nhwc_kernel");
}
