// torgb_sm100.cu — ToRGB: 1x1 modulated convolution to 3 channels (no demodulation) + bias + skip.
//
// Reference: models/RestoreNet.py:647-666 / e4e/models/stylegan2/model.py:345-364.  With N = 3 the
// "GEMM" is a memory-bound read of the feature map, so it does not go through
// the tensor-core kernel (whose 128x16 tile would be 81 % padding and per-tile latency bound): each
// thread owns 4 pixels, streams their NHWC bf16 channel vectors with 128-bit loads, and applies the
// per-sample modulated weights wscale*W[o,c]*s[b,c] held in shared memory as one float4 per channel.
// Output is fp32 NCHW (the RGB skip chain stays full precision), coalesced across pixels.
#include "common.cuh"

namespace vsp {
namespace {

constexpr int kThreads = 256;
constexpr int kPix = 4;  // pixels per thread
constexpr int kWarpPix = 8;  // pixels per warp in the warp-per-pixel variant

__global__ void __launch_bounds__(kThreads)
torgb_kernel(const uint4 *__restrict__ x, const float *__restrict__ w, const float *__restrict__ s,
             const float *__restrict__ bias, const float *__restrict__ skip, float *__restrict__ out,
             long long hw, int c, float wscale) {
  extern __shared__ float4 wm[];  // [c] : (w0, w1, w2, 0) * s[b,c] * wscale
  const long long b = blockIdx.y;
  for (int ch = threadIdx.x; ch < c; ch += kThreads) {
    const float f = wscale * (s ? __ldg(s + b * c + ch) : 1.f);
    wm[ch] = make_float4(__ldg(w + ch) * f, __ldg(w + c + ch) * f, __ldg(w + 2 * c + ch) * f, 0.f);
  }
  __syncthreads();
  const int cg = c / 8;
  const long long p0 = (long long)blockIdx.x * (kThreads * kPix) + threadIdx.x;
  float acc[kPix][3];
#pragma unroll
  for (int j = 0; j < kPix; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
  const uint4 *xb = x + b * hw * cg;
  for (int g = 0; g < cg; ++g) {
    uint4 v[kPix];
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
      const long long p = p0 + (long long)j * kThreads;
      v[j] = p < hw ? __ldg(xb + p * cg + g) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float4 wv = wm[g * 8 + e];
#pragma unroll
      for (int j = 0; j < kPix; ++j) {
        const uint32_t word = (&v[j].x)[e >> 1];
        const float xv = __uint_as_float((e & 1) ? (word & 0xFFFF0000u) : (word << 16));   // bf16 -> fp32
        acc[j][0] = fmaf(xv, wv.x, acc[j][0]);
        acc[j][1] = fmaf(xv, wv.y, acc[j][1]);
        acc[j][2] = fmaf(xv, wv.z, acc[j][2]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kPix; ++j) {
    const long long p = p0 + (long long)j * kThreads;
    if (p >= hw) continue;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const long long off = (b * 3 + o) * hw + p;
      float v = acc[j][o] + (bias ? __ldg(bias + o) : 0.f);
      if (skip) v += ld_stream_f1(skip + off);
      st_stream_f1(out + off, v);
    }
  }
}

// Final ToRGB of the style decoder fused with everything that follows it in the reference (psp.py:245-246 face_pool):
//   image = AdaptiveAvgPool2d(H/2)( conv1x1(x) + bias + Upsample(skip) )
// Pooling is linear, so one thread sums the 2x2 pixel block's channel vectors, does ONE dot per output channel, and adds
// the previous level's skip through the 3x3 composite of (2x FIR upsample, then 2x2 mean) — neither the full-resolution
// RGB image nor the upsampled skip is ever written (saves ~1.6 GB of traffic per 32 faces at 1024^2).
struct PoolTaps {
  float k[9];
};

__global__ void __launch_bounds__(kThreads)
torgb_pool2_kernel(const uint4 *__restrict__ x, const float *__restrict__ w, const float *__restrict__ s,
                   const float *__restrict__ bias, const float *__restrict__ skip, float *__restrict__ out,
                   int oh, int ow, int c, float wscale, const PoolTaps taps) {
  extern __shared__ float4 wm[];
  const long long b = blockIdx.y;
  for (int ch = threadIdx.x; ch < c; ch += kThreads) {
    const float f = 0.25f * wscale * (s ? __ldg(s + b * c + ch) : 1.f);      // 0.25 = the 2x2 mean
    wm[ch] = make_float4(__ldg(w + ch) * f, __ldg(w + c + ch) * f, __ldg(w + 2 * c + ch) * f, 0.f);
  }
  __syncthreads();
  const long long p = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (p >= (long long)oh * ow) return;
  const int oy = (int)(p / ow), ox = (int)(p % ow);
  const int cg = c / 8;
  const long long iw = 2LL * ow;
  const uint4 *x00 = x + ((b * 2 * oh + 2 * oy) * iw + 2 * ox) * cg;
  const uint4 *x10 = x00 + iw * cg;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int g = 0; g < cg; ++g) {
    const uint4 v[4] = {__ldg(x00 + g), __ldg(x00 + cg + g), __ldg(x10 + g), __ldg(x10 + cg + g)};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float xv = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t word = (&v[q].x)[e >> 1];
        xv += __uint_as_float((e & 1) ? (word & 0xFFFF0000u) : (word << 16));
      }
      const float4 wv = wm[g * 8 + e];
      a0 = fmaf(xv, wv.x, a0);
      a1 = fmaf(xv, wv.y, a1);
      a2 = fmaf(xv, wv.z, a2);
    }
  }
  const float acc[3] = {a0, a1, a2};
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    float v = acc[o] + (bias ? __ldg(bias + o) : 0.f);
    if (skip) {
      const float *sp = skip + (b * 3 + o) * (long long)oh * ow;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int yy = oy + dy - 1, xx = ox + dx - 1;
          if (yy >= 0 && yy < oh && xx >= 0 && xx < ow) v = fmaf(taps.k[dy * 3 + dx], __ldg(sp + (long long)yy * ow + xx), v);
        }
    }
    st_stream_f1(out + (b * 3 + o) * (long long)oh * ow + p, v);
  }
}

// Small images (hw <= 4096): one warp per pixel, lanes split the channel groups and reduce with shuffles —
// the pixel-per-thread form above would leave most SMs idle and serialise 64 dependent loads per thread.
__global__ void __launch_bounds__(kThreads)
torgb_warp_kernel(const uint4 *__restrict__ x, const float *__restrict__ w, const float *__restrict__ s,
                  const float *__restrict__ bias, const float *__restrict__ skip, float *__restrict__ out,
                  long long hw, int c, float wscale) {
  extern __shared__ float4 wm[];
  const long long b = blockIdx.y;
  for (int ch = threadIdx.x; ch < c; ch += kThreads) {
    const float f = wscale * (s ? __ldg(s + b * c + ch) : 1.f);
    wm[ch] = make_float4(__ldg(w + ch) * f, __ldg(w + c + ch) * f, __ldg(w + 2 * c + ch) * f, 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int cg = c / 8;
  // kWarpPix consecutive pixels per warp (amortises the per-block weight staging above), four at a time so that
  // every lane keeps 4 x cg/32 independent 128-bit loads in flight
  const long long p0 = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * kWarpPix;
  for (int i0 = 0; i0 < kWarpPix; i0 += 4) {
    if (p0 + i0 >= hw) return;
    float acc[4][3];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
    for (int g = lane; g < cg; g += 32) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        v[q] = (p0 + i0 + q < hw) ? __ldg(x + (b * hw + p0 + i0 + q) * cg + g) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 wv = wm[g * 8 + e];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t word = (&v[q].x)[e >> 1];
          const float xv = __uint_as_float((e & 1) ? (word & 0xFFFF0000u) : (word << 16));
          acc[q][0] = fmaf(xv, wv.x, acc[q][0]);
          acc[q][1] = fmaf(xv, wv.y, acc[q][1]);
          acc[q][2] = fmaf(xv, wv.z, acc[q][2]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int o = 0; o < 3; ++o) acc[q][o] = warp_sum(acc[q][o]);
    if (lane < 12) {
      const int q = lane / 3, o = lane % 3;
      const long long p = p0 + i0 + q;
      if (p < hw) {
        float v = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
#pragma unroll
          for (int oo = 0; oo < 3; ++oo) v = (qq == q && oo == o) ? acc[qq][oo] : v;
        const long long off = (b * 3 + o) * hw + p;
        v += bias ? __ldg(bias + o) : 0.f;
        if (skip) v += __ldg(skip + off);
        out[off] = v;
      }
    }
  }
}

}  // namespace
}  // namespace vsp

extern "C" int vsp_torgb_nhwc_bf16(const void *x, const float *w, const float *s, const float *bias,
                                   const float *skip, float *out, int64_t batch, int64_t hw, int64_t c,
                                   float wscale, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(batch >= 0 && hw >= 0 && c >= 8 && c % 8 == 0, "torgb: channels must be a positive multiple of 8");
  if (batch == 0 || hw == 0) return 0;
  VSP_REQUIRE(x && w && out, "torgb: null pointer");
  VSP_REQUIRE(batch <= 65535 && c <= 2048, "torgb: batch/channel extent too large");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "torgb: x must be 16-byte aligned");
  if (hw <= 4096) {
    dim3 grid((unsigned)ceil_div64(hw, (kThreads / 32) * kWarpPix), (unsigned)batch);
    torgb_warp_kernel<<<grid, kThreads, sizeof(float4) * c, stream>>>(static_cast<const uint4 *>(x), w, s, bias, skip, out,
                                                                      hw, (int)c, wscale);
    return check_launch("torgb_warp_kernel");
  }
  dim3 grid((unsigned)ceil_div64(hw, kThreads * kPix), (unsigned)batch);
  torgb_kernel<<<grid, kThreads, sizeof(float4) * c, stream>>>(static_cast<const uint4 *>(x), w, s, bias, skip, out,
                                                               hw, (int)c, wscale);
  return check_launch("torgb_kernel");
}

extern "C" int vsp_torgb_pool2_nhwc_bf16(const void *x, const float *w, const float *s, const float *bias,
                                         const float *skip, const float *k3_host, float *out, int64_t batch,
                                         int64_t out_h, int64_t out_w, int64_t c, float wscale, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(batch >= 0 && out_h >= 0 && out_w >= 0 && c >= 8 && c % 8 == 0, "torgb_pool2: channels must be a positive multiple of 8");
  if (batch == 0 || out_h == 0 || out_w == 0) return 0;
  VSP_REQUIRE(x && w && out && (skip == nullptr || k3_host != nullptr), "torgb_pool2: null pointer");
  VSP_REQUIRE(batch <= 65535 && c <= 2048 && out_h < 32768 && out_w < 32768, "torgb_pool2: extent too large");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "torgb_pool2: x must be 16-byte aligned");
  PoolTaps t;
  for (int i = 0; i < 9; ++i) t.k[i] = k3_host ? k3_host[i] : 0.f;
  dim3 grid((unsigned)ceil_div64(out_h * out_w, kThreads), (unsigned)batch);
  torgb_pool2_kernel<<<grid, kThreads, sizeof(float4) * c, stream>>>(static_cast<const uint4 *>(x), w, s, bias, skip, out,
                                                                    (int)out_h, (int)out_w, (int)c, wscale, t);
  return check_launch("torgb_pool2_kernel");
}
