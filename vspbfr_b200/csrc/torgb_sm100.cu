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
