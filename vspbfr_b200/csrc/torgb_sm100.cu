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


// ---- lane-split variants (the default for the model's channel counts) --------------------------------------------------
// The pixel-per-thread kernels above issue one 16-byte load per lane at a 2*C-byte stride: 32 different 128-byte lines per
// warp instruction, and the L1 moves one line per clock — a hard ceiling of 16 B/clk/SM (~4 TB/s; measured 3.2-4.3 TB/s).
// Here L = C/8 lanes share one pixel (each owns one 8-channel group, its 24 modulated weights live in registers), so a
// warp-wide 128-bit load covers 512 contiguous bytes (4 lines).  Products use the packed fp32 FMA (fma.rn.f32x2: even /
// odd channels accumulate in the two halves), and the 4 rows x 3 colours of partial sums are combined with a
// reduce-scatter over the lane group (12 shuffles for L = 8 instead of 36 for a butterfly per value).
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float sum2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

constexpr int kRows = 4;  // independent pixel rows (loads in flight) per lane and iteration

// modulated weights of channel group g as (even, odd) channel pairs: wr[o][i] = wscale * s[b, 8g+2i(+1)] * W[o, 8g+2i(+1)]
__device__ __forceinline__ void load_lane_weights(unsigned long long (&wr)[3][4], const float *w, const float *s, long long b,
                                                  int c, int g, float wscale) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ch = g * 8 + 2 * i;
    const float f0 = wscale * (s ? __ldg(s + b * c + ch) : 1.f), f1 = wscale * (s ? __ldg(s + b * c + ch + 1) : 1.f);
#pragma unroll
    for (int o = 0; o < 3; ++o) wr[o][i] = pack2(__ldg(w + o * c + ch) * f0, __ldg(w + o * c + ch + 1) * f1);
  }
}

__device__ __forceinline__ void dot8(unsigned long long (&acc)[3], const uint4 &v, const unsigned long long (&wr)[3][4]) {
  const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned long long xv = pack2(__uint_as_float(wd[i] << 16), __uint_as_float(wd[i] & 0xFFFF0000u));
#pragma unroll
    for (int o = 0; o < 3; ++o) acc[o] = ffma2(xv, wr[o][i], acc[o]);
  }
}

// Reduce-scatter of t[row][colour] over the L lanes of a group (L >= 4): afterwards the lanes whose (bit0, bit1) of `sub`
// encode row = 2*bit0 + bit1 hold that row's three complete sums in k[0..2].
template <int L>
__device__ __forceinline__ int reduce_rows(const float (&t)[kRows][3], float (&k)[3], int sub) {
  const bool b0 = (sub & 1) != 0, b1 = (sub & 2) != 0;
  float k2[2][3];
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    const float s0 = b0 ? t[0][o] : t[2][o], s1 = b0 ? t[1][o] : t[3][o];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    k2[0][o] = (b0 ? t[2][o] : t[0][o]) + r0;
    k2[1][o] = (b0 ? t[3][o] : t[1][o]) + r1;
  }
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    const float sn = b1 ? k2[0][o] : k2[1][o];
    const float rc = __shfl_xor_sync(0xffffffffu, sn, 2);
    k[o] = (b1 ? k2[1][o] : k2[0][o]) + rc;
  }
#pragma unroll
  for (int m = 4; m < L; m <<= 1)
#pragma unroll
    for (int o = 0; o < 3; ++o) k[o] += __shfl_xor_sync(0xffffffffu, k[o], m);
  return (b0 ? 2 : 0) + (b1 ? 1 : 0);
}

template <int L, int V>
__global__ void __launch_bounds__(kThreads)
torgb_lanes_kernel(const uint4 *__restrict__ x, const float *__restrict__ w, const float *__restrict__ s,
                   const float *__restrict__ bias, const float *__restrict__ skip, float *__restrict__ out,
                   long long hw, int c, float wscale, int pix_per_warp) {
  constexpr int P = 32 / L;  // pixels per warp-wide load
  const int lane = threadIdx.x & 31, sub = lane % L, pq = lane / L;
  const long long b = blockIdx.y;
  const int cg = L * V;
  unsigned long long wr[V][3][4];
#pragma unroll
  for (int v = 0; v < V; ++v) load_lane_weights(wr[v], w, s, b, c, sub + v * L, wscale);
  const long long wp0 = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * pix_per_warp;
  long long wp1 = wp0 + pix_per_warp;
  if (wp1 > hw) wp1 = hw;
  const uint4 *xb = x + b * hw * cg + sub;
  uint4 nxt[kRows][V];      // software prefetch: the next iteration's loads are in flight while this one is reduced
  auto fetch = [&](long long p0) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const long long p = p0 + r * P + pq;
#pragma unroll
      for (int q = 0; q < V; ++q) nxt[r][q] = p < wp1 ? ld_stream_u4(xb + p * cg + q * L) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if (wp0 < wp1) fetch(wp0);
  for (long long p0 = wp0; p0 < wp1; p0 += P * kRows) {
    uint4 v[kRows][V];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int q = 0; q < V; ++q) v[r][q] = nxt[r][q];
    if (p0 + P * kRows < wp1) fetch(p0 + P * kRows);
    float t[kRows][3];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      unsigned long long acc[3] = {0ull, 0ull, 0ull};
#pragma unroll
      for (int q = 0; q < V; ++q) dot8(acc, v[r][q], wr[q]);
#pragma unroll
      for (int o = 0; o < 3; ++o) t[r][o] = sum2(acc[o]);
    }
    float k[3];
    const int row = reduce_rows<L>(t, k, sub);
    const long long p = p0 + row * P + pq;
    if ((sub >> 2) == 0 && p < wp1) {
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const long long off = (b * 3 + o) * hw + p;
        float r = k[o] + (bias ? __ldg(bias + o) : 0.f);
        if (skip) r += ld_stream_f1(skip + off);
        st_stream_f1(out + off, r);
      }
    }
  }
}

// Pooled form (see torgb_pool2_kernel): the L = 2*C/8 lanes of a group cover the two horizontally adjacent input pixels of
// a 2x2 block (contiguous in NHWC); each lane accumulates the block's two rows for its channel group.  `skip` here is
// already at the OUTPUT resolution (the caller applies the 3x3 upsample-then-pool composite with upfirdn2d) and is just added.
template <int L>
__global__ void __launch_bounds__(kThreads)
torgb_pool2_lanes_kernel(const uint4 *__restrict__ x, const float *__restrict__ w, const float *__restrict__ s,
                         const float *__restrict__ bias, const float *__restrict__ skip, float *__restrict__ out,
                         int oh, int ow, int c, float wscale, int pix_per_warp) {
  constexpr int P = 32 / L;
  const int lane = threadIdx.x & 31, sub = lane % L, pq = lane / L;
  const long long b = blockIdx.y;
  const int cg = L / 2;
  unsigned long long wr[3][4];
  load_lane_weights(wr, w, s, b, c, sub % cg, 0.25f * wscale);      // 0.25 = the 2x2 mean
  const long long hw = (long long)oh * ow, iw = 2LL * ow;
  const long long wp0 = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * pix_per_warp;
  long long wp1 = wp0 + pix_per_warp;
  if (wp1 > hw) wp1 = hw;
  const uint4 *xb = x + b * 4 * hw * cg + sub;
  uint4 nxt[kRows][2];
  auto fetch = [&](long long p0) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const long long p = p0 + r * P + pq;
      const int oy = (int)(p / ow), ox = (int)(p % ow);
      const uint4 *x00 = xb + ((2LL * oy) * iw + 2 * ox) * cg;
      nxt[r][0] = p < wp1 ? ld_stream_u4(x00) : make_uint4(0u, 0u, 0u, 0u);
      nxt[r][1] = p < wp1 ? ld_stream_u4(x00 + iw * cg) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  if (wp0 < wp1) fetch(wp0);
  for (long long p0 = wp0; p0 < wp1; p0 += P * kRows) {
    uint4 v[kRows][2];
#pragma unroll
    for (int r = 0; r < kRows; ++r) v[r][0] = nxt[r][0], v[r][1] = nxt[r][1];
    if (p0 + P * kRows < wp1) fetch(p0 + P * kRows);
    float t[kRows][3];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      unsigned long long acc[3] = {0ull, 0ull, 0ull};
      dot8(acc, v[r][0], wr);
      dot8(acc, v[r][1], wr);
#pragma unroll
      for (int o = 0; o < 3; ++o) t[r][o] = sum2(acc[o]);
    }
    float k[3];
    const int row = reduce_rows<L>(t, k, sub);
    const long long p = p0 + row * P + pq;
    if ((sub >> 2) == 0 && p < wp1) {
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const long long off = (b * 3 + o) * hw + p;
        float r = k[o] + (bias ? __ldg(bias + o) : 0.f);
        if (skip) r += ld_stream_f1(skip + off);
        st_stream_f1(out + off, r);
      }
    }
  }
}

// pixels per warp: a multiple of the P*kRows pixels of one iteration, sized for >= ~8 blocks per SM when the image allows
inline int lanes_pix_per_warp(long long hw, long long batch, int pix_per_iter) {
  const long long iters_total = ceil_div64(hw, pix_per_iter) * batch;
  long long it = iters_total / ((long long)num_sms() * 8 * (kThreads / 32));
  if (it < 1) it = 1;
  if (it > 64) it = 64;
  return (int)it * pix_per_iter;
}

template <int L, int V>
int launch_torgb_lanes(const void *x, const float *w, const float *s, const float *bias, const float *skip, float *out,
                       int64_t batch, int64_t hw, int64_t c, float wscale, cudaStream_t stream) {
  const int ppw = lanes_pix_per_warp(hw, batch, (32 / L) * kRows);
  dim3 grid((unsigned)ceil_div64(hw, (long long)ppw * (kThreads / 32)), (unsigned)batch);
  torgb_lanes_kernel<L, V><<<grid, kThreads, 0, stream>>>(static_cast<const uint4 *>(x), w, s, bias, skip, out, hw, (int)c,
                                                           wscale, ppw);
  return check_launch("torgb_lanes_kernel");
}

template <int L>
int launch_torgb_pool2_lanes(const void *x, const float *w, const float *s, const float *bias, const float *skip,
                             float *out, int64_t batch, int64_t oh, int64_t ow, int64_t c, float wscale,
                             cudaStream_t stream) {
  const int ppw = lanes_pix_per_warp(oh * ow, batch, (32 / L) * kRows);
  dim3 grid((unsigned)ceil_div64(oh * ow, (long long)ppw * (kThreads / 32)), (unsigned)batch);
  torgb_pool2_lanes_kernel<L><<<grid, kThreads, 0, stream>>>(static_cast<const uint4 *>(x), w, s, bias, skip, out, (int)oh,
                                                             (int)ow, (int)c, wscale, ppw);
  return check_launch("torgb_pool2_lanes_kernel");
}

static bool lanes_enabled() {
  static const bool on = getenv("VSP_NO_TORGB_LANES") == nullptr;
  return on;
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
  if (lanes_enabled()) {     // C = 32 (64-byte pixels) already streams at ~5.9 TB/s in the pixel-per-thread form
    switch (c) {
      case 64: return launch_torgb_lanes<8, 1>(x, w, s, bias, skip, out, batch, hw, c, wscale, stream);
      case 128: return launch_torgb_lanes<16, 1>(x, w, s, bias, skip, out, batch, hw, c, wscale, stream);
      case 256: return launch_torgb_lanes<32, 1>(x, w, s, bias, skip, out, batch, hw, c, wscale, stream);
      case 512: return launch_torgb_lanes<32, 2>(x, w, s, bias, skip, out, batch, hw, c, wscale, stream);
      case 1024: return launch_torgb_lanes<32, 4>(x, w, s, bias, skip, out, batch, hw, c, wscale, stream);
      default: break;
    }
  }
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
  VSP_REQUIRE(x && w && out, "torgb_pool2: null pointer");
  VSP_REQUIRE(batch <= 65535 && c <= 2048 && out_h < 32768 && out_w < 32768, "torgb_pool2: extent too large");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "torgb_pool2: x must be 16-byte aligned");
  PoolTaps t;
  for (int i = 0; i < 9; ++i) t.k[i] = k3_host ? k3_host[i] : 0.f;
  if (lanes_enabled() && (skip == nullptr || k3_host == nullptr)) {
    switch (c) {
      case 16: return launch_torgb_pool2_lanes<4>(x, w, s, bias, skip, out, batch, out_h, out_w, c, wscale, stream);
      case 32: return launch_torgb_pool2_lanes<8>(x, w, s, bias, skip, out, batch, out_h, out_w, c, wscale, stream);
      case 64: return launch_torgb_pool2_lanes<16>(x, w, s, bias, skip, out, batch, out_h, out_w, c, wscale, stream);
      case 128: return launch_torgb_pool2_lanes<32>(x, w, s, bias, skip, out, batch, out_h, out_w, c, wscale, stream);
      default: break;
    }
  }
  if (skip != nullptr && k3_host == nullptr) {   // pre-pooled skip on the pixel-per-thread kernel: centre tap only
    for (int i = 0; i < 9; ++i) t.k[i] = (i == 4) ? 1.f : 0.f;
  }
  dim3 grid((unsigned)ceil_div64(out_h * out_w, kThreads), (unsigned)batch);
  torgb_pool2_kernel<<<grid, kThreads, sizeof(float4) * c, stream>>>(static_cast<const uint4 *>(x), w, s, bias, skip, out,
                                                                    (int)out_h, (int)out_w, (int)c, wscale, t);
  return check_launch("torgb_pool2_kernel");
}
