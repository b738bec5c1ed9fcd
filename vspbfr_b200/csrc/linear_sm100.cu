// linear_sm100.cu — grouped EqualLinear: every style-modulation linear of a network pass in ONE launch.
//
// Reference: EqualLinear (models/RestoreNet.py:142-176) as used by ModulatedConv2d.modulation / SMART_layer.modulation
// (:467,:510,:211,:227): s_j = F.linear(style_j, W_j * scale_j, bias_j * lr_mul_j), one tiny [B,512..2048] x
// [Cin, 512..2048] product per layer — 60+ library launches per forward in the reference.  All styles of a pass are
// known up front, so the host builds one descriptor table and this kernel computes every output row (problem j,
// channel o) with one warp: the weight row is streamed once (128-bit loads) and dotted with the B style rows.
// Memory-bound on the weights (HBM/L2), fp32 throughout.
#include "common.cuh"

namespace vsp {
namespace {

constexpr int kLinThreads = 256;
constexpr int kLinRows = 8;     // output rows per block (one per warp)
// kLinB samples are accumulated per pass over kLinK staged style elements (kLinB x kLinK floats of shared memory).
// Two instantiations: 8 x 512 for small batches, 32 x 256 (32 KB) for micro-batches > 8 — a 32-face micro-batch then streams every
// weight row ONCE (with 8 samples per pass it was re-streamed and the styles re-staged four times: 0.37 TB/s on the weights).

// One block = kLinRows consecutive output rows of ONE problem (the host pads every problem to a multiple of
// kLinRows rows in `row_start`), so the block's warps share the problem's style rows: they are staged in shared
// memory in [kLinB x kLinK] slabs (coalesced) and every warp dots its own weight row (streamed once per slab of
// samples, 128-bit loads) against them.
template <int kLinB, int kLinK>
__global__ void __launch_bounds__(kLinThreads)
grouped_linear_kernel(const vsp_linear_desc *__restrict__ descs, const int *__restrict__ row_start, int n_problems,
                      const float *__restrict__ x, long long x_bstride, float *__restrict__ y, int batch) {
  __shared__ __align__(16) float xs[kLinB][kLinK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * kLinRows;
  int lo = 0, hi = n_problems - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(row_start + mid) <= row0) lo = mid; else hi = mid - 1;
  }
  const vsp_linear_desc d = descs[lo];
  const int o = row0 - __ldg(row_start + lo) + warp;
  const bool live = o < d.out_dim;
  const float *w = d.w + (long long)(live ? o : 0) * d.in_dim;
  const float *xb = x + d.x_off;
  if (d.x_bstride != 0) x_bstride = d.x_bstride;
  const float bias = (live && d.bias) ? __ldg(d.bias + o) * d.bscale : 0.f;
  for (int b0 = 0; b0 < batch; b0 += kLinB) {
    float acc[kLinB];
#pragma unroll
    for (int q = 0; q < kLinB; ++q) acc[q] = 0.f;
    for (int k0 = 0; k0 < d.in_dim; k0 += kLinK) {
      const int kn = min(kLinK, d.in_dim - k0);
      __syncthreads();
      for (int i = threadIdx.x; i < kLinB * kLinK; i += kLinThreads) {
        const int q = i / kLinK, k = i % kLinK;
        xs[q][k] = (b0 + q < batch && k < kn) ? __ldg(xb + (long long)(b0 + q) * x_bstride + k0 + k) : 0.f;
      }
      __syncthreads();
      if (live) {
        for (int k = lane * 4; k < kn; k += 128) {
          float4 wv;
          if (k + 3 < kn && ((reinterpret_cast<uintptr_t>(w + k0 + k) & 15) == 0)) {
            wv = __ldg(reinterpret_cast<const float4 *>(w + k0 + k));
          } else {
            wv.x = __ldg(w + k0 + k);
            wv.y = k + 1 < kn ? __ldg(w + k0 + k + 1) : 0.f;
            wv.z = k + 2 < kn ? __ldg(w + k0 + k + 2) : 0.f;
            wv.w = k + 3 < kn ? __ldg(w + k0 + k + 3) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < kLinB; ++q) {
            const float4 xv = *reinterpret_cast<const float4 *>(&xs[q][k]);
            acc[q] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[q]))));
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kLinB; ++q) {
      const float v = warp_sum(acc[q]);
      if (live && lane == 0 && b0 + q < batch) y[d.y_off + (long long)(b0 + q) * d.out_dim + o] = v * d.wscale + bias;
    }
  }
}

// Micro-batches of 9..32 samples: lane = sample.  The 32 x kLinK accumulate-per-lane form above issues one 128-bit
// shared-memory load per FMA quadruple and sample (ncu launch list of a 32-face micro-batch: 6 launches, 525 us at
// 0.29 TB/s on the weights); here a block still owns 8 rows of one problem, but its 8 warps split every K slab between
// them (split-K), each lane keeps 8 row accumulators for ITS sample, and the weights are read as shared-memory
// broadcasts: 12 shared-memory wavefronts per 32 FMA instructions instead of 128.  Style rows are staged with a pitch
// of kLanK + 4 floats so the per-lane 128-bit reads of a quarter warp fall into distinct banks.
constexpr int kLanK = 256, kLanPitch = kLanK + 4, kLanB = 32;

__device__ __forceinline__ float4 ld4_guarded(const float *p, int k, int kn, bool ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!ok || k >= kn) return v;
  if (k + 3 < kn && ((reinterpret_cast<uintptr_t>(p + k) & 15) == 0)) return __ldg(reinterpret_cast<const float4 *>(p + k));
  v.x = __ldg(p + k);
  if (k + 1 < kn) v.y = __ldg(p + k + 1);
  if (k + 2 < kn) v.z = __ldg(p + k + 2);
  if (k + 3 < kn) v.w = __ldg(p + k + 3);
  return v;
}

__global__ void __launch_bounds__(kLinThreads)
grouped_linear_lanes_kernel(const vsp_linear_desc *__restrict__ descs, const int *__restrict__ row_start, int n_problems,
                            const float *__restrict__ x, long long x_bstride, float *__restrict__ y, int batch) {
  __shared__ __align__(16) float xs[kLanB][kLanPitch];
  __shared__ __align__(16) float ws[kLinRows][kLanK];
  static_assert(sizeof(xs) >= sizeof(float) * (kLinThreads / 32) * kLinRows * kLanB, "reduction buffer aliases xs");
  constexpr int kQuads = kLanK / 4;                              // float4 columns of a slab
  constexpr int kXPer = kLanB * kQuads / kLinThreads;            // 8 style quads per thread and slab
  constexpr int kWPer = kLinRows * kQuads / kLinThreads;         // 2 weight quads
  constexpr int kWarpK = kLanK / (kLinThreads / 32);             // 32 k per warp and slab
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * kLinRows;
  int lo = 0, hi = n_problems - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(row_start + mid) <= row0) lo = mid; else hi = mid - 1;
  }
  const vsp_linear_desc d = descs[lo];
  const int o0 = row0 - __ldg(row_start + lo);
  const float *xb = x + d.x_off;
  if (d.x_bstride != 0) x_bstride = d.x_bstride;
  // every row of both operands starts 16-byte aligned and holds whole quads: branch-free 128-bit loads, all of a slab's
  // loads of a thread in flight together (the guarded form below serialises them behind its branches)
  const bool vec = ((reinterpret_cast<uintptr_t>(xb) | reinterpret_cast<uintptr_t>(d.w)) & 15) == 0 &&
                   (x_bstride & 3) == 0 && (d.in_dim & 3) == 0;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 xr[kXPer], wr[kWPer];
  auto fetch = [&](int b0, int k0) {                             // slab (b0, k0) -> registers
    const int kn = min(kLanK, d.in_dim - k0);
#pragma unroll
    for (int j = 0; j < kXPer; ++j) {
      const int i = threadIdx.x + j * kLinThreads, q = i / kQuads, k = (i % kQuads) * 4;
      const bool ok = b0 + q < batch;
      const float *p = xb + (long long)(ok ? b0 + q : 0) * x_bstride + k0;
      if (vec) xr[j] = (ok && k < kn) ? __ldg(reinterpret_cast<const float4 *>(p + k)) : zero4;
      else xr[j] = ld4_guarded(p, k, kn, ok);
    }
#pragma unroll
    for (int j = 0; j < kWPer; ++j) {
      const int i = threadIdx.x + j * kLinThreads, r = i / kQuads, k = (i % kQuads) * 4;
      const bool live = o0 + r < d.out_dim;
      const float *p = d.w + (long long)(live ? o0 + r : 0) * d.in_dim + k0;
      if (vec) wr[j] = (live && k < kn) ? __ldg(reinterpret_cast<const float4 *>(p + k)) : zero4;
      else wr[j] = ld4_guarded(p, k, kn, live);
    }
  };
  for (int b0 = 0; b0 < batch; b0 += kLanB) {
    float acc[kLinRows];
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) acc[r] = 0.f;
    fetch(b0, 0);
    for (int k0 = 0; k0 < d.in_dim; k0 += kLanK) {
      const int kn = min(kLanK, d.in_dim - k0);
      __syncthreads();                                           // the previous slab (or reduction) has been read
#pragma unroll
      for (int j = 0; j < kXPer; ++j) {
        const int i = threadIdx.x + j * kLinThreads;
        *reinterpret_cast<float4 *>(&xs[i / kQuads][(i % kQuads) * 4]) = xr[j];
      }
#pragma unroll
      for (int j = 0; j < kWPer; ++j) {
        const int i = threadIdx.x + j * kLinThreads;
        *reinterpret_cast<float4 *>(&ws[i / kQuads][(i % kQuads) * 4]) = wr[j];
      }
      __syncthreads();
      if (k0 + kLanK < d.in_dim) fetch(b0, k0 + kLanK);          // next slab's loads fly under this slab's FMAs
      const int kw = warp * kWarpK;
#pragma unroll 2
      for (int k = kw; k < kw + kWarpK; k += 4) {
        if (k >= kn) break;                                      // warp-uniform
        const float4 xv = *reinterpret_cast<const float4 *>(&xs[lane][k]);
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) {
          const float4 wv = *reinterpret_cast<const float4 *>(&ws[r][k]);
          acc[r] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[r]))));
        }
      }
    }
    __syncthreads();
    float *red = &xs[0][0];                                      // [warp][row][sample]
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) red[(warp * kLinRows + r) * kLanB + lane] = acc[r];
    __syncthreads();
    {
      const int r = warp, o = o0 + r;            // kLinThreads / 32 == kLinRows: one output row per warp, lane = sample
      float v = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kLinThreads / 32; ++w2) v += red[(w2 * kLinRows + r) * kLanB + lane];
      if (o < d.out_dim && b0 + lane < batch) {
        const float bias = d.bias ? __ldg(d.bias + o) * d.bscale : 0.f;
        y[d.y_off + (long long)(b0 + lane) * d.out_dim + o] = v * d.wscale + bias;
      }
    }
  }
}

}  // namespace
}  // namespace vsp

extern "C" int vsp_grouped_linear_f32(const vsp_linear_desc *descs_dev, const int *row_start_dev, int n_problems,
                                      int total_rows, const float *x, int64_t x_bstride, float *y, int batch,
                                      void *stream_) {
  using namespace vsp;
  VSP_REQUIRE(n_problems >= 0 && total_rows >= 0 && batch >= 0, "grouped_linear: negative extent");
  if (n_problems == 0 || total_rows == 0 || batch == 0) return 0;
  VSP_REQUIRE(descs_dev && row_start_dev && x && y, "grouped_linear: null pointer");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_bstride & 3) == 0,
              "grouped_linear: x must be 16-byte aligned with a row stride that is a multiple of 4 floats");
  VSP_REQUIRE(total_rows % kLinRows == 0, "grouped_linear: row_start must pad every problem to a multiple of 8 rows");
  const unsigned blocks = (unsigned)(total_rows / kLinRows);
  if (batch <= 8)
    grouped_linear_kernel<8, 512><<<blocks, kLinThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
        descs_dev, row_start_dev, n_problems, x, x_bstride, y, batch);
  else
    grouped_linear_lanes_kernel<<<blocks, kLinThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
        descs_dev, row_start_dev, n_problems, x, x_bstride, y, batch);
  return check_launch("grouped_linear_kernel");
}
