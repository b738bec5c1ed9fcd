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
constexpr int kLinB = 8;   // samples accumulated per pass over a weight row

__global__ void __launch_bounds__(kLinThreads)
grouped_linear_kernel(const vsp_linear_desc *__restrict__ descs, const int *__restrict__ row_start, int n_problems,
                      int total_rows, const float *__restrict__ x, long long x_bstride, float *__restrict__ y, int batch) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (kLinThreads / 32) + (threadIdx.x >> 5);
  if (row >= total_rows) return;
  // problem of this row: largest j with row_start[j] <= row
  int lo = 0, hi = n_problems - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(row_start + mid) <= row) lo = mid; else hi = mid - 1;
  }
  const vsp_linear_desc d = descs[lo];
  const int o = row - __ldg(row_start + lo);
  const float *w = d.w + (long long)o * d.in_dim;
  const float *xb = x + d.x_off;
  const float bias = d.bias ? __ldg(d.bias + o) * d.bscale : 0.f;
  for (int b0 = 0; b0 < batch; b0 += kLinB) {
    float acc[kLinB];
#pragma unroll
    for (int q = 0; q < kLinB; ++q) acc[q] = 0.f;
    if ((d.in_dim & 3) == 0) {
      for (int i = lane * 4; i < d.in_dim; i += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + i));
#pragma unroll
        for (int q = 0; q < kLinB; ++q)
          if (b0 + q < batch) {
            const float4 xv = __ldg(reinterpret_cast<const float4 *>(xb + (long long)(b0 + q) * x_bstride + i));
            acc[q] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[q]))));
          }
      }
    } else {
      for (int i = lane; i < d.in_dim; i += 32) {
        const float wv = __ldg(w + i);
#pragma unroll
        for (int q = 0; q < kLinB; ++q)
          if (b0 + q < batch) acc[q] = fmaf(wv, __ldg(xb + (long long)(b0 + q) * x_bstride + i), acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < kLinB; ++q) {
      const float v = warp_sum(acc[q]);
      if (lane == 0 && b0 + q < batch) y[d.y_off + (long long)(b0 + q) * d.out_dim + o] = v * d.wscale + bias;
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
  const int rows_per_block = kLinThreads / 32;
  const unsigned blocks = (unsigned)((total_rows + rows_per_block - 1) / rows_per_block);
  grouped_linear_kernel<<<blocks, kLinThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      descs_dev, row_start_dev, n_problems, total_rows, x, x_bstride, y, batch);
  return check_launch("grouped_linear_kernel");
}
