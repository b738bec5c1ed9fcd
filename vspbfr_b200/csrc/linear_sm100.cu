// linear_sm100.cu — grouped EqualLinear: every style-modulation linear of a network pass in ONE launch.
//
// Reference: EqualLinear (models/RestoreNet.py:142-176) as used by ModulatedConv2d.modulation / SMART_layer.modulation
// (:467,:510,:211,:227): s_j = F.linear(style_j, W_j * scale_j, bias_j * lr_mul_j), one tiny [B,512..2048] x
// [Cin, 512..2048] product per layer — 60+ library launches per forward in the reference.  All styles of a pass are
// known up front, so the host builds one descriptor table and one launch computes every output row (problem j,
// channel o).  Two kernels: micro-batches <= 8 use one warp per row (the weight row streamed once with 128-bit loads and
// dotted with the staged style rows), 9..32 samples the lane-per-sample split-K form further down.  With `act` set in the
// descriptor the same launch is a layer of the style MLP (bias * lr_mul + leaky relu * sqrt 2 in the epilogue).
// Memory-bound on the weights (HBM/L2), fp32 throughout.
#include "common.cuh"

namespace vsp {
namespace {

__device__ __forceinline__ float lin_act(const vsp_linear_desc &d, float v) {
  return d.act == 3 ? (v > 0.f ? v : v * d.alpha) * d.gain : v;
}

constexpr int kLinThreads = 256;
constexpr int kLinRows = 8;     // output rows per block (one per warp)
// kLinB samples are accumulated per pass over kLinK staged style elements (kLinB x kLinK floats of shared memory);
// instantiated as 8 x 512 (batch > 8 goes to grouped_linear_lanes_kernel: with 8 samples per pass a 32-face micro-batch
// re-streamed every weight row and re-staged the styles four times, 0.37 TB/s on the weights).

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
      if (live && lane == 0 && b0 + q < batch) y[d.y_off + (long long)(b0 + q) * d.out_dim + o] = lin_act(d, v * d.wscale + bias);
    }
  }
}

// Micro-batches of 9..32 samples: lane = sample.  The 32 x kLinK accumulate-per-lane form above issues one 128-bit
// shared-memory load per FMA quadruple and sample (ncu launch list of a 32-face micro-batch: 6 launches, 525 us at
// 0.29 TB/s on the weights).  Here a block owns up to kLanGroups consecutive 8-row groups of ONE problem: the style slab
// [32 samples x kLanK] is staged once per K slab and shared by the groups (every block of a problem reads the same style
// rows — re-staging them per 8 rows made the launch wait on a few L2 lines), the 8 warps split the slab between them
// (split-K), each lane keeps 8 row accumulators per group for ITS sample, and the weights are read as shared-memory
// broadcasts: 12 shared-memory wavefronts per 32 FMA instructions instead of 128.  Operands are fetched into registers
// one stage ahead (styles: one slab, weights: one group), so their latency runs under the FMAs.  Style rows are staged
// with a pitch of kLanK + 4 floats so the per-lane 128-bit reads of a quarter warp fall into distinct banks.
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

template <int kLanGroups, int kMinBlocks>
__global__ void __launch_bounds__(kLinThreads, kMinBlocks)
grouped_linear_lanes_kernel(const vsp_linear_desc *__restrict__ descs, const int *__restrict__ row_start, int n_problems,
                            int total_rows, const float *__restrict__ x, long long x_bstride_call, float *__restrict__ y,
                            int batch) {
  __shared__ __align__(16) float xs[kLanB][kLanPitch];
  __shared__ __align__(16) float ws[kLinRows][kLanK];
  static_assert(sizeof(xs) >= sizeof(float) * kLanGroups * (kLinThreads / 32) * kLinRows * kLanB,
                "reduction buffer aliases xs");
  constexpr int kQuads = kLanK / 4;                              // float4 columns of a slab
  constexpr int kXPer = kLanB * kQuads / kLinThreads;            // 8 style quads per thread and slab
  constexpr int kWPer = kLinRows * kQuads / kLinThreads;         // 2 weight quads per thread, group and slab
  constexpr int kWarpK = kLanK / (kLinThreads / 32);             // 32 k per warp and slab
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int g_done = 0; g_done < kLanGroups;) {
    const int row0 = (blockIdx.x * kLanGroups + g_done) * kLinRows;
    if (row0 >= total_rows) break;
    int lo = 0, hi = n_problems - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(row_start + mid) <= row0) lo = mid; else hi = mid - 1;
    }
    const vsp_linear_desc d = descs[lo];
    const int o0 = row0 - __ldg(row_start + lo);
    const int prob_end = lo + 1 < n_problems ? __ldg(row_start + lo + 1) : total_rows;
    const int run = min(kLanGroups - g_done, (prob_end - row0) / kLinRows);   // groups of this problem in this block
    g_done += run;
    const float *xb = x + d.x_off;
    const long long x_bstride = d.x_bstride != 0 ? d.x_bstride : x_bstride_call;
    // every row of both operands starts 16-byte aligned and holds whole quads: branch-free 128-bit loads, all of a
    // stage's loads of a thread in flight together (the guarded form serialises them behind its branches)
    const bool vec = ((reinterpret_cast<uintptr_t>(xb) | reinterpret_cast<uintptr_t>(d.w)) & 15) == 0 &&
                     (x_bstride & 3) == 0 && (d.in_dim & 3) == 0;
    float4 xr[kXPer], wr[kWPer];
    auto fetch_x = [&](int b0, int k0) {                         // style slab (b0, k0) -> registers
      const int kn = min(kLanK, d.in_dim - k0);
#pragma unroll
      for (int j = 0; j < kXPer; ++j) {
        const int i = threadIdx.x + j * kLinThreads, q = i / kQuads, k = (i % kQuads) * 4;
        const bool ok = b0 + q < batch;
        const float *p = xb + (long long)(ok ? b0 + q : 0) * x_bstride + k0;
        if (vec) xr[j] = (ok && k < kn) ? __ldg(reinterpret_cast<const float4 *>(p + k)) : zero4;
        else xr[j] = ld4_guarded(p, k, kn, ok);
      }
    };
    auto fetch_w = [&](int g, int k0) {                          // weight rows of group g, slab k0 -> registers
      const int kn = min(kLanK, d.in_dim - k0);
#pragma unroll
      for (int j = 0; j < kWPer; ++j) {
        const int i = threadIdx.x + j * kLinThreads, r = i / kQuads, k = (i % kQuads) * 4;
        const int o = o0 + g * kLinRows + r;
        const bool live = o < d.out_dim;
        const float *p = d.w + (long long)(live ? o : 0) * d.in_dim + k0;
        if (vec) wr[j] = (live && k < kn) ? __ldg(reinterpret_cast<const float4 *>(p + k)) : zero4;
        else wr[j] = ld4_guarded(p, k, kn, live);
      }
    };
    for (int b0 = 0; b0 < batch; b0 += kLanB) {
      float acc[kLanGroups][kLinRows];
#pragma unroll
      for (int g = 0; g < kLanGroups; ++g)
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) acc[g][r] = 0.f;
      fetch_x(b0, 0);
      fetch_w(0, 0);
      for (int k0 = 0; k0 < d.in_dim; k0 += kLanK) {
        const int kn = min(kLanK, d.in_dim - k0);
        const bool more = k0 + kLanK < d.in_dim;
#pragma unroll
        for (int g = 0; g < kLanGroups; ++g) {
          if (g >= run) break;                                   // block-uniform
          __syncthreads();                                       // the previous stage (or reduction) has been read
          if (g == 0) {
#pragma unroll
            for (int j = 0; j < kXPer; ++j) {
              const int i = threadIdx.x + j * kLinThreads;
              *reinterpret_cast<float4 *>(&xs[i / kQuads][(i % kQuads) * 4]) = xr[j];
            }
          }
#pragma unroll
          for (int j = 0; j < kWPer; ++j) {
            const int i = threadIdx.x + j * kLinThreads;
            *reinterpret_cast<float4 *>(&ws[i / kQuads][(i % kQuads) * 4]) = wr[j];
          }
          __syncthreads();
          if (g == 0 && more) fetch_x(b0, k0 + kLanK);           // the next stages' loads fly under this stage's FMAs
          if (g + 1 < run) fetch_w(g + 1, k0);
          else if (more) fetch_w(0, k0 + kLanK);
          const int kw = warp * kWarpK;
#pragma unroll 2
          for (int k = kw; k < kw + kWarpK; k += 4) {
            if (k >= kn) break;                                  // warp-uniform
            const float4 xv = *reinterpret_cast<const float4 *>(&xs[lane][k]);
#pragma unroll
            for (int r = 0; r < kLinRows; ++r) {
              const float4 wv = *reinterpret_cast<const float4 *>(&ws[r][k]);
              acc[g][r] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[g][r]))));
            }
          }
        }
      }
      __syncthreads();
      float *red = &xs[0][0];                                    // [group][warp][row][sample]
#pragma unroll
      for (int g = 0; g < kLanGroups; ++g)
#pragma unroll
        for (int r = 0; r < kLinRows; ++r) red[((g * (kLinThreads / 32) + warp) * kLinRows + r) * kLanB + lane] = acc[g][r];
      __syncthreads();
      for (int g = 0; g < run; ++g) {
        const int r = warp, o = o0 + g * kLinRows + r;   // kLinThreads / 32 == kLinRows: one output row per warp, lane = sample
        float v = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kLinThreads / 32; ++w2) v += red[((g * (kLinThreads / 32) + w2) * kLinRows + r) * kLanB + lane];
        if (o < d.out_dim && b0 + lane < batch) {
          const float bias = d.bias ? __ldg(d.bias + o) * d.bscale : 0.f;
          y[d.y_off + (long long)(b0 + lane) * d.out_dim + o] = lin_act(d, v * d.wscale + bias);
        }
      }
    }
    __syncthreads();                                             // the reduction buffer is the next run's style slab
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
  {
    // four 8-row groups per block (one block per SM, 255 registers) when that still gives every SM four blocks, else one
    // group per block at two blocks per SM (the style MLP's 512-row layers); VSP_LINEAR_GROUPS=1|4 forces either
    static const int forced = getenv("VSP_LINEAR_GROUPS") ? atoi(getenv("VSP_LINEAR_GROUPS")) : 0;
    const int groups = forced ? forced : (blocks >= 4u * (unsigned)num_sms() ? 4 : 1);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (groups >= 4)
      grouped_linear_lanes_kernel<4, 1><<<(blocks + 3) / 4, kLinThreads, 0, st>>>(
          descs_dev, row_start_dev, n_problems, total_rows, x, x_bstride, y, batch);
    else
      grouped_linear_lanes_kernel<1, 2><<<blocks, kLinThreads, 0, st>>>(
          descs_dev, row_start_dev, n_problems, total_rows, x, x_bstride, y, batch);
  }
  return check_launch("grouped_linear_kernel");
}
