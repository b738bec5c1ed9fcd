// bias_act_sm100.cu — fused bias + activation (+gain), its first-order backward
// with the bias-gradient reduction fused in, for sm_100a.
//
// Behaviour contract: op/fused_bias_act_kernel.cu:18-105 and
// op/fused_act.py:126-196 of the reference.  Design: pure streaming kernels,
// 128-bit non-allocating loads / streaming stores, one channel lookup per
// 128-bit vector (no per-element div/mod), grid sized to the SM count with a
// grid-stride loop; the backward reduces dbias with warp shuffles -> shared
// memory -> one fp32 atomic per (block, channel run).
#include "common.cuh"

namespace vsp {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float act_apply(float x, float ref, int mode, float alpha) {
  // mode = act*10 + grad, as the switch at op/fused_bias_act_kernel.cu:40-61
  switch (mode) {
    case 30: return x > 0.f ? x : x * alpha;
    case 31: return ref > 0.f ? x : x * alpha;
    case 32: return 0.f;
    case 12: return 0.f;
    default: return x;  // 10, 11 and anything else: linear
  }
}

// Vector path: n % 4 == 0, step_b % 4 == 0 (a float4 never straddles channels),
// all pointers 16-byte aligned.
template <bool HAS_B, bool HAS_REF>
__global__ void __launch_bounds__(kThreads)
bias_act_vec_kernel(const float4 *__restrict__ x, const float *__restrict__ b,
                    const float4 *__restrict__ ref, float4 *__restrict__ y, long long nvec,
                    long long step_v, int size_b, int mode, float alpha, float scale) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < nvec; i += stride) {
    float4 v = ld_stream_f4(x + i);
    if (HAS_B) {
      const float bb = __ldg(b + (int)((i / step_v) % size_b));
      v.x += bb; v.y += bb; v.z += bb; v.w += bb;
    }
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HAS_REF) r = ld_stream_f4(ref + i);
    float4 o;
    o.x = act_apply(v.x, r.x, mode, alpha) * scale;
    o.y = act_apply(v.y, r.y, mode, alpha) * scale;
    o.z = act_apply(v.z, r.z, mode, alpha) * scale;
    o.w = act_apply(v.w, r.w, mode, alpha) * scale;
    st_stream_f4(y + i, o);
  }
}

// Plane form of the forward: grid = (4096-element chunks of a plane, planes).  The bias of the block is ONE load (the flat
// kernel above spends a 64-bit divide per vector on finding its channel) and every thread has four independent 128-bit
// loads in flight before the first use.  step_b % 4 == 0, all pointers 16-byte aligned.
template <bool HAS_REF>
__global__ void __launch_bounds__(kThreads)
bias_act_plane_kernel(const float *__restrict__ x, const float *__restrict__ b, const float *__restrict__ ref,
                      float *__restrict__ y, long long step_b, int size_b, long long plane0, int mode, float alpha,
                      float scale) {
  const long long plane = plane0 + blockIdx.y;
  const float bb = b != nullptr ? __ldg(b + (int)(plane % size_b)) : 0.f;
  const long long base = plane * step_b;
  const long long nv = step_b >> 2;                                   // vectors in the plane
  const float4 *x4 = reinterpret_cast<const float4 *>(x + base);
  const float4 *r4 = reinterpret_cast<const float4 *>(ref + (HAS_REF ? base : 0));
  float4 *y4 = reinterpret_cast<float4 *>(y + base);
  const long long i0 = (long long)blockIdx.x * (kThreads * 4) + threadIdx.x;
  float4 v[4], r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = i0 + k * kThreads;
    if (i < nv) {
      v[k] = ld_stream_f4(x4 + i);
      if (HAS_REF) r[k] = ld_stream_f4(r4 + i);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = i0 + k * kThreads;
    if (i < nv) {
      const float4 rr = HAS_REF ? r[k] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o;
      o.x = act_apply(v[k].x + bb, rr.x, mode, alpha) * scale;
      o.y = act_apply(v[k].y + bb, rr.y, mode, alpha) * scale;
      o.z = act_apply(v[k].z + bb, rr.z, mode, alpha) * scale;
      o.w = act_apply(v[k].w + bb, rr.w, mode, alpha) * scale;
      st_stream_f4(y4 + i, o);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
bias_act_scalar_kernel(const float *__restrict__ x, const float *__restrict__ b,
                       const float *__restrict__ ref, float *__restrict__ y, long long n,
                       long long step_b, int size_b, int mode, float alpha, float scale) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n; i += stride) {
    float v = x[i];
    if (b != nullptr) v += __ldg(b + (int)((i / step_b) % size_b));
    const float r = ref != nullptr ? ref[i] : 0.f;
    y[i] = act_apply(v, r, mode, alpha) * scale;
  }
}

// Backward, large planes: grid = (chunks per plane, planes). One block reduces
// its chunk of dx and issues ONE atomicAdd on dbias[plane % size_b].
template <bool VEC>
__global__ void __launch_bounds__(kThreads)
bias_act_bwd_plane_kernel(const float *__restrict__ dy, const float *__restrict__ ref,
                          float *__restrict__ dx, float *__restrict__ dbias, long long step_b,
                          int size_b, long long chunk, float alpha, float scale) {
  const long long plane = blockIdx.y;
  const long long base = plane * step_b;
  const long long lo = blockIdx.x * chunk;
  const long long hi = min(lo + chunk, step_b);
  float acc = 0.f;
  if (VEC) {
    const float4 *dy4 = reinterpret_cast<const float4 *>(dy + base);
    const float4 *rf4 = reinterpret_cast<const float4 *>(ref + base);
    float4 *dx4 = reinterpret_cast<float4 *>(dx + base);
    for (long long i = lo / 4 + threadIdx.x; i < hi / 4; i += kThreads) {
      const float4 g = ld_stream_f4(dy4 + i);
      const float4 r = ld_stream_f4(rf4 + i);
      float4 o;
      o.x = (r.x > 0.f ? g.x : g.x * alpha) * scale;
      o.y = (r.y > 0.f ? g.y : g.y * alpha) * scale;
      o.z = (r.z > 0.f ? g.z : g.z * alpha) * scale;
      o.w = (r.w > 0.f ? g.w : g.w * alpha) * scale;
      st_stream_f4(dx4 + i, o);
      acc += (o.x + o.y) + (o.z + o.w);
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
      const float g = dy[base + i], r = ref[base + i];
      const float o = (r > 0.f ? g : g * alpha) * scale;
      dx[base + i] = o;
      acc += o;
    }
  }
  if (dbias == nullptr) return;
  __shared__ float part[kThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(dbias + (int)(plane % size_b), v);
  }
}

// Backward, small planes (step_b < 256, e.g. the [B, C] inputs of EqualLinear):
// one thread per (channel, s) column walks the batch; coalesced across columns.
__global__ void __launch_bounds__(kThreads)
bias_act_bwd_small_kernel(const float *__restrict__ dy, const float *__restrict__ ref,
                          float *__restrict__ dx, float *__restrict__ dbias, long long outer,
                          long long step_b, int size_b, float alpha, float scale) {
  const long long cols = step_b * size_b;
  const long long col = blockIdx.x * (long long)kThreads + threadIdx.x;
  if (col >= cols) return;
  float acc = 0.f;
  for (long long o = 0; o < outer; ++o) {
    const long long i = o * cols + col;
    const float g = dy[i], r = ref[i];
    const float v = (r > 0.f ? g : g * alpha) * scale;
    dx[i] = v;
    acc += v;
  }
  if (dbias != nullptr) {
    if (step_b == 1) dbias[col] = acc;          // exactly one thread per channel
    else atomicAdd(dbias + (int)(col / step_b), acc);
  }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace vsp

extern "C" int vsp_bias_act_f32(const float *x, const float *b, const float *ref, float *y,
                                int64_t n, int64_t step_b, int64_t size_b, int act, int grad,
                                float alpha, float scale, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(n >= 0, "bias_act: negative element count");
  if (n == 0) return 0;
  VSP_REQUIRE(x && y, "bias_act: null pointer");
  if (b != nullptr) VSP_REQUIRE(step_b >= 1 && size_b >= 1 && size_b < (1LL << 31), "bias_act: bad bias geometry");
  const int mode = act * 10 + grad;
  const bool vec = (n % 4 == 0) && (b == nullptr || step_b % 4 == 0) && aligned16(x) && aligned16(y) &&
                   (ref == nullptr || aligned16(ref));
  if (vec && b != nullptr && step_b >= 1024 && n % step_b == 0) {
    const long long planes = n / step_b;
    const unsigned cx = (unsigned)((step_b / 4 + kThreads * 4 - 1) / (kThreads * 4));
    for (long long p0 = 0; p0 < planes; p0 += 65535) {
      const unsigned np = (unsigned)(planes - p0 < 65535 ? planes - p0 : 65535);
      if (ref) bias_act_plane_kernel<true><<<dim3(cx, np), kThreads, 0, stream>>>(x, b, ref, y, step_b, (int)size_b, p0, mode, alpha, scale);
      else bias_act_plane_kernel<false><<<dim3(cx, np), kThreads, 0, stream>>>(x, b, ref, y, step_b, (int)size_b, p0, mode, alpha, scale);
      if (int rc = check_launch("bias_act_plane_kernel")) return rc;
    }
    return 0;
  }
  const long long work = vec ? n / 4 : n;
  long long blocks = (work + kThreads - 1) / kThreads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (vec) {
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    const float4 *r4 = reinterpret_cast<const float4 *>(ref);
    float4 *y4 = reinterpret_cast<float4 *>(y);
    const long long sv = b ? step_b / 4 : 1;
    const int sb = b ? (int)size_b : 1;
#define VSP_BA_LAUNCH(HB, HR) \
  bias_act_vec_kernel<HB, HR><<<(unsigned)blocks, kThreads, 0, stream>>>(x4, b, r4, y4, work, sv, sb, mode, alpha, scale)
    if (b && ref) VSP_BA_LAUNCH(true, true);
    else if (b) VSP_BA_LAUNCH(true, false);
    else if (ref) VSP_BA_LAUNCH(false, true);
    else VSP_BA_LAUNCH(false, false);
#undef VSP_BA_LAUNCH
    return check_launch("bias_act_vec_kernel");
  }
  bias_act_scalar_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(x, b, ref, y, n, b ? step_b : 1,
                                                                    b ? (int)size_b : 1, mode, alpha, scale);
  return check_launch("bias_act_scalar_kernel");
}

extern "C" int vsp_bias_act_bwd_f32(const float *dy, const float *ref, float *dx, float *dbias,
                                    int64_t n, int64_t step_b, int64_t size_b, float alpha,
                                    float scale, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(n >= 0 && step_b >= 1 && size_b >= 1 && size_b < (1LL << 31), "bias_act_bwd: bad geometry");
  if (dbias != nullptr) VSP_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * size_b, stream));
  if (n == 0) return 0;
  VSP_REQUIRE(dy && ref && dx, "bias_act_bwd: null pointer");
  VSP_REQUIRE(n % (step_b * size_b) == 0, "bias_act_bwd: n (%lld) is not a multiple of step_b*size_b (%lld)",
              (long long)n, (long long)(step_b * size_b));
  const long long outer = n / (step_b * size_b);
  if (step_b >= 256) {
    const long long planes = outer * size_b;
    const bool vec = (step_b % 4 == 0) && aligned16(dy) && aligned16(ref) && aligned16(dx);
    // chunk: multiple of 4 elements, >= 4096, sized so the grid has ~16 blocks per SM
    long long want = (long long)num_sms() * 16;
    long long per_plane = (want + planes - 1) / planes;
    if (per_plane < 1) per_plane = 1;
    long long chunk = (step_b + per_plane - 1) / per_plane;
    if (chunk < 4096) chunk = 4096;
    chunk = (chunk + 3) & ~3LL;
    const long long cx = (step_b + chunk - 1) / chunk;
    VSP_REQUIRE(planes <= 65535 * 32768LL, "bias_act_bwd: too many planes");
    if (planes <= 65535) {
      dim3 grid((unsigned)cx, (unsigned)planes);
      if (vec) bias_act_bwd_plane_kernel<true><<<grid, kThreads, 0, stream>>>(dy, ref, dx, dbias, step_b, (int)size_b, chunk, alpha, scale);
      else bias_act_bwd_plane_kernel<false><<<grid, kThreads, 0, stream>>>(dy, ref, dx, dbias, step_b, (int)size_b, chunk, alpha, scale);
      return check_launch("bias_act_bwd_plane_kernel");
    }
    // more planes than gridDim.y allows: walk them in slabs that keep the channel phase
    const long long slab = (65535 / size_b) * size_b > 0 ? (65535 / size_b) * size_b : 0;
    VSP_REQUIRE(slab > 0, "bias_act_bwd: size_b too large for the plane kernel");
    for (long long p0 = 0; p0 < planes; p0 += slab) {
      const long long np = planes - p0 < slab ? planes - p0 : slab;
      dim3 grid((unsigned)cx, (unsigned)np);
      const long long off = p0 * step_b;
      if (vec) bias_act_bwd_plane_kernel<true><<<grid, kThreads, 0, stream>>>(dy + off, ref + off, dx + off, dbias, step_b, (int)size_b, chunk, alpha, scale);
      else bias_act_bwd_plane_kernel<false><<<grid, kThreads, 0, stream>>>(dy + off, ref + off, dx + off, dbias, step_b, (int)size_b, chunk, alpha, scale);
      if (int rc = check_launch("bias_act_bwd_plane_kernel")) return rc;
    }
    return 0;
  }
  const long long cols = step_b * size_b;
  const long long blocks = (cols + kThreads - 1) / kThreads;
  bias_act_bwd_small_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(dy, ref, dx, dbias, outer, step_b,
                                                                       (int)size_b, alpha, scale);
  return check_launch("bias_act_bwd_small_kernel");
}
