// conv_common.cuh — pieces shared by the tcgen05 convolution kernels (conv_sm100.cu, conv_ring_sm100.cu):
// launch parameters, tile constants and the fused epilogue applied to one chunk of accumulator columns.
#pragma once
#include "common.cuh"

namespace vsp {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;              // bf16 elements = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumThreads = 320;         // 10 warps: TMA producer, MMA issuer, 8 epilogue (two per TMEM lane quadrant)
constexpr int kMaxTaps = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB

struct ConvParams {
  int batch, groups;
  int cin, cout;
  int out_h, out_w;        // logical output extent of this launch
  int stride;
  int ntaps;
  int tap_w[kMaxTaps], tap_dy[kMaxTaps], tap_dx[kMaxTaps];
  int tw, th, tiles_w, tiles_h, tiles_n;
  int tb, tiles_b;         // samples stacked in one 128-row tile (shared weights, images smaller than 128 pixels)
  int branch_mode;         // channel tile n_i is branch n_i of a SMART layer: tap offsets scaled by n_dil[n_i]
  int n_dil[4];
  long long total_tiles;
  int mpairs;              // CTA-pair kernel: pairs of spatial tiles per sample, pair-tiles in the launch
  long long total_pairs;
  int kc;                  // channel blocks per tap = ceil(cin / 64)
  int halo_d, halo_w;      // row-halo / row-ring kernels: dilation and padded halo row length (pixels)
  // row-ring kernel: R output rows per accumulator hand-off, S row slots, segments of L rows per chain
  int rr_R, rr_S, rr_L, rr_segs, rr_chains, rr_strips, rr_nb, rr_staged, rr_nslices;
  // output addressing: logical (oh, ow) -> (oh*os + oo_h, ow*os + oo_w) inside [full_h, full_w]
  void *out;
  int out_nhwc;
  int full_h, full_w, os, oo_h, oo_w;
  long long ldo, co_off;
  // pixel-shuffle epilogue (fused up-conv): GEMM column n = class * shuffle_cout + channel, class = (pa, pb) in
  // row-major order, written to (2*oh + pa, 2*ow + pb); 0 = off
  int shuffle_cout;
  // class mode (stride-2 transposed convolution as ONE launch): channel tile n_i belongs to output parity class
  // n_i / cls_tpc and runs only that class's taps — tap j of class c reads the activation at shift
  // (tap_dy, tap_dx)[cls_shift[c][j]] and weight tap cls_w[c][j], rows (n_i % cls_tpc) * BLOCK_N of the layer's OWN packed
  // weights (no composite tensor, no structural zeros: 1x the layer's FLOPs).  Tiles are dealt so that a CTA's successive
  // tiles cycle through the classes (their work is 4 : 2 : 2 : 1 taps).
  int cls_mode, cls_tpc;
  int cls_ntaps[4], cls_shift[4][4], cls_w[4][4];
  int staged;              // conv_fprop_kernel: epilogue through swizzled shared memory + TMA stores
  // epilogue
  const float *row_scale;
  const float *noise;
  long long noise_bstride;
  float noise_weight;
  const float *noise_weight_dev;
  const float *bias;
  const float *pre_bias;
  int pre_act;
  int act;
  float alpha, scale;
  const void *residual;
  const void *residual2;
  const float *alpha_vec;  // per-channel negative slope (PReLU) replacing `alpha` in the second stage
};

// channel-tile index of a persistent-loop tile (rotated in class mode so that one CTA sees every class in turn)
__device__ __forceinline__ int tile_n_index(const ConvParams &p, long long tile) {
  const int n_i = (int)(tile % p.tiles_n);
  return p.cls_mode ? (int)((n_i + tile / p.tiles_n) % p.tiles_n) : n_i;
}

__device__ __forceinline__ uint4 pack8_bf16(const float *v) {
  __nv_bfloat162 q0 = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 q1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 q2 = __floats2bfloat162_rn(v[4], v[5]);
  __nv_bfloat162 q3 = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t *>(&q0);
  u.y = *reinterpret_cast<uint32_t *>(&q1);
  u.z = *reinterpret_cast<uint32_t *>(&q2);
  u.w = *reinterpret_cast<uint32_t *>(&q3);
  return u;
}

__device__ __forceinline__ float epi_act(float v, int act, float alpha, float scale) {
  if (act == 3) v = (v > 0.f ? v : v * alpha) * scale;
  return v;
}

// Fused epilogue arithmetic of CHUNK accumulator columns of one pixel: TMEM -> registers -> demod, [bias1 + lrelu],
// noise + bias + lrelu, residuals.  The chunk covers output channels c0 .. c0+CHUNK-1 of a tensor with `ctot`
// channels (c0 differs from the GEMM column in the pixel-shuffle mapping); results in v[] (zeros outside the output).
// `vrs`/`vb1`/`vb2` point at this chunk's slice of the per-channel vectors staged in shared memory; `rs_row`
// (optional, global) replaces vrs with per-thread demod values when the tile stacks several samples.
// The tcgen05.ld is executed unconditionally (warp-collective).
template <int CHUNK>
__device__ __forceinline__ void epi_compute(const ConvParams &p, uint32_t taddr, int c0, int ctot, int b, long long pix,
                                            long long plane, bool pix_ok, float nz, const float *vrs,
                                            const float *vb1, const float *vb2, const float *rs_row,
                                            float (&v)[CHUNK]) {
  const bool live = pix_ok && c0 < ctot;
  const bool fullc = c0 + CHUNK <= ctot;
  // residual prefetch (independent of the accumulator)
  float rsd[CHUNK];
#pragma unroll
  for (int j = 0; j < CHUNK; ++j) rsd[j] = 0.f;
  if (live && (p.residual || p.residual2)) {
    if (!p.out_nhwc) {
      const long long off = ((long long)b * ctot + c0) * plane + pix;
      const float *r1 = static_cast<const float *>(p.residual);
      const float *r2 = static_cast<const float *>(p.residual2);
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (fullc || c0 + j < ctot) {
          if (r1) rsd[j] += __ldg(r1 + off + (long long)j * plane);
          if (r2) rsd[j] += __ldg(r2 + off + (long long)j * plane);
        }
    } else {
      const long long off = ((long long)b * plane + pix) * p.ldo + p.co_off + c0;
      const __nv_bfloat16 *rr[2] = {static_cast<const __nv_bfloat16 *>(p.residual),
                                    static_cast<const __nv_bfloat16 *>(p.residual2)};
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!rr[q]) continue;
        if (fullc && (((p.ldo | p.co_off) & 7) == 0)) {
#pragma unroll
          for (int j = 0; j < CHUNK; j += 8) {
            const uint4 u = __ldg(reinterpret_cast<const uint4 *>(rr[q] + off + j));
            const __nv_bfloat162 *h2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h2[e]);
              rsd[j + 2 * e] += f.x;
              rsd[j + 2 * e + 1] += f.y;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < CHUNK; ++j)
            if (c0 + j < ctot) rsd[j] += __bfloat162float(rr[q][off + j]);
        }
      }
    }
  }
  uint32_t r[CHUNK];
  if constexpr (CHUNK == 32) tmem_ld_32x32b_x32(taddr, r);
  else tmem_ld_32x32b_x16(taddr, reinterpret_cast<uint32_t(&)[16]>(r));
  tmem_ld_wait();
  const float4 *srs = reinterpret_cast<const float4 *>(vrs);
  const float4 *sb1 = reinterpret_cast<const float4 *>(vb1);
  const float4 *sb2 = reinterpret_cast<const float4 *>(vb2);
#pragma unroll
  for (int j = 0; j < CHUNK; j += 4) {
    const float4 a = srs[j / 4], c1 = sb1[j / 4], c2 = sb2[j / 4];
    float aa[4] = {a.x, a.y, a.z, a.w};
    const float b1[4] = {c1.x, c1.y, c1.z, c1.w}, b2[4] = {c2.x, c2.y, c2.z, c2.w};
    if (rs_row != nullptr) {
#pragma unroll
      for (int e = 0; e < 4; ++e) aa[e] = (live && (fullc || c0 + j + e < ctot)) ? __ldg(rs_row + j + e) : 1.f;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x = __uint_as_float(r[j + e]) * aa[e];
      if (p.pre_act) x = epi_act(x + b1[e], p.pre_act, p.alpha, p.scale);   // stage 1 (SMART fusion conv)
      const float al = (p.alpha_vec != nullptr && (fullc || c0 + j + e < ctot)) ? __ldg(p.alpha_vec + c0 + j + e) : p.alpha;
      x = epi_act(x + nz + b2[e], p.act, al, p.scale);                        // noise + bias + activation
      v[j + e] = live ? x + rsd[j + e] : 0.f;
    }
  }
}

// Direct per-thread stores of one computed chunk: NCHW fp32 (coalesced across the warp's pixels) or NHWC bf16.
template <int CHUNK>
__device__ __forceinline__ void epi_store_direct(const ConvParams &p, int c0, int ctot, int b, long long pix,
                                                 long long plane, bool pix_ok, const float (&v)[CHUNK]) {
  if (!(pix_ok && c0 < ctot)) return;
  const bool fullc = c0 + CHUNK <= ctot;
  if (!p.out_nhwc) {
    float *o = static_cast<float *>(p.out) + ((long long)b * ctot + c0) * plane + pix;
#pragma unroll
    for (int j = 0; j < CHUNK; ++j)
      if (fullc || c0 + j < ctot) o[(long long)j * plane] = v[j];
  } else {
    __nv_bfloat16 *o = static_cast<__nv_bfloat16 *>(p.out) + ((long long)b * plane + pix) * p.ldo + p.co_off + c0;
    if (fullc && (((p.ldo | p.co_off) & 7) == 0)) {
#pragma unroll
      for (int j = 0; j < CHUNK; j += 8) *reinterpret_cast<uint4 *>(o + j) = pack8_bf16(&v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j)
        if (c0 + j < ctot) o[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// Lean staged-epilogue arithmetic for 8 channels: demod, [bias1 + lrelu], noise + bias + lrelu, branch-free.
// lrelu(t) * s == max(t * s, t * s * a) for 0 <= a <= 1, s > 0, so the gains are folded into the per-channel vectors
// when they are staged (lean_scale_*): vrs = demod * (pre ? m1 : m2), vb1 = bias1 * m1, vb2 = bias * m2, and the
// caller passes nzs = noise * m2.  A disabled stage has m = a = 1 (max(t, t) = t).
//   no first stage : t = acc * vrs + (nzs + vb2);                      out = max(t, t * a2)        (4 ops / element)
//   first stage    : t1 = acc * vrs + vb1; y = max(t1, t1 * a1);  t2 = y * m2 + (nzs + vb2);  out = max(t2, t2 * a2)
struct LeanK {
  float m1, a1, m2, a2;
  int pre;
};
__device__ __forceinline__ LeanK lean_consts(const ConvParams &p) {
  LeanK k;
  k.pre = p.pre_act != 0;
  k.m1 = k.pre ? p.scale : 1.f;
  k.a1 = k.pre ? p.alpha : 1.f;
  k.m2 = p.act ? p.scale : 1.f;
  k.a2 = p.act ? p.alpha : 1.f;
  return k;
}
__device__ __forceinline__ float lean_scale_rs(const LeanK &k, float rs) { return rs * (k.pre ? k.m1 : k.m2); }
__device__ __forceinline__ float lean_scale_b1(const LeanK &k, float b1) { return b1 * k.m1; }
__device__ __forceinline__ float lean_scale_b2(const LeanK &k, float b2) { return b2 * k.m2; }

template <bool PRE>
__device__ __forceinline__ void epi_lean8f_t(const uint32_t *r, const float *vrs, const float *vb1, const float *vb2,
                                             float nzs, const LeanK &k, float (&v)[8]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 a = reinterpret_cast<const float4 *>(vrs)[h];
    const float4 c2 = reinterpret_cast<const float4 *>(vb2)[h];
    const float aa[4] = {a.x, a.y, a.z, a.w}, b2[4] = {c2.x, c2.y, c2.z, c2.w};
    if constexpr (PRE) {
      const float4 c1 = reinterpret_cast<const float4 *>(vb1)[h];
      const float b1[4] = {c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float t1 = fmaf(__uint_as_float(r[4 * h + e]), aa[e], b1[e]);
        const float y = fmaxf(t1, t1 * k.a1);
        const float t2 = fmaf(y, k.m2, nzs + b2[e]);
        v[4 * h + e] = fmaxf(t2, t2 * k.a2);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float t = fmaf(__uint_as_float(r[4 * h + e]), aa[e], nzs + b2[e]);
        v[4 * h + e] = fmaxf(t, t * k.a2);
      }
    }
  }
}
// PReLU form of the single-stage epilogue (per-channel slope, any sign): t = acc * vrs + (nzs + vb2); out = max(t, 0) + a min(t, 0)
__device__ __forceinline__ void epi_lean8f_prelu(const uint32_t *r, const float *vrs, const float *vb2, const float *va, float nzs,
                                                 float (&v)[8]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 a = reinterpret_cast<const float4 *>(vrs)[h];
    const float4 c2 = reinterpret_cast<const float4 *>(vb2)[h];
    const float4 s = reinterpret_cast<const float4 *>(va)[h];
    const float aa[4] = {a.x, a.y, a.z, a.w}, b2[4] = {c2.x, c2.y, c2.z, c2.w}, sl[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float t = fmaf(__uint_as_float(r[4 * h + e]), aa[e], nzs + b2[e]);
      v[4 * h + e] = fmaf(sl[e], fminf(t, 0.f), fmaxf(t, 0.f));
    }
  }
}
__device__ __forceinline__ void epi_lean8f(const uint32_t *r, const float *vrs, const float *vb1, const float *vb2,
                                           float nzs, const LeanK &k, float (&v)[8]) {
  if (k.pre) epi_lean8f_t<true>(r, vrs, vb1, vb2, nzs, k, v);
  else epi_lean8f_t<false>(r, vrs, vb1, vb2, nzs, k, v);
}
__device__ __forceinline__ uint4 epi_lean8(const uint32_t *r, const float *vrs, const float *vb1, const float *vb2,
                                           float nzs, const LeanK &k) {
  float v[8];
  epi_lean8f(r, vrs, vb1, vb2, nzs, k, v);
  return pack8_bf16(v);
}

// v[0..7] += a + b (packed bf16 residuals)
__device__ __forceinline__ void add2_bf16x8(float (&v)[8], const uint4 &a, const uint4 &b) {
  const __nv_bfloat162 *ha = reinterpret_cast<const __nv_bfloat162 *>(&a);
  const __nv_bfloat162 *hb = reinterpret_cast<const __nv_bfloat162 *>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __bfloat1622float2(ha[i]), fb = __bfloat1622float2(hb[i]);
    v[2 * i] += fa.x + fb.x;
    v[2 * i + 1] += fa.y + fb.y;
  }
}

// Output tensor maps of the staged (shared memory -> TMA store) epilogue: one per output parity class
// (plain convs use m[0]; a stride-2 class launch uses its own; the pixel-shuffle epilogue uses all four).
struct OutMaps {
  CUtensorMap m[4];
};

template <int CHUNK>
__device__ __forceinline__ void epi_chunk(const ConvParams &p, uint32_t taddr, int c0, int ctot, int b, long long pix,
                                          long long plane, bool pix_ok, float nz, const float *vrs,
                                          const float *vb1, const float *vb2, const float *rs_row) {
  float v[CHUNK];
  epi_compute<CHUNK>(p, taddr, c0, ctot, b, pix, plane, pix_ok, nz, vrs, vb1, vb2, rs_row, v);
  epi_store_direct<CHUNK>(p, c0, ctot, b, pix, plane, pix_ok, v);
}

// conv_ring_sm100.cu: returns -1 when the shape is not eligible (caller falls through to the other kernels),
// 0 on success, >0 on error.  `p` must already carry cin/cout/extent/output/epilogue fields.
int conv_ring_try_launch(ConvParams &p, const void *x, const void *wq, int64_t in_h, int64_t in_w,
                         int64_t cout_pad, int taps_total, int dil, cudaStream_t stream);

}  // namespace vsp
