// upfirdn2d_sm100.cu — up-sample / FIR / down-sample of fp32 planes for sm_100a.
//
// Behaviour contract: op/upfirdn2d.py:346-406 of the reference (zero-stuff by
// `up`, pad / crop, true convolution with `filt`, keep every `down`-th sample).
// Design (NOT the reference's 2019 tiled kernel):
//   * polyphase, phases resolved at compile time: every thread owns a 2x4
//     micro-tile of outputs whose tap->input mapping is a template constant, so
//     there is no per-element div/mod and zero-stuffed taps are never visited;
//   * input tiles (+halo) for PZ planes are staged in shared memory either by
//     one TMA box per stage (cp.async.bulk.tensor, OOB zero fill == padding,
//     2-stage mbarrier ring across plane groups) when the row pitch is 16-byte
//     aligned, or by coalesced LDGs otherwise (odd widths such as 65, 513);
//   * micro-tile windows are read with 64/128-bit LDS, outputs leave as 128-bit
//     streaming stores; optional bias + leaky-ReLU epilogue.
// A generic one-thread-per-output kernel covers every other parameter set
// (tuple factors, 12-tap filters, ... used by non_leaking.py:879-905).
#include "common.cuh"

#include <stdlib.h>

namespace vsp {
namespace {

constexpr int kThreads = 256;
constexpr int kK = 4;   // fast path filter extent (smaller filters are zero-extended)
// (output rows per thread follow from the tile: Cfg::RO = 2 for the 2048-output tiles, 4 for the 4096-output one)
constexpr int kCO = 4;  // outputs per thread, cols

struct UfdParams {
  const float *x;
  const float *filt;
  float *y;
  const float *bias;
  long long major;
  int in_h, in_w, out_h, out_w;
  int kh, kw;
  int up_x, up_y, down_x, down_y;
  int pad_x0, pad_y0;
  int channels, act;
  float alpha, scale;
  int tiles_x, tiles_y, zgroups, iters;  // fast path decomposition
  int use_tma;
};

__host__ __device__ constexpr int round_up4(int v) { return (v + 3) & ~3; }

template <int U, int D, int QX, int QY, int TOW, int TOH, int PZ>
struct Cfg {
  static constexpr int TX = TOW / kCO, TY = kThreads / (TX * PZ);
  static constexpr int RO = TOH / TY;                              // output rows per thread
  static_assert(TX * TY * PZ == kThreads && RO * TY == TOH && RO >= 1, "tile must map onto 256 threads");
  static_assert((RO * D) % U == 0 && (kCO * D) % U == 0, "micro tile must keep the phase");
  static constexpr int SY = RO * D / U, SX = kCO * D / U;        // thread stride in input samples
  static constexpr int WR = (QY + (RO - 1) * D + kK - 1) / U + 1;   // window rows
  static constexpr int WC = (QX + (kCO - 1) * D + kK - 1) / U + 1;  // window cols
  static constexpr int V = (SX % 4 == 0) ? 4 : 2;                 // LDS vector width (floats)
  // TMA needs the innermost box coordinate 16-byte aligned (measured on B200: a misaligned
  // start coordinate raises "illegal instruction"), so the smem tile starts at the input
  // column rounded DOWN to a multiple of 4 and every window is read at a shift XS in 0..3.
  static constexpr int WCV_MAX = ((3 + WC + V - 1) / V) * V;
  static constexpr int TIH = (TY - 1) * SY + WR;
  static constexpr int TIW = round_up4((TX - 1) * SX + WCV_MAX);
  static constexpr int TILE_FLOATS = PZ * TIH * TIW;               // what one TMA box delivers
  static constexpr int TILE_BYTES = TILE_FLOATS * 4;
  static constexpr int STAGE_FLOATS = (TILE_FLOATS + 31) & ~31;   // stages stay 128-byte aligned
  static constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
  static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 128 /*align slack*/ + 64 /*barriers*/;
};

template <int N, int ALIGN>
__device__ __forceinline__ void lds_row(const float *p, float (&dst)[N]) {
  // N is a multiple of ALIGN elements; p is ALIGN*4-byte aligned.
  if constexpr (ALIGN == 4) {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      float4 v = *reinterpret_cast<const float4 *>(p + i);
      dst[i] = v.x; dst[i + 1] = v.y; dst[i + 2] = v.z; dst[i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      float2 v = *reinterpret_cast<const float2 *>(p + i);
      dst[i] = v.x; dst[i + 1] = v.y;
    }
  }
}

// Micro-tile of RO x kCO outputs from the staged tile; XS = column shift of the window
// inside the 16-byte-aligned tile (compile time so every LDS stays a 64/128-bit access).
template <int U, int D, int QX, int QY, int TOW, int TOH, int PZ, int XS>
__device__ __forceinline__ void micro_tile(const float *wp, const float (&w)[kK][kK],
                                           float (&acc)[Cfg<U, D, QX, QY, TOW, TOH, PZ>::RO][kCO]) {
  using C = Cfg<U, D, QX, QY, TOW, TOH, PZ>;
  constexpr int kRO = C::RO;
  constexpr int LO = (XS / C::V) * C::V;                               // first aligned column read
  constexpr int NV = ((XS + C::WC + C::V - 1) / C::V) * C::V - LO;     // columns read (multiple of V)
  float win[C::WR][NV];
#pragma unroll
  for (int r = 0; r < C::WR; ++r) lds_row<NV, C::V>(wp + r * C::TIW + LO, win[r]);
#pragma unroll
  for (int r = 0; r < kRO; ++r)
#pragma unroll
    for (int c = 0; c < kCO; ++c) acc[r][c] = 0.f;
  // tap order: y outer, x inner, sequential FMA (as op/upfirdn2d_kernel.cu:193-198)
#pragma unroll
  for (int r = 0; r < kRO; ++r)
#pragma unroll
    for (int jy = 0; jy < kK; ++jy) {
      const int ty_u = QY + r * D + jy;
      if (ty_u % U != 0) continue;
#pragma unroll
      for (int c = 0; c < kCO; ++c)
#pragma unroll
        for (int jx = 0; jx < kK; ++jx) {
          const int tx_u = QX + c * D + jx;
          if (tx_u % U != 0) continue;
          acc[r][c] = fmaf(win[ty_u / U][tx_u / U + XS - LO], w[jy][jx], acc[r][c]);
        }
    }
}

template <int U, int D, int QX, int QY, int TOW, int TOH, int PZ>
__global__ void __launch_bounds__(kThreads)
upfirdn2d_tile_kernel(const UfdParams p, const __grid_constant__ CUtensorMap tmap) {
  using C = Cfg<U, D, QX, QY, TOW, TOH, PZ>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem_al =
      reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float *stage0 = reinterpret_cast<float *>(smem_al);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_al + 2 * C::STAGE_BYTES);

  const int tid = threadIdx.x;
  // block -> (tile_x, tile_y, z chunk)
  int bid = blockIdx.x;
  const int tile_x = bid % p.tiles_x; bid /= p.tiles_x;
  const int tile_y = bid % p.tiles_y; bid /= p.tiles_y;
  const int zchunk = bid;
  const int ox0 = tile_x * TOW, oy0 = tile_y * TOH;
  // first input sample of the tile (may be negative: padding)
  const int ix_first = (ox0 * D - p.pad_x0 - QX) / U;  // exact by construction of QX
  const int iy_base = (oy0 * D - p.pad_y0 - QY) / U;
  const int xs = ix_first & 3;                         // two's complement: also right for negatives
  const int ix_base = ix_first - xs;                   // multiple of 4 floats = 16 bytes

  // flipped, zero-extended filter in registers: w[jy][jx] = filt[kh-1-jy][kw-1-jx]
  float w[kK][kK];
#pragma unroll
  for (int jy = 0; jy < kK; ++jy)
#pragma unroll
    for (int jx = 0; jx < kK; ++jx)
      w[jy][jx] = (jy < p.kh && jx < p.kw) ? __ldg(p.filt + (p.kh - 1 - jy) * p.kw + (p.kw - 1 - jx)) : 0.f;

  const int tz = tid / (C::TX * C::TY);
  const int trem = tid - tz * (C::TX * C::TY);
  const int ty = trem / C::TX;
  const int tx = trem - ty * C::TX;

  const int zg_begin = zchunk * p.iters;
  const int zg_end = min(zg_begin + p.iters, p.zgroups);
  const int n_it = zg_end - zg_begin;
  if (n_it <= 0) return;

  const bool use_tma = p.use_tma != 0;
  if (use_tma) {
    if (tid == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(&bars[0], C::TILE_BYTES);
      tma_load_3d(stage0, &tmap, &bars[0], ix_base, iy_base, zg_begin * PZ);
    }
  }

  for (int it = 0; it < n_it; ++it) {
    const long long pz0 = (long long)(zg_begin + it) * PZ;
    float *tile = stage0 + (use_tma ? (it & 1) * C::STAGE_FLOATS : 0);
    if (use_tma) {
      if (tid == 0 && it + 1 < n_it) {
        const int s = (it + 1) & 1;
        mbar_arrive_expect_tx(&bars[s], C::TILE_BYTES);
        tma_load_3d(stage0 + s * C::STAGE_FLOATS, &tmap, &bars[s], ix_base, iy_base,
                    (int)(pz0 + PZ));
      }
      mbar_wait(&bars[it & 1], (it >> 1) & 1);
    } else {
      // Rows whose pitch is not a multiple of 16 bytes (odd widths: the (2H+1)^2 output of a stride-2 transposed conv)
      // cannot be described to the TMA.  Staging walks each tile row in ALIGNED 16-byte vectors of the flat tensor (one
      // LDG.128 + four predicated STS per vector instead of four LDG.32 with per-element index arithmetic; a 4-byte
      // cp.async ring was tried and is slower); vectors that touch the tensor's first / last bytes fall back to scalars.
      constexpr int NV = (C::TIW + 6) / 4;                 // vectors overlapping a row at any misalignment
      const bool vec_ok = (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
      const long long total = p.major * (long long)p.in_h * p.in_w;
      for (int t = tid; t < PZ * C::TIH * NV; t += kThreads) {
        const int vi = t % NV;
        const int rz = t / NV;
        const int r = rz % C::TIH;
        const int z = rz / C::TIH;
        const int iy = iy_base + r;
        const long long pl = pz0 + z;
        const bool row_ok = iy >= 0 && iy < p.in_h && pl < p.major;
        const long long g0 = (pl * p.in_h + iy) * (long long)p.in_w + ix_base;      // flat index of column 0 (may be < 0)
        const int a = (int)(g0 & 3);                                               // two's complement: right for negatives
        const long long gv = g0 - a + 4LL * vi;                                     // aligned vector start
        const int c0 = 4 * vi - a;                                                  // its first column in the tile row
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (row_ok) {
          if (vec_ok && gv >= 0 && gv + 3 < total) {
            const float4 q = ld_stream_f4(reinterpret_cast<const float4 *>(p.x + gv));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (gv + e >= 0 && gv + e < total) v[e] = ld_stream_f1(p.x + gv + e);
          }
        }
        float *dst = tile + (z * C::TIH + r) * C::TIW;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + e, ix = ix_base + c;
          if (c >= 0 && c < C::TIW) dst[c] = (row_ok && ix >= 0 && ix < p.in_w) ? v[e] : 0.f;
        }
      }
      __syncthreads();
    }

    // ---- compute the 2x4 micro tile (uniform switch on the alignment shift)
    constexpr int kRO = C::RO;
    float acc[kRO][kCO];
    const float *wp = tile + (tz * C::TIH + ty * C::SY) * C::TIW + tx * C::SX;
    switch (xs) {
      case 0: micro_tile<U, D, QX, QY, TOW, TOH, PZ, 0>(wp, w, acc); break;
      case 1: micro_tile<U, D, QX, QY, TOW, TOH, PZ, 1>(wp, w, acc); break;
      case 2: micro_tile<U, D, QX, QY, TOW, TOH, PZ, 2>(wp, w, acc); break;
      default: micro_tile<U, D, QX, QY, TOW, TOH, PZ, 3>(wp, w, acc); break;
    }

    // ---- epilogue + store
    const long long pl = pz0 + tz;
    if (pl < p.major) {
      float b = 0.f;
      if (p.act != 0 && p.bias != nullptr) b = __ldg(p.bias + (int)(pl % p.channels));
      const int ox = ox0 + tx * kCO;
#pragma unroll
      for (int r = 0; r < kRO; ++r) {
        const int oy = oy0 + ty * kRO + r;
        if (oy >= p.out_h || ox >= p.out_w) continue;
        float o[kCO];
#pragma unroll
        for (int c = 0; c < kCO; ++c) {
          float v = acc[r][c];
          if (p.act != 0) {
            v += b;
            v = (v > 0.f ? v : v * p.alpha) * p.scale;
          }
          o[c] = v;
        }
        float *dst = p.y + (pl * p.out_h + oy) * (long long)p.out_w + ox;
        if ((p.out_w & 3) == 0) {
          st_stream_f4(reinterpret_cast<float4 *>(dst), make_float4(o[0], o[1], o[2], o[3]));
        } else {
#pragma unroll
          for (int c = 0; c < kCO; ++c)
            if (ox + c < p.out_w) st_stream_f1(dst + c, o[c]);
        }
      }
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
  }
}

// Generic fallback: any factors / filter size / pads. One thread per output.
__global__ void __launch_bounds__(kThreads)
upfirdn2d_generic_kernel(const UfdParams p, long long total) {
  for (long long idx = blockIdx.x * (long long)kThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kThreads) {
    const int ox = (int)(idx % p.out_w);
    const long long t = idx / p.out_w;
    const int oy = (int)(t % p.out_h);
    const long long pl = t / p.out_h;
    const float *xp = p.x + pl * p.in_h * (long long)p.in_w;
    const int base_y = oy * p.down_y - p.pad_y0;
    const int base_x = ox * p.down_x - p.pad_x0;
    float acc = 0.f;
    for (int jy = 0; jy < p.kh; ++jy) {
      const int uy = base_y + jy;
      if (uy < 0 || uy % p.up_y != 0) continue;
      const int iy = uy / p.up_y;
      if (iy >= p.in_h) continue;
      for (int jx = 0; jx < p.kw; ++jx) {
        const int ux = base_x + jx;
        if (ux < 0 || ux % p.up_x != 0) continue;
        const int ix = ux / p.up_x;
        if (ix >= p.in_w) continue;
        acc = fmaf(__ldg(xp + (long long)iy * p.in_w + ix),
                   __ldg(p.filt + (p.kh - 1 - jy) * p.kw + (p.kw - 1 - jx)), acc);
      }
    }
    if (p.act != 0) {
      if (p.bias != nullptr) acc += __ldg(p.bias + (int)(pl % p.channels));
      acc = (acc > 0.f ? acc : acc * p.alpha) * p.scale;
    }
    p.y[idx] = acc;
  }
}

// Channels-last bf16 variant: one thread per (output pixel, 8-channel group), 128-bit
// loads/stores along the contiguous channel axis, fp32 accumulation, optional epilogue
// (noise + bias + leaky ReLU, then up to two residual adds — StyledConv(upsample) tail and
// the decoder skip fusion of models/RestoreNet.py:1031-1035).
struct NhwcEpi {
  const float *noise;
  long long noise_bstride;
  float noise_weight;
  const float *noise_weight_dev;
  const float *bias;
  int act;
  float alpha, scale;
  const uint4 *residual;
  const uint4 *residual2;
};

__device__ __forceinline__ void add_bf16x8(float (&acc)[8], const uint4 &v) {
  const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    acc[2 * i] += f.x;
    acc[2 * i + 1] += f.y;
  }
}

__global__ void __launch_bounds__(kThreads)
upfirdn2d_nhwc_kernel(const uint4 *__restrict__ x, const float *__restrict__ filt, uint4 *__restrict__ y,
                      const UfdParams p, const NhwcEpi e, int cg, long long total) {
  const float nw = e.noise ? (e.noise_weight_dev ? __ldg(e.noise_weight_dev) : e.noise_weight) : 0.f;
  for (long long idx = blockIdx.x * (long long)kThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kThreads) {
    const int g = (int)(idx % cg);
    long long t = idx / cg;
    const int ox = (int)(t % p.out_w); t /= p.out_w;
    const int oy = (int)(t % p.out_h);
    const long long b = t / p.out_h;
    const int base_y = oy * p.down_y - p.pad_y0;
    const int base_x = ox * p.down_x - p.pad_x0;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int jy = 0; jy < p.kh; ++jy) {
      const int uy = base_y + jy;
      if (uy < 0 || uy % p.up_y != 0) continue;
      const int iy = uy / p.up_y;
      if (iy >= p.in_h) continue;
      for (int jx = 0; jx < p.kw; ++jx) {
        const int ux = base_x + jx;
        if (ux < 0 || ux % p.up_x != 0) continue;
        const int ix = ux / p.up_x;
        if (ix >= p.in_w) continue;
        const float w = __ldg(filt + (p.kh - 1 - jy) * p.kw + (p.kw - 1 - jx));
        const uint4 v = __ldg(x + ((b * p.in_h + iy) * (long long)p.in_w + ix) * cg + g);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          acc[2 * i] = fmaf(f.x, w, acc[2 * i]);
          acc[2 * i + 1] = fmaf(f.y, w, acc[2 * i + 1]);
        }
      }
    }
    if (e.noise != nullptr || e.bias != nullptr || e.act != 0) {
      const float nz = e.noise ? nw * __ldg(e.noise + b * e.noise_bstride + (long long)oy * p.out_w + ox) : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = acc[i] + nz;
        if (e.bias) v += __ldg(e.bias + g * 8 + i);
        if (e.act == 3) v = (v > 0.f ? v : v * e.alpha) * e.scale;
        acc[i] = v;
      }
    }
    if (e.residual) add_bf16x8(acc, __ldg(e.residual + idx));
    if (e.residual2) add_bf16x8(acc, __ldg(e.residual2 + idx));
    uint4 o;
    __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
    y[idx] = o;
  }
}

// Channels-last fast path for the model's blurs (up = down = 1, filter <= 4x4): one thread owns
// 8 channels x R consecutive output rows of one column, walks the R+3 input rows once (4 x 128-bit
// loads per row, neighbouring lanes = neighbouring channel groups / pixels -> full 128 B lines out of
// L1) and scatters each row into the <= 4 output rows it feeds: (R+3)*4 loads for R outputs instead
// of 16 per output, no div/mod in the tap loops.
template <int R>
__global__ void __launch_bounds__(kThreads)
blur_nhwc_kernel(const uint4 *__restrict__ x, const float *__restrict__ filt, uint4 *__restrict__ y,
                 const UfdParams p, const NhwcEpi e, int cg, int row_blocks, long long total) {
  float w[kK][kK];
#pragma unroll
  for (int jy = 0; jy < kK; ++jy)
#pragma unroll
    for (int jx = 0; jx < kK; ++jx)
      w[jy][jx] = (jy < p.kh && jx < p.kw) ? __ldg(filt + (p.kh - 1 - jy) * p.kw + (p.kw - 1 - jx)) : 0.f;
  const float nw = e.noise ? (e.noise_weight_dev ? __ldg(e.noise_weight_dev) : e.noise_weight) : 0.f;
  for (long long idx = blockIdx.x * (long long)kThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kThreads) {
    const int g = (int)(idx % cg);
    long long t = idx / cg;
    const int ox = (int)(t % p.out_w); t /= p.out_w;
    const int rb = (int)(t % row_blocks);
    const long long b = t / row_blocks;
    const int oy0 = rb * R;
    const int iy0 = oy0 - p.pad_y0, ix0 = ox - p.pad_x0;
    float acc[R][8];
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
    const uint4 *xb = x + b * p.in_h * (long long)p.in_w * cg + g;
#pragma unroll
    for (int r = 0; r < R + kK - 1; ++r) {
      const int iy = iy0 + r;
      if (iy < 0 || iy >= p.in_h) continue;
      float in[kK][8];
#pragma unroll
      for (int j = 0; j < kK; ++j) {
        const int ix = ix0 + j;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ix >= 0 && ix < p.in_w) v = __ldg(xb + ((long long)iy * p.in_w + ix) * cg);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          in[j][2 * i] = f.x;
          in[j][2 * i + 1] = f.y;
        }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int jy = r - q;               // output row q reads input row r with filter row jy
        if (jy < 0 || jy >= kK) continue;
#pragma unroll
        for (int j = 0; j < kK; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[q][i] = fmaf(in[j][i], w[jy][j], acc[q][i]);
      }
    }
    float bias[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bias[i] = e.bias ? __ldg(e.bias + g * 8 + i) : 0.f;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int oy = oy0 + q;
      if (oy >= p.out_h) break;
      const long long opix = (b * p.out_h + oy) * (long long)p.out_w + ox;
      if (e.noise != nullptr || e.bias != nullptr || e.act != 0) {
        const float nz = e.noise ? nw * __ldg(e.noise + b * e.noise_bstride + (long long)oy * p.out_w + ox) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = acc[q][i] + nz + bias[i];
          if (e.act == 3) v = (v > 0.f ? v : v * e.alpha) * e.scale;
          acc[q][i] = v;
        }
      }
      if (e.residual) add_bf16x8(acc[q], __ldg(e.residual + opix * cg + g));
      if (e.residual2) add_bf16x8(acc[q], __ldg(e.residual2 + opix * cg + g));
      uint4 o;
      __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(acc[q][2 * i], acc[q][2 * i + 1]);
      y[opix * cg + g] = o;
    }
  }
}

// Separable form of the same blur (the model's filters are outer products of [1,3,3,1]; the host factorises the
// filter once and passes the 1-D taps by value): one thread owns 8 channels x 2 adjacent output columns x R output
// rows.  Each of the R+3 input rows gets ONE horizontal pass for both columns (5 x 128-bit loads instead of 8, 64 FMAs)
// and is then scattered into the <= 4 output rows it feeds with one FMA per value: 11 FMAs and 4.4 unpack operations
// per output element instead of 16 and 7 — the 2-D kernel above is issue-bound (ncu: 73 % issue slots busy).
struct SepTaps {
  float fx[kK], fy[kK];
};

template <int R>
__global__ void __launch_bounds__(kThreads)
blur_sep_nhwc_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, const UfdParams p, const NhwcEpi e,
                     const SepTaps t, int cg, int col_pairs, int row_blocks, long long total) {
  const float nw = e.noise ? (e.noise_weight_dev ? __ldg(e.noise_weight_dev) : e.noise_weight) : 0.f;
  for (long long idx = blockIdx.x * (long long)kThreads + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * kThreads) {
    const int g = (int)(idx % cg);
    long long q_ = idx / cg;
    const int cp = (int)(q_ % col_pairs); q_ /= col_pairs;
    const int rb = (int)(q_ % row_blocks);
    const long long b = q_ / row_blocks;
    const int ox0 = cp * 2, oy0 = rb * R;
    const int iy0 = oy0 - p.pad_y0, ix0 = ox0 - p.pad_x0;
    float acc[R][2][8];
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[q][c2][i] = 0.f;
    const uint4 *xb = x + b * p.in_h * (long long)p.in_w * cg + g;
#pragma unroll
    for (int r = 0; r < R + kK - 1; ++r) {
      const int iy = iy0 + r;
      if (iy < 0 || iy >= p.in_h) continue;
      float h0[8], h1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h0[i] = h1[i] = 0.f;
      const uint4 *xr = xb + (long long)iy * p.in_w * cg;
#pragma unroll
      for (int j = 0; j < kK + 1; ++j) {
        const int ix = ix0 + j;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ix >= 0 && ix < p.in_w) v = __ldg(xr + (long long)ix * cg);
        const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lo = __uint_as_float(wd[i] << 16), hi = __uint_as_float(wd[i] & 0xFFFF0000u);
          if (j < kK) {
            h0[2 * i] = fmaf(lo, t.fx[j], h0[2 * i]);
            h0[2 * i + 1] = fmaf(hi, t.fx[j], h0[2 * i + 1]);
          }
          if (j > 0) {
            h1[2 * i] = fmaf(lo, t.fx[j - 1], h1[2 * i]);
            h1[2 * i + 1] = fmaf(hi, t.fx[j - 1], h1[2 * i + 1]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int jy = r - q;               // output row q reads input row r with vertical tap jy
        if (jy < 0 || jy >= kK) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[q][0][i] = fmaf(h0[i], t.fy[jy], acc[q][0][i]);
          acc[q][1][i] = fmaf(h1[i], t.fy[jy], acc[q][1][i]);
        }
      }
    }
    float bias[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bias[i] = e.bias ? __ldg(e.bias + g * 8 + i) : 0.f;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int oy = oy0 + q;
      if (oy >= p.out_h) break;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int ox = ox0 + c2;
        if (ox >= p.out_w) continue;
        const long long opix = (b * p.out_h + oy) * (long long)p.out_w + ox;
        if (e.noise != nullptr || e.bias != nullptr || e.act != 0) {
          const float nz = e.noise ? nw * __ldg(e.noise + b * e.noise_bstride + (long long)oy * p.out_w + ox) : 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v = acc[q][c2][i] + nz + bias[i];
            if (e.act == 3) v = (v > 0.f ? v : v * e.alpha) * e.scale;
            acc[q][c2][i] = v;
          }
        }
        if (e.residual) add_bf16x8(acc[q][c2], __ldg(e.residual + opix * cg + g));
        if (e.residual2) add_bf16x8(acc[q][c2], __ldg(e.residual2 + opix * cg + g));
        uint4 o;
        __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(acc[q][c2][2 * i], acc[q][c2][2 * i + 1]);
        y[opix * cg + g] = o;
      }
    }
  }
}

// ---- TMA strip form of the separable NHWC blur -------------------------------------------------------------------------
// blur_sep_nhwc_kernel is latency- and issue-bound (119 registers -> 16 warps/SM, per-load address arithmetic and bounds
// tests, ~28 instructions per output element: 2.3 TB/s).  Here a CTA owns a strip of 32 output columns x 64 channels and
// walks DOWN it: input rows arrive as TMA boxes [64 ch, 35 cols, 4 rows] in a 4-stage mbarrier ring (out-of-bounds zero
// fill == the blur padding, no address math or predicates in the loop, every input row read once — no vertical halo), each
// thread (8 channels x 2 adjacent columns) does one horizontal pass per input row from 5 conflict-free 128-bit LDS and
// scatters it into a rolling window of 4 output-row accumulators; with 4 rows per stage == 4 filter taps the window rotates
// with compile-time indices.  All arithmetic is packed fp32 (fma.rn.f32x2 on (even, odd) channel pairs): ~11 instructions
// per output element.
constexpr int kStripW = 32;
constexpr int kStripCols = kStripW + kK - 1;
constexpr int kStageRows = 4;
constexpr int kStripStages = 4;
constexpr int kStripThreads = 128;
constexpr int kStageBytes = kStageRows * kStripCols * 128;
static_assert(kStageRows == kK, "rolling window rotation assumes rows per stage == filter taps");

struct StripParams {
  int in_h, in_w, out_h, out_w, pad_x0, pad_y0;
  int n_strips, n_chunks, n_seg, seg_rows, cg;
};

__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2_(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fmul2_(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// kEpiAhead: rows of look-ahead of the epilogue operand loads (< kStageRows)
template <bool EPI, int kEpiAhead, int kMinBlocks>
__global__ void __launch_bounds__(kStripThreads, kMinBlocks)
blur_strip_kernel(const __grid_constant__ CUtensorMap tmx, uint4 *__restrict__ y, const StripParams p, const NhwcEpi e,
                  const SepTaps t) {
  extern __shared__ __align__(128) unsigned char strip_smem[];
  __shared__ uint64_t full[kStripStages], empty[kStripStages];
  const int tid = threadIdx.x, lane = tid & 31;
  int bid = blockIdx.x;
  const int sx = bid % p.n_strips; bid /= p.n_strips;
  const int cz = bid % p.n_chunks; bid /= p.n_chunks;
  const int sy = bid % p.n_seg;
  const int b = bid / p.n_seg;
  const int q_lo = sy * p.seg_rows;
  const int rows_out = min(p.seg_rows, p.out_h - q_lo);
  const int n_stage = (rows_out + kK - 1 + kStageRows - 1) / kStageRows;
  const int ox0 = sx * kStripW, ix0 = ox0 - p.pad_x0, iy0 = q_lo - p.pad_y0;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStripStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kStripThreads / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    for (int k = 0; k < kStripStages && k < n_stage; ++k) {
      mbar_arrive_expect_tx(&full[k], kStageBytes);
      tma_load_4d(strip_smem + k * kStageBytes, &tmx, &full[k], cz * 64, ix0, iy0 + k * kStageRows, b);
    }
  }
  const int cg8 = tid & 7, cp = tid >> 3;
  const int ox = ox0 + 2 * cp;
  const bool ok0 = ox < p.out_w, ok1 = ox + 1 < p.out_w;
  unsigned long long fxp[kK], fyp[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) {
    fxp[j] = pk2(t.fx[j], t.fx[j]);
    fyp[j] = pk2(t.fy[j], t.fy[j]);
  }
  unsigned long long acc[kK][2][4];
#pragma unroll
  for (int q = 0; q < kK; ++q)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[q][c][i] = 0ull;
  float bias[8];
  float nw = 0.f;
  if (EPI) {
#pragma unroll
    for (int i = 0; i < 8; ++i) bias[i] = e.bias ? __ldg(e.bias + (cz * 8 + cg8) * 8 + i) : 0.f;
    nw = e.noise ? (e.noise_weight_dev ? __ldg(e.noise_weight_dev) : e.noise_weight) : 0.f;
  }
  const uint32_t tbase = smem_u32(strip_smem) + (uint32_t)(2 * cp) * 128u + (uint32_t)cg8 * 16u;
  const long long ybase = (long long)b * p.out_h * p.out_w * p.cg + cz * 8 + cg8;

  // ring of epilogue operands, slot = input row % kStageRows (compile-time inside the unrolled stage)
  uint4 r1[kStageRows][2], r2[kStageRows][2];
  float nz[kStageRows][2];
  auto epi_issue = [&](int i, int slot) {
    if (!(i >= kK - 1 && i - (kK - 1) < rows_out)) return;
    const long long pix = (long long)(q_lo + i - (kK - 1)) * p.out_w + ox;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const bool ok = c ? ok1 : ok0;
      r1[slot][c] = (e.residual && ok) ? __ldg(e.residual + ybase + (pix + c) * p.cg) : make_uint4(0u, 0u, 0u, 0u);
      r2[slot][c] = (e.residual2 && ok) ? __ldg(e.residual2 + ybase + (pix + c) * p.cg) : make_uint4(0u, 0u, 0u, 0u);
      nz[slot][c] = (e.noise && ok) ? __ldg(e.noise + b * e.noise_bstride + pix + c) : 0.f;   // raw: no use at issue time
    }
  };
  if (EPI) {
#pragma unroll
    for (int i = 0; i < kEpiAhead; ++i) epi_issue(i, i);
  }

  for (int k = 0; k < n_stage; ++k) {
    const int s = k % kStripStages;
    const uint32_t phase = (uint32_t)(k / kStripStages) & 1u;
    mbar_wait(&full[s], phase);
    const uint32_t st = tbase + (uint32_t)s * kStageBytes;
#pragma unroll
    for (int rr = 0; rr < kStageRows; ++rr) {
      const int i = k * kStageRows + rr;              // input row of this segment
      const int oy = q_lo + i - (kK - 1);             // the output row this input row completes
      const bool emit = i >= kK - 1 && i - (kK - 1) < rows_out;
      // epilogue operands: row i + kEpiAhead's residual / noise loads are issued here, kEpiAhead rows of arithmetic before
      // their use (issued only one row ahead the kernel sat on these loads: 3.2 TB/s vs 5.5 TB/s without an epilogue)
      if (EPI) epi_issue(i + kEpiAhead, (rr + kEpiAhead) % kStageRows);
      unsigned long long h0[4], h1[4];
#pragma unroll
      for (int j = 0; j < kK + 1; ++j) {
        const uint4 v = lds128(st + (uint32_t)(rr * kStripCols + j) * 128u);
        const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned long long xv = pk2(__uint_as_float(wd[q] << 16), __uint_as_float(wd[q] & 0xFFFF0000u));
          if (j == 0) h0[q] = fmul2_(xv, fxp[0]);
          else if (j < kK) h0[q] = ffma2_(xv, fxp[j], h0[q]);
          if (j == 1) h1[q] = fmul2_(xv, fxp[0]);
          else if (j > 1) h1[q] = ffma2_(xv, fxp[j - 1], h1[q]);
        }
      }
      // rolling window: slot rr starts output row i (tap 0); slots rr-1, rr-2, rr-3 take taps 1, 2, 3
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[rr][0][q] = fmul2_(h0[q], fyp[0]);
        acc[rr][1][q] = fmul2_(h1[q], fyp[0]);
#pragma unroll
        for (int j = 1; j < kK; ++j) {
          acc[(rr + kK - j) % kK][0][q] = ffma2_(h0[q], fyp[j], acc[(rr + kK - j) % kK][0][q]);
          acc[(rr + kK - j) % kK][1][q] = ffma2_(h1[q], fyp[j], acc[(rr + kK - j) % kK][1][q]);
        }
      }
      if (emit) {
        constexpr int dummy = 0;
        (void)dummy;
        const int slot = (rr + 1) % kK;               // the row that has now seen all kK taps
        const long long pix = (long long)oy * p.out_w + ox;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (!(c ? ok1 : ok0)) continue;
          float v[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) unpk2(acc[slot][c][q], v[2 * q], v[2 * q + 1]);
          if (EPI) {
            if (e.noise != nullptr || e.bias != nullptr || e.act != 0) {
              const float nzw = nw * nz[rr][c];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float u = v[q] + nzw + bias[q];
                if (e.act == 3) u = (u > 0.f ? u : u * e.alpha) * e.scale;
                v[q] = u;
              }
            }
            if (e.residual) add_bf16x8(v, r1[rr][c]);
            if (e.residual2) add_bf16x8(v, r2[rr][c]);
          }
          uint4 o;
          __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
          for (int q = 0; q < 4; ++q) oh[q] = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
          y[ybase + (pix + c) * p.cg] = o;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (tid == 0 && k + kStripStages < n_stage) {
      mbar_wait(&empty[s], phase);
      mbar_arrive_expect_tx(&full[s], kStageBytes);
      tma_load_4d(strip_smem + s * kStageBytes, &tmx, &full[s], cz * 64, ix0, iy0 + (k + kStripStages) * kStageRows, b);
    }
  }
}

// ---- epilogue form with its operands on the bulk-copy path ------------------------------------------------------------
// ncu on blur_strip_kernel<EPI> with both skip residuals (profiles/r02_ncu_prof_blur_epi*.summary.txt): 54 % of the stall
// samples sit on the FIRST use of the epilogue operands of a row although their loads were issued three rows earlier — a
// row of a 4-warp CTA takes ~300 issue slots, three rows cover ~0.7 us, a loaded-DRAM round trip is longer, and registers
// for a deeper ring do not exist (204 at three rows).  Here the residual rows of a stage arrive like the input rows do: one
// TMA box [64 ch, 32 cols, 4 rows] per residual into the SAME stage buffer (the four output rows a stage completes are
// rows q_lo + 4k - 3 ... q_lo + 4k: out-of-range rows / columns zero-fill and are never emitted), so they are in flight a
// whole stage ahead at no register cost, read back with one conflict-free LDS.128 per pixel; the noise values of the NEXT
// stage (8 floats) are fetched while the current one is processed.  NRES = 0 / 1 / 2 residuals: 4 / 3 / 2 stages of
// 17.5 / 33.5 / 49.5 KB, two CTAs per SM.
template <int NRES>
struct StripEpiCfg {
  static constexpr int kResBytes = kStageRows * kStripW * 128;                     // one residual's rows of a stage
  static constexpr int kBytes = kStageBytes + NRES * kResBytes;
  static constexpr int kStages = NRES == 0 ? 4 : (NRES == 1 ? 3 : 2);
};

template <int NRES>
__global__ void __launch_bounds__(kStripThreads, NRES == 0 ? 3 : 2)
blur_strip_epi_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmr1,
                      const __grid_constant__ CUtensorMap tmr2, uint4 *__restrict__ y, const StripParams p, const NhwcEpi e,
                      const SepTaps t) {
  using Cfg = StripEpiCfg<NRES>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(128) unsigned char strip_smem[];
  __shared__ uint64_t full[kStages], empty[kStages];
  const int tid = threadIdx.x, lane = tid & 31;
  int bid = blockIdx.x;
  const int sx = bid % p.n_strips; bid /= p.n_strips;
  const int cz = bid % p.n_chunks; bid /= p.n_chunks;
  const int sy = bid % p.n_seg;
  const int b = bid / p.n_seg;
  const int q_lo = sy * p.seg_rows;
  const int rows_out = min(p.seg_rows, p.out_h - q_lo);
  const int n_stage = (rows_out + kK - 1 + kStageRows - 1) / kStageRows;
  const int ox0 = sx * kStripW, ix0 = ox0 - p.pad_x0, iy0 = q_lo - p.pad_y0;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kStripThreads / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  auto issue_stage = [&](int k) {                // one elected thread: input rows + the residual rows of the outputs they complete
    unsigned char *dst = strip_smem + (k % kStages) * Cfg::kBytes;
    uint64_t *bar = &full[k % kStages];
    mbar_arrive_expect_tx(bar, Cfg::kBytes);
    tma_load_4d(dst, &tmx, bar, cz * 64, ix0, iy0 + k * kStageRows, b);
    if (NRES >= 1) tma_load_4d(dst + kStageBytes, &tmr1, bar, cz * 64, ox0, q_lo + k * kStageRows - (kK - 1), b);
    if (NRES >= 2) tma_load_4d(dst + kStageBytes + Cfg::kResBytes, &tmr2, bar, cz * 64, ox0, q_lo + k * kStageRows - (kK - 1), b);
  };
  if (tid == 0) {
    tma_prefetch_desc(&tmx);
    if (NRES >= 1) tma_prefetch_desc(&tmr1);
    if (NRES >= 2) tma_prefetch_desc(&tmr2);
    for (int k = 0; k < kStages && k < n_stage; ++k) issue_stage(k);
  }
  const int cg8 = tid & 7, cp = tid >> 3;
  const int ox = ox0 + 2 * cp;
  const bool ok0 = ox < p.out_w, ok1 = ox + 1 < p.out_w;
  unsigned long long fxp[kK], fyp[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) {
    fxp[j] = pk2(t.fx[j], t.fx[j]);
    fyp[j] = pk2(t.fy[j], t.fy[j]);
  }
  unsigned long long acc[kK][2][4];
#pragma unroll
  for (int q = 0; q < kK; ++q)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[q][c][i] = 0ull;
  float bias[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bias[i] = e.bias ? __ldg(e.bias + (cz * 8 + cg8) * 8 + i) : 0.f;
  const float nw = e.noise ? (e.noise_weight_dev ? __ldg(e.noise_weight_dev) : e.noise_weight) : 0.f;
  const bool pointwise = e.noise != nullptr || e.bias != nullptr || e.act != 0;
  const uint32_t tbase = smem_u32(strip_smem) + (uint32_t)(2 * cp) * 128u + (uint32_t)cg8 * 16u;
  const long long ybase = (long long)b * p.out_h * p.out_w * p.cg + cz * 8 + cg8;
  // noise of the rows a stage emits, fetched one stage ahead (raw values: scaled at their use)
  float nz_cur[kStageRows][2], nz_nxt[kStageRows][2];
  auto fetch_noise = [&](int k, float (&nz)[kStageRows][2]) {
#pragma unroll
    for (int rr = 0; rr < kStageRows; ++rr) {
      const int i = k * kStageRows + rr;
      const bool emit = e.noise != nullptr && i >= kK - 1 && i - (kK - 1) < rows_out;
      const long long pix = (long long)(q_lo + i - (kK - 1)) * p.out_w + ox;
      nz[rr][0] = (emit && ok0) ? __ldg(e.noise + b * e.noise_bstride + pix) : 0.f;
      nz[rr][1] = (emit && ok1) ? __ldg(e.noise + b * e.noise_bstride + pix + 1) : 0.f;
    }
  };
  fetch_noise(0, nz_cur);

  for (int k = 0; k < n_stage; ++k) {
    const int s = k % kStages;
    const uint32_t phase = (uint32_t)(k / kStages) & 1u;
    if (k + 1 < n_stage) fetch_noise(k + 1, nz_nxt);
    mbar_wait(&full[s], phase);
    const uint32_t st = tbase + (uint32_t)s * Cfg::kBytes;
#pragma unroll
    for (int rr = 0; rr < kStageRows; ++rr) {
      const int i = k * kStageRows + rr;              // input row of this segment
      const int oy = q_lo + i - (kK - 1);             // the output row this input row completes
      const bool emit = i >= kK - 1 && i - (kK - 1) < rows_out;
      unsigned long long h0[4], h1[4];
#pragma unroll
      for (int j = 0; j < kK + 1; ++j) {
        const uint4 v = lds128(st + (uint32_t)(rr * kStripCols + j) * 128u);
        const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const unsigned long long xv = pk2(__uint_as_float(wd[q] << 16), __uint_as_float(wd[q] & 0xFFFF0000u));
          if (j == 0) h0[q] = fmul2_(xv, fxp[0]);
          else if (j < kK) h0[q] = ffma2_(xv, fxp[j], h0[q]);
          if (j == 1) h1[q] = fmul2_(xv, fxp[0]);
          else if (j > 1) h1[q] = ffma2_(xv, fxp[j - 1], h1[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[rr][0][q] = fmul2_(h0[q], fyp[0]);
        acc[rr][1][q] = fmul2_(h1[q], fyp[0]);
#pragma unroll
        for (int j = 1; j < kK; ++j) {
          acc[(rr + kK - j) % kK][0][q] = ffma2_(h0[q], fyp[j], acc[(rr + kK - j) % kK][0][q]);
          acc[(rr + kK - j) % kK][1][q] = ffma2_(h1[q], fyp[j], acc[(rr + kK - j) % kK][1][q]);
        }
      }
      if (emit) {
        const int slot = (rr + 1) % kK;               // the row that has now seen all kK taps
        const long long pix = (long long)oy * p.out_w + ox;
        // residual rows of this stage: [row rr][column][64 channels], this thread's pixel pair at column 2 cp
        const uint32_t rbase = st - (uint32_t)(2 * cp) * 128u + kStageBytes + (uint32_t)(rr * kStripW + 2 * cp) * 128u;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (!(c ? ok1 : ok0)) continue;
          float v[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) unpk2(acc[slot][c][q], v[2 * q], v[2 * q + 1]);
          if (pointwise) {
            const float nzw = nw * nz_cur[rr][c];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float u = v[q] + nzw + bias[q];
              if (e.act == 3) u = (u > 0.f ? u : u * e.alpha) * e.scale;
              v[q] = u;
            }
          }
          if (NRES >= 1) add_bf16x8(v, lds128(rbase + (uint32_t)c * 128u));
          if (NRES >= 2) add_bf16x8(v, lds128(rbase + Cfg::kResBytes + (uint32_t)c * 128u));
          uint4 o;
          __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
          for (int q = 0; q < 4; ++q) oh[q] = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
          y[ybase + (pix + c) * p.cg] = o;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (tid == 0 && k + kStages < n_stage) {
      mbar_wait(&empty[s], phase);
      issue_stage(k + kStages);
    }
#pragma unroll
    for (int rr = 0; rr < kStageRows; ++rr) { nz_cur[rr][0] = nz_nxt[rr][0]; nz_cur[rr][1] = nz_nxt[rr][1]; }
  }
}

template <int NRES>
int launch_blur_strip_epi(const void *x, void *y, StripParams p, const NhwcEpi &e, const void *r1, const void *r2,
                          const SepTaps &t, int64_t n, int64_t c, cudaStream_t stream) {
  using Cfg = StripEpiCfg<NRES>;
  auto kern = blur_strip_epi_kernel<NRES>;
  constexpr int smem = Cfg::kStages * Cfg::kBytes;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  p.n_strips = (p.out_w + kStripW - 1) / kStripW;
  p.n_chunks = (int)(c / 64);
  p.cg = (int)(c / 8);
  const long long base_blocks = (long long)p.n_strips * p.n_chunks * n;
  long long n_seg = ((long long)num_sms() * (NRES == 0 ? 6 : 4) + base_blocks - 1) / base_blocks;     // two waves of resident CTAs
  const long long max_seg = p.out_h >= 32 ? p.out_h / 32 : 1;
  if (n_seg > max_seg) n_seg = max_seg;
  if (n_seg < 1) n_seg = 1;
  p.seg_rows = (int)((p.out_h + n_seg - 1) / n_seg);
  p.n_seg = (p.out_h + p.seg_rows - 1) / p.seg_rows;
  const long long blocks = base_blocks * p.n_seg;
  VSP_REQUIRE(blocks < 2147483647LL, "blur_strip_epi: grid too large (%lld blocks)", blocks);
  CUtensorMap tmx, tmr[2];
  memset(&tmx, 0, sizeof(tmx));
  memset(tmr, 0, sizeof(tmr));
  {
    uint64_t dims[4] = {(uint64_t)c, (uint64_t)p.in_w, (uint64_t)p.in_h, (uint64_t)n};
    uint64_t strides[4] = {0, (uint64_t)c * 2, (uint64_t)c * p.in_w * 2, (uint64_t)c * p.in_w * p.in_h * 2};
    uint32_t box[4] = {64, (uint32_t)kStripCols, (uint32_t)kStageRows, 1};
    if (int rc = encode_tma(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
  }
  const void *res[2] = {r1, r2};
  for (int i = 0; i < NRES; ++i) {
    uint64_t dims[4] = {(uint64_t)c, (uint64_t)p.out_w, (uint64_t)p.out_h, (uint64_t)n};
    uint64_t strides[4] = {0, (uint64_t)c * 2, (uint64_t)c * p.out_w * 2, (uint64_t)c * p.out_w * p.out_h * 2};
    uint32_t box[4] = {64, (uint32_t)kStripW, (uint32_t)kStageRows, 1};
    if (int rc = encode_tma(&tmr[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, res[i], dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_NONE))
      return rc;
  }
  kern<<<(unsigned)blocks, kStripThreads, smem, stream>>>(tmx, tmr[0], tmr[1], static_cast<uint4 *>(y), p, e, t);
  return check_launch("blur_strip_epi_kernel");
}

template <bool EPI, int AHEAD = 0, int MINB = 3>
int launch_blur_strip(const void *x, void *y, StripParams p, const NhwcEpi &e, const SepTaps &t, int64_t n, int64_t c,
                      cudaStream_t stream) {
  auto kern = blur_strip_kernel<EPI, AHEAD, MINB>;
  constexpr int smem = kStripStages * kStageBytes;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  p.n_strips = (p.out_w + kStripW - 1) / kStripW;
  p.n_chunks = (int)(c / 64);
  p.cg = (int)(c / 8);
  // vertical segments only while the strips alone cannot fill the machine (each segment re-reads 3 halo rows)
  const long long base_blocks = (long long)p.n_strips * p.n_chunks * n;
  long long n_seg = ((long long)num_sms() * 6 + base_blocks - 1) / base_blocks;
  const long long max_seg = p.out_h >= 32 ? p.out_h / 32 : 1;
  if (n_seg > max_seg) n_seg = max_seg;
  if (n_seg < 1) n_seg = 1;
  p.seg_rows = (int)((p.out_h + n_seg - 1) / n_seg);
  p.n_seg = (p.out_h + p.seg_rows - 1) / p.seg_rows;
  const long long blocks = base_blocks * p.n_seg;
  VSP_REQUIRE(blocks < 2147483647LL, "blur_strip: grid too large (%lld blocks)", blocks);
  CUtensorMap tmx;
  memset(&tmx, 0, sizeof(tmx));
  uint64_t dims[4] = {(uint64_t)c, (uint64_t)p.in_w, (uint64_t)p.in_h, (uint64_t)n};
  uint64_t strides[4] = {0, (uint64_t)c * 2, (uint64_t)c * p.in_w * 2, (uint64_t)c * p.in_w * p.in_h * 2};
  uint32_t box[4] = {64, (uint32_t)kStripCols, (uint32_t)kStageRows, 1};
  if (int rc = encode_tma(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, nullptr,
                          CU_TENSOR_MAP_SWIZZLE_NONE))
    return rc;
  kern<<<(unsigned)blocks, kStripThreads, smem, stream>>>(tmx, static_cast<uint4 *>(y), p, e, t);
  return check_launch("blur_strip_kernel");
}

template <int U, int D, int QX, int QY, int TOW, int TOH, int PZ>
int launch_tile(UfdParams p, cudaStream_t stream) {
  using C = Cfg<U, D, QX, QY, TOW, TOH, PZ>;
  auto kern = upfirdn2d_tile_kernel<U, D, QX, QY, TOW, TOH, PZ>;
  static bool attr_done[64] = {false};  // per device; benign race: the call is idempotent
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  p.tiles_x = (p.out_w + TOW - 1) / TOW;
  p.tiles_y = (p.out_h + TOH - 1) / TOH;
  p.zgroups = (int)((p.major + PZ - 1) / PZ);
  const long long tiles_xy = (long long)p.tiles_x * p.tiles_y;
  // Grid: measured on B200 — the up-sampling and blur tiles (small input tile, many blocks per SM) run best oversubscribed
  // (~16 blocks per SM, at most 8 plane groups per block: config 1 34.8 us vs 37.6 us as one resident wave); the decimating
  // tile (36 KB per stage, 3 blocks per SM) runs best as exactly one resident wave (39.1 vs 41.0 us).
  long long want_blocks = (long long)num_sms() * 16, cap = 8;
  if (D == 2) {
    static int occ[64] = {0};
    if (dev >= 0 && dev < 64 && occ[dev] == 0) {
      int nb = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, C::SMEM_BYTES) != cudaSuccess || nb < 1) nb = 1;
      occ[dev] = nb;
    }
    want_blocks = (long long)num_sms() * ((dev >= 0 && dev < 64) ? occ[dev] : 3);
    cap = 64;
  }
  long long iters = (tiles_xy * p.zgroups + want_blocks - 1) / want_blocks;
  if (iters < 1) iters = 1;
  if (iters > cap) iters = cap;
  if (iters > p.zgroups) iters = p.zgroups;
  p.iters = (int)iters;
  const long long zchunks = (p.zgroups + iters - 1) / iters;
  const long long blocks = tiles_xy * zchunks;
  VSP_REQUIRE(blocks < 2147483647LL, "upfirdn2d: grid too large (%lld blocks)", blocks);

  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (p.use_tma) {
    uint64_t dims[3] = {(uint64_t)p.in_w, (uint64_t)p.in_h, (uint64_t)p.major};
    uint64_t strides[3] = {0, (uint64_t)p.in_w * 4, (uint64_t)p.in_w * p.in_h * 4};
    uint32_t box[3] = {(uint32_t)C::TIW, (uint32_t)C::TIH, (uint32_t)PZ};
    if (encode_tma(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.x, dims, strides, box, nullptr,
                   CU_TENSOR_MAP_SWIZZLE_NONE) != 0)
      p.use_tma = 0;  // fall back to LDG staging; error text is overwritten on success
  }
  kern<<<(unsigned)blocks, kThreads, C::SMEM_BYTES, stream>>>(p, tmap);
  return check_launch("upfirdn2d_tile_kernel");
}

template <int U, int D, int QX, int QY>
int dispatch_tile(const UfdParams &p, cudaStream_t stream) {
  const int ow = p.out_w, oh = p.out_h;
  if constexpr (D == 2) {
    if (ow > 32) return launch_tile<U, D, QX, QY, 64, 32, 1>(p, stream);
  } else {
    if constexpr (U == 2) {        // four output rows per thread: the 2 x 4 micro-tile was issue-bound (24 instructions per output)
      static const bool ro2 = getenv("VSP_UFD_RO2") != nullptr;
      if (ow > 64 && oh >= 32 && !ro2) return launch_tile<U, D, QX, QY, 128, 32, 1>(p, stream);
    }
    if (ow > 64) return launch_tile<U, D, QX, QY, 128, 16, 1>(p, stream);
    if (ow > 32) return launch_tile<U, D, QX, QY, 64, 32, 1>(p, stream);
  }
  if (ow > 16 || oh > 16) return launch_tile<U, D, QX, QY, 32, 32, 2>(p, stream);
  if (ow > 8 || oh > 8) return launch_tile<U, D, QX, QY, 16, 16, 8>(p, stream);
  return launch_tile<U, D, QX, QY, 8, 8, 32>(p, stream);
}


// ---------------------------------------------------------------------------------------------------------------
// Row-band kernel for up = down = 1 (the model's Blur: models/RestoreNet.py:84-101 after every transposed / before every
// strided convolution, almost always on ODD extents — [.,.,65,65], [.,.,513,513], [.,.,1025,1025] — whose row pitch no
// tensor map can describe).  A unit is a run of whole input rows of the flat tensor — R output rows (+ kh-1 halo rows)
// of one wide plane, or several whole small planes — i.e. ONE contiguous byte range, brought in by a single 1-D bulk copy
// (cp.async.bulk, 16-byte aligned by over-fetching <= 12 bytes at each end) into a 3-stage mbarrier ring by a producer
// warp: no address arithmetic, predicates or registers on the load side, every input row read once per unit.
// Compute (16 warps): item = (output column, group of R rows); the 32 lanes of a warp own 32 consecutive columns
// (conflict-free scalar LDS, 128-byte coalesced stores), warps take 32-item chunks round-robin with a rotation per unit so
// that ragged chunk counts (1026 columns = 32 chunks + 2 columns) average out over the ring instead of idling 15 warps.
// The filter is factorised in the kernel (rank-1 test on the 16 taps): separable filters — every filter the model
// builds, make_kernel's outer([1,3,3,1]) — take one 4-tap horizontal pass per input row and a vertical scatter into
// the rolling accumulators of the <= 4 output rows that use it (8 FMAs per output instead of 16; fp32 results differ from
// the 2-D tap order by rounding only); other filters run the 2-D form.
constexpr int kBandWarps = 16;
constexpr int kBandThreads = (kBandWarps + 1) * 32;   // + producer warp
constexpr int kBandStages = 3;
constexpr int kBandStageBytes = 73728;                // 3 x 72 KiB + barriers < 227 KiB
constexpr int kBandR = 8;                             // output rows per item

struct BandParams {
  UfdParams u;
  int P;                    // whole planes per unit (> 1 only when a plane is small)
  int G;                    // row groups per unit
  int upp;                  // units per plane (P == 1)
  long long units;
  long long total_floats;   // elements of x (bulk copies never read past the last whole 16 bytes)
  unsigned magic_ow, magic_pp;   // floor(2^32 / d) + 1 for d = out_w and d = G * out_w (items per plane when P > 1)
  int dn;                   // decimation factor (1: blur, 2: the Downsample / Upsample-backward mode), both axes
};

__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct BandUnit {
  long long plane0;         // first plane
  int np;                   // planes in the unit
  int oy0, rows;            // output rows [oy0, oy0 + rows) of each plane
  int iy_lo, iy_hi;         // input rows [iy_lo, iy_hi) of each plane held in the stage
  long long g_lo, g_hi;     // aligned float range of x in shared memory
};

__device__ __forceinline__ BandUnit band_unit(const BandParams &p, long long unit) {
  BandUnit b;
  const UfdParams &u = p.u;
  int iy_hi;
  if (p.P > 1) {
    b.plane0 = unit * p.P;
    b.np = (int)min((long long)p.P, u.major - b.plane0);
    b.oy0 = 0; b.rows = u.out_h;
    b.iy_lo = 0; iy_hi = u.in_h;
    b.iy_hi = iy_hi;
  } else {
    b.plane0 = unit / p.upp;
    b.np = 1;
    b.oy0 = (int)(unit % p.upp) * p.G * kBandR;
    b.rows = min(p.G * kBandR, u.out_h - b.oy0);
    b.iy_lo = max(0, b.oy0 * p.dn - u.pad_y0);
    // (kK rows per output even when the filter has fewer: the compute loop walks the zero-extended 4-tap window, and a
    // zero tap times a row that was never loaded — stale shared memory, possibly NaN bits — would not be zero)
    iy_hi = min(u.in_h, (b.oy0 + b.rows - 1) * p.dn - u.pad_y0 + kK);
    if (iy_hi < b.iy_lo) iy_hi = b.iy_lo;
    b.iy_hi = iy_hi;
  }
  const long long plane_sz = (long long)u.in_h * u.in_w;
  b.g_lo = (b.plane0 * plane_sz + (long long)b.iy_lo * u.in_w) & ~3LL;
  b.g_hi = ((b.plane0 + b.np - 1) * plane_sz + (long long)iy_hi * u.in_w + 3) & ~3LL;
  return b;
}

// Compute side of the band kernel.  SEP: the filter is fy (x) fx.  Interior items (all taps inside the image, full row
// group) run a predicate-free body: one shared-memory pointer bumped by the row pitch, four LDS at immediate offsets, the
// horizontal 4-tap product and the vertical scatter — 13 instructions per input row and column; items touching the image
// border take the predicated form.
template <bool SEP, int DN>
__device__ __forceinline__ void band_compute(const BandParams &p, unsigned char *band_smem, uint64_t *full, uint64_t *empty,
                                             const float *filt, const float *sep) {
  const UfdParams &u = p.u;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  float fy[kK], fx[kK], w2[SEP ? 1 : kK * kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) { fy[j] = sep[j]; fx[j] = sep[kK + j]; }
  if constexpr (!SEP) {
#pragma unroll
    for (int j = 0; j < kK * kK; ++j) w2[j] = filt[j];
  }
  const long long total_al = p.total_floats & ~3LL;
  const long long plane_sz = (long long)u.in_h * u.in_w;
  const int in_w = u.in_w, in_h = u.in_h, out_w = u.out_w;
  long long k = 0;
  for (long long unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++k) {
    const int s = (int)(k % kBandStages);
    const BandUnit b = band_unit(p, unit);
    mbar_wait(&full[s], (uint32_t)((k / kBandStages) & 1));
    float *sm = reinterpret_cast<float *>(band_smem + (size_t)s * kBandStageBytes);
    if (b.g_hi > total_al) {        // the (< 16 byte) tail of the tensor that a bulk copy may not touch: plain loads
      const long long n_tail = p.total_floats - total_al;
      if (tid < n_tail && total_al + tid >= b.g_lo) sm[total_al + tid - b.g_lo] = __ldg(u.x + total_al + tid);
      asm volatile("bar.sync 1, %0;" ::"n"(kBandWarps * 32) : "memory");
    }
    const int groups = (b.rows + kBandR - 1) / kBandR;
    const int per_plane = groups * out_w;
    const int items = b.np * per_plane;
    const int chunks = (items + 31) >> 5;
    const int rot = (int)((k * 5) & (kBandWarps - 1));
    const int base0 = (int)(b.plane0 * plane_sz - b.g_lo);           // smem index of x[plane0, 0, 0] (may be negative)
    // rows of each plane that the stage holds (loads outside the image are clamped into this range, then discarded)
    const int ylo = b.iy_lo, yhi = b.iy_hi;
    for (int c = (warp - rot) & (kBandWarps - 1); c < chunks; c += kBandWarps) {
      const int it = c * 32 + lane;
      const bool live = it < items;
      // (plane, row group, column) of the item: multiply-high by host-computed reciprocals (exact for it * divisor < 2^32)
      int pl = 0, rem = live ? it : 0;
      if (p.P > 1) {
        pl = (int)__umulhi((unsigned)rem, p.magic_pp);
        rem -= pl * per_plane;
      }
      const int g = (int)__umulhi((unsigned)rem, p.magic_ow), ox = rem - g * out_w;
      const int r0 = b.oy0 + g * kBandR;                            // first output row of this item
      const int nr = live ? min(kBandR, b.oy0 + b.rows - r0) : 0;
      constexpr int kRowsIn = DN * (kBandR - 1) + kK;              // input rows one item walks
      const int cx = ox * DN - u.pad_x0;                             // input column of tap jx = 0
      const int iy0 = r0 * DN - u.pad_y0;                            // input row of tap jy = 0 of output row r0
      const float *colp = sm + (base0 + pl * (int)plane_sz + cx);    // &x[plane, 0, cx]
      float acc[kBandR];
#pragma unroll
      for (int q = 0; q < kBandR; ++q) acc[q] = 0.f;
      const bool interior = nr == kBandR && cx >= 0 && cx + kK - 1 < in_w && iy0 >= 0 && iy0 + kRowsIn - 1 < in_h;
      if (__all_sync(0xffffffffu, interior)) {
        const float *sr = colp + iy0 * in_w;
#pragma unroll
        for (int t = 0; t < kRowsIn; ++t) {
          const float v0 = sr[0], v1 = sr[1], v2 = sr[2], v3 = sr[3];
          sr += in_w;
          if constexpr (SEP) {
            const float h = fmaf(v3, fx[3], fmaf(v2, fx[2], fmaf(v1, fx[1], v0 * fx[0])));
#pragma unroll
            for (int jy = 0; jy < kK; ++jy) {
              const int q = (t - jy) / DN;
              if (t - jy >= 0 && (t - jy) % DN == 0 && q < kBandR) acc[q] = fmaf(h, fy[jy], acc[q]);
            }
          } else {
#pragma unroll
            for (int jy = 0; jy < kK; ++jy) {
              const int q = (t - jy) / DN;
              if (t - jy >= 0 && (t - jy) % DN == 0 && q < kBandR)
                acc[q] = fmaf(v3, w2[jy * kK + 3], fmaf(v2, w2[jy * kK + 2], fmaf(v1, w2[jy * kK + 1], fmaf(v0, w2[jy * kK], acc[q]))));
            }
          }
        }
      } else {
        // border form, branch-free: column taps outside the image are predicated off, rows outside it are read from
        // the nearest held row and their contribution replaced by zero
        const bool any = live && yhi > ylo;
        const bool in0 = any && cx >= 0 && cx < in_w, in1 = any && cx + 1 >= 0 && cx + 1 < in_w,
                   in2 = any && cx + 2 >= 0 && cx + 2 < in_w, in3 = any && cx + 3 >= 0 && cx + 3 < in_w;
#pragma unroll
        for (int t = 0; t < kRowsIn; ++t) {
          const int iy = iy0 + t;
          const float *sr = colp + min(max(iy, ylo), yhi - 1) * in_w;
          const float v0 = in0 ? sr[0] : 0.f, v1 = in1 ? sr[1] : 0.f, v2 = in2 ? sr[2] : 0.f, v3 = in3 ? sr[3] : 0.f;
          const bool rv = (unsigned)iy < (unsigned)in_h && t < DN * (nr - 1) + kK;
          if constexpr (SEP) {
            float h = fmaf(v3, fx[3], fmaf(v2, fx[2], fmaf(v1, fx[1], v0 * fx[0])));
            h = rv ? h : 0.f;
#pragma unroll
            for (int jy = 0; jy < kK; ++jy) {
              const int q = (t - jy) / DN;
              if (t - jy >= 0 && (t - jy) % DN == 0 && q < kBandR) acc[q] = fmaf(h, fy[jy], acc[q]);
            }
          } else {
            const float z0 = rv ? v0 : 0.f, z1 = rv ? v1 : 0.f, z2 = rv ? v2 : 0.f, z3 = rv ? v3 : 0.f;
#pragma unroll
            for (int jy = 0; jy < kK; ++jy) {
              const int q = (t - jy) / DN;
              if (t - jy >= 0 && (t - jy) % DN == 0 && q < kBandR)
                acc[q] = fmaf(z3, w2[jy * kK + 3], fmaf(z2, w2[jy * kK + 2], fmaf(z1, w2[jy * kK + 1], fmaf(z0, w2[jy * kK], acc[q]))));
            }
          }
        }
      }
      const long long plane = b.plane0 + pl;
      float *yp = u.y + (plane * u.out_h + r0) * (long long)out_w + ox;
      if (u.act != 0) {
        const float bias = (u.bias != nullptr && live) ? __ldg(u.bias + (int)(plane % u.channels)) : 0.f;
#pragma unroll
        for (int q = 0; q < kBandR; ++q) {
          const float a = acc[q] + bias;
          acc[q] = (a > 0.f ? a : a * u.alpha) * u.scale;
        }
      }
      if (__all_sync(0xffffffffu, nr == kBandR)) {
#pragma unroll
        for (int q = 0; q < kBandR; ++q) {
          st_stream_f1(yp, acc[q]);
          yp += out_w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < kBandR; ++q) {
          if (q < nr) st_stream_f1(yp, acc[q]);
          yp += out_w;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
}

__global__ void __launch_bounds__(kBandThreads, 1)
upfirdn2d_band_kernel(const BandParams p) {
  extern __shared__ __align__(128) unsigned char band_smem[];
  auto stage = [&](int i) { return reinterpret_cast<float *>(band_smem + (size_t)i * kBandStageBytes); };
  uint64_t *full = reinterpret_cast<uint64_t *>(band_smem + (size_t)kBandStages * kBandStageBytes);
  uint64_t *empty = full + kBandStages;
  __shared__ float filt[kK * kK];        // flipped, zero-extended: filt[jy][jx] multiplies x[oy - pad_y0 + jy][ox - pad_x0 + jx]
  __shared__ float sep[2 * kK + 1];      // fy[4], fx[4], flag
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const UfdParams &u = p.u;
  if (tid < kK * kK) {
    const int jy = tid / kK, jx = tid % kK;
    filt[tid] = (jy < u.kh && jx < u.kw) ? __ldg(u.filt + (u.kh - 1 - jy) * u.kw + (u.kw - 1 - jx)) : 0.f;
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < kBandStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kBandWarps);
    }
    fence_barrier_init();
    // rank-1 test: pivot on the largest tap, fx = its row, fy = its column / pivot
    int best = 0;
    for (int i = 1; i < kK * kK; ++i)
      if (fabsf(filt[i]) > fabsf(filt[best])) best = i;
    const int j0 = best / kK, i0 = best % kK;
    const float piv = filt[best];
    float ok = piv != 0.f ? 1.f : 0.f;
    for (int j = 0; j < kK; ++j) {
      sep[j] = piv != 0.f ? filt[j * kK + i0] / piv : 0.f;
      sep[kK + j] = filt[j0 * kK + j];
    }
    for (int j = 0; j < kK; ++j)
      for (int i = 0; i < kK; ++i)
        if (fabsf(sep[j] * sep[kK + i] - filt[j * kK + i]) > 2e-7f * fabsf(piv)) ok = 0.f;
    sep[2 * kK] = ok;
  }
  __syncthreads();
  const long long total_al = p.total_floats & ~3LL;

  if (warp == kBandWarps) {
    // ===================== producer: one bulk copy per unit =====================
    if (lane == 0) {
      long long k = 0;
      for (long long unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++k) {
        const int s = (int)(k % kBandStages);
        if (k >= kBandStages) mbar_wait(&empty[s], (uint32_t)(((k / kBandStages) - 1) & 1));
        const BandUnit b = band_unit(p, unit);
        const long long hi = b.g_hi < total_al ? b.g_hi : total_al;
        const uint32_t bytes = hi > b.g_lo ? (uint32_t)((hi - b.g_lo) * 4) : 0u;
        mbar_arrive_expect_tx(&full[s], bytes);
        if (bytes) bulk_load_1d(stage(s), u.x + b.g_lo, bytes, &full[s]);
      }
    }
    return;
  }

  // ===================== compute warps =====================
  if (p.dn == 2) {
    if (sep[2 * kK] != 0.f) band_compute<true, 2>(p, band_smem, full, empty, filt, sep);
    else band_compute<false, 2>(p, band_smem, full, empty, filt, sep);
  } else {
    if (sep[2 * kK] != 0.f) band_compute<true, 1>(p, band_smem, full, empty, filt, sep);
    else band_compute<false, 1>(p, band_smem, full, empty, filt, sep);
  }
}

int launch_band(const UfdParams &u, int dn, cudaStream_t stream) {
  BandParams p;
  p.u = u;
  p.dn = dn;
  const long long row_bytes = (long long)u.in_w * 4;
  const long long plane_bytes = row_bytes * u.in_h;
  const long long budget = kBandStageBytes - 64;           // alignment over-fetch at both ends
  if (2 * plane_bytes <= budget) {
    // small planes: several whole planes per unit, but keep >= 3 units per SM in flight when the tensor allows it
    // the P that wastes least of the last wave of units over the SMs (each SM holds one CTA); larger P on near-ties
    long long P = 1;
    double best = -1.0;
    for (long long c = budget / plane_bytes; c >= 1; --c) {
      const long long un = (u.major + c - 1) / c;
      const long long waves = (un + num_sms() - 1) / num_sms();
      const double eff = (double)un / (double)(waves * num_sms());
      if (eff > best + 0.03) { best = eff; P = c; }
    }
    p.P = (int)P;
    p.G = (u.out_h + kBandR - 1) / kBandR;
    p.upp = 1;
    p.units = (u.major + P - 1) / P;
  } else {
    long long in_rows = budget / row_bytes;                 // input rows a stage holds: dn * (G * R - 1) + kh of them are needed
    long long G = ((in_rows - kK) / dn + 1) / kBandR;
    if (G < 1) return -1;                                   // one row group does not fit: caller falls back
    const long long gmax = (u.out_h + kBandR - 1) / kBandR;
    if (G > gmax) G = gmax;
    while (G > 1 && u.major * ((gmax + G - 1) / G) < 3LL * num_sms()) --G;
    p.P = 1;
    p.G = (int)G;
    p.upp = (int)((gmax + G - 1) / G);
    p.units = u.major * p.upp;
  }
  p.total_floats = u.major * (long long)u.in_h * u.in_w;
  // items of a unit are < 2^15 and the divisors < 2^15, so the multiply-high quotients are exact
  if ((long long)p.G * u.out_w >= (1 << 15)) return -1;
  p.magic_ow = (unsigned)((1ULL << 32) / (unsigned)u.out_w) + 1u;
  p.magic_pp = (unsigned)((1ULL << 32) / (unsigned)(p.G * u.out_w)) + 1u;
  auto kern = upfirdn2d_band_kernel;
  constexpr int smem = kBandStages * kBandStageBytes + 2 * kBandStages * 8;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  long long grid = p.units < num_sms() ? p.units : num_sms();
  kern<<<(unsigned)grid, kBandThreads, smem, stream>>>(p);
  return check_launch("upfirdn2d_band_kernel");
}

inline int pmod(int a, int m) { return ((a % m) + m) % m; }

}  // namespace
}  // namespace vsp

extern "C" int vsp_upfirdn2d_f32(const float *x, const float *filt, float *y, int64_t major,
                                 int64_t in_h, int64_t in_w, int kh, int kw, int up_x, int up_y,
                                 int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                                 int pad_y1, const float *bias, int64_t channels, int act,
                                 float alpha, float scale, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1,
              "upfirdn2d: up/down factors must be >= 1");
  VSP_REQUIRE(kh >= 1 && kw >= 1, "upfirdn2d: empty filter");
  VSP_REQUIRE(major >= 0 && in_h >= 0 && in_w >= 0, "upfirdn2d: negative extent");
  VSP_REQUIRE(in_h < (1 << 30) && in_w < (1 << 30), "upfirdn2d: plane extent too large");
  VSP_REQUIRE(act == 0 || act == 3, "upfirdn2d: epilogue act must be 0 (none) or 3 (lrelu)");
  const int64_t out_h = vsp_upfirdn2d_out_size(in_h, kh, up_y, down_y, pad_y0, pad_y1);
  const int64_t out_w = vsp_upfirdn2d_out_size(in_w, kw, up_x, down_x, pad_x0, pad_x1);
  if (major == 0 || out_h <= 0 || out_w <= 0) return 0;  // empty output: nothing to do
  VSP_REQUIRE(x && filt && y, "upfirdn2d: null pointer");
  VSP_REQUIRE(out_h < (1 << 30) && out_w < (1 << 30), "upfirdn2d: output extent too large");
  if (act != 0 && bias != nullptr) VSP_REQUIRE(channels > 0, "upfirdn2d: channels must be > 0 with a bias");

  UfdParams p;
  p.x = x; p.filt = filt; p.y = y; p.bias = bias;
  p.major = major; p.in_h = (int)in_h; p.in_w = (int)in_w; p.out_h = (int)out_h; p.out_w = (int)out_w;
  p.kh = kh; p.kw = kw; p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.channels = (int)(channels > 0 ? channels : 1); p.act = act; p.alpha = alpha; p.scale = scale;
  p.tiles_x = p.tiles_y = p.zgroups = p.iters = 0;
  // TMA staging needs a 16-byte aligned base and row pitch, and a non-empty input
  static const bool no_tma = getenv("VSP_NO_TMA") != nullptr;  // debugging aid: force LDG staging
  p.use_tma = !no_tma && (in_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && in_h > 0 && in_w > 0 &&
              major < (1LL << 31);

  const bool small_filt = kh <= kK && kw <= kK;
  const bool sane_pad = pad_x0 > -(1 << 28) && pad_x0 < (1 << 28) && pad_y0 > -(1 << 28) && pad_y0 < (1 << 28);
  if (small_filt && sane_pad && in_h > 0 && in_w > 0) {
    if (up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1) {
      // contiguous row bands through 1-D bulk copies (any row pitch; VSP_BAND_ODD_ONLY=1 keeps even pitches on the tile kernel)
      static const bool no_band = getenv("VSP_NO_BAND") != nullptr;
      static const bool odd_only = getenv("VSP_BAND_ODD_ONLY") != nullptr;
      if ((!p.use_tma || !odd_only) && !no_band && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && out_w >= 32 &&
          major < (1LL << 31)) {
        const int rc = launch_band(p, 1, stream);
        if (rc >= 0) return rc;
      }
      return dispatch_tile<1, 1, 0, 0>(p, stream);
    }
    if (up_x == 1 && up_y == 1 && down_x == 2 && down_y == 2) {
      // decimation by 2 (Downsample, the backward of Upsample): the same row bands, two input rows / columns per output
      static const bool no_band2 = getenv("VSP_NO_BAND2") != nullptr;
      if (!no_band2 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && out_w >= 32 && major < (1LL << 31)) {
        const int rc = launch_band(p, 2, stream);
        if (rc >= 0) return rc;
      }
      return dispatch_tile<1, 2, 0, 0>(p, stream);
    }
    if (up_x == 2 && up_y == 2 && down_x == 1 && down_y == 1) {
      // tile origins are even, so the phase of the first tap is fixed by the pad parity
      const int qx = pmod(-pad_x0, 2), qy = pmod(-pad_y0, 2);
      if (qx == 0 && qy == 0) return dispatch_tile<2, 1, 0, 0>(p, stream);
      if (qx == 1 && qy == 0) return dispatch_tile<2, 1, 1, 0>(p, stream);
      if (qx == 0 && qy == 1) return dispatch_tile<2, 1, 0, 1>(p, stream);
      return dispatch_tile<2, 1, 1, 1>(p, stream);
    }
  }
  const long long total = (long long)major * out_h * out_w;
  long long blocks = (total + kThreads - 1) / kThreads;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  upfirdn2d_generic_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(p, total);
  return check_launch("upfirdn2d_generic_kernel");
}

extern "C" int vsp_upfirdn2d_nhwc_bf16(const void *x, const float *filt, void *y, int64_t n, int64_t in_h,
                                       int64_t in_w, int64_t c, int kh, int kw, int up_x, int up_y,
                                       int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                                       int pad_y1, const vsp_conv_epilogue *epi, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d_nhwc: factors must be >= 1");
  VSP_REQUIRE(kh >= 1 && kw >= 1, "upfirdn2d_nhwc: empty filter");
  VSP_REQUIRE(c >= 0 && c % 8 == 0, "upfirdn2d_nhwc: channels must be a multiple of 8");
  VSP_REQUIRE(in_h < (1 << 30) && in_w < (1 << 30), "upfirdn2d_nhwc: extent too large");
  const int64_t out_h = vsp_upfirdn2d_out_size(in_h, kh, up_y, down_y, pad_y0, pad_y1);
  const int64_t out_w = vsp_upfirdn2d_out_size(in_w, kw, up_x, down_x, pad_x0, pad_x1);
  if (n <= 0 || c == 0 || out_h <= 0 || out_w <= 0) return 0;
  VSP_REQUIRE(x && filt && y, "upfirdn2d_nhwc: null pointer");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
              "upfirdn2d_nhwc: tensors must be 16-byte aligned");
  UfdParams p;
  memset(&p, 0, sizeof(p));
  p.in_h = (int)in_h; p.in_w = (int)in_w; p.out_h = (int)out_h; p.out_w = (int)out_w;
  p.kh = kh; p.kw = kw; p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  const int cg = (int)(c / 8);
  const long long total = (long long)n * out_h * out_w * cg;
  long long blocks = (total + kThreads - 1) / kThreads;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  NhwcEpi e;
  memset(&e, 0, sizeof(e));
  if (epi) {
    VSP_REQUIRE(epi->act == 0 || epi->act == 3, "upfirdn2d_nhwc: epilogue act must be 0 or 3");
    VSP_REQUIRE(epi->row_scale == nullptr && epi->pre_act == 0, "upfirdn2d_nhwc: row_scale / pre_act are conv-only");
    e.noise = epi->noise; e.noise_bstride = epi->noise_bstride; e.noise_weight = epi->noise_weight;
    e.noise_weight_dev = epi->noise_weight_dev; e.bias = epi->bias; e.act = epi->act; e.alpha = epi->alpha;
    e.scale = epi->scale;
    e.residual = static_cast<const uint4 *>(epi->residual);
    e.residual2 = static_cast<const uint4 *>(epi->residual2);
  }
  if (up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && kh <= kK && kw <= kK) {
    static const int blur_r = getenv("VSP_BLUR_R") ? atoi(getenv("VSP_BLUR_R")) : 4;
    const int R = blur_r == 8 ? 8 : (blur_r == 2 ? 2 : 4);
    const int row_blocks = (int)((out_h + R - 1) / R);
    const long long tot = (long long)n * row_blocks * out_w * cg;
    long long nb = (tot + kThreads - 1) / kThreads;
    if (nb > (long long)num_sms() * 64) nb = (long long)num_sms() * 64;
    if (R == 8)
      blur_nhwc_kernel<8><<<(unsigned)nb, kThreads, 0, stream>>>(static_cast<const uint4 *>(x), filt,
                                                                static_cast<uint4 *>(y), p, e, cg, row_blocks, tot);
    else if (R == 2)
      blur_nhwc_kernel<2><<<(unsigned)nb, kThreads, 0, stream>>>(static_cast<const uint4 *>(x), filt,
                                                                static_cast<uint4 *>(y), p, e, cg, row_blocks, tot);
    else
      blur_nhwc_kernel<4><<<(unsigned)nb, kThreads, 0, stream>>>(static_cast<const uint4 *>(x), filt,
                                                                static_cast<uint4 *>(y), p, e, cg, row_blocks, tot);
    return check_launch("blur_nhwc_kernel");
  }
  upfirdn2d_nhwc_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(
      static_cast<const uint4 *>(x), filt, static_cast<uint4 *>(y), p, e, cg, total);
  return check_launch("upfirdn2d_nhwc_kernel");
}

extern "C" int vsp_blur_sep_nhwc_bf16(const void *x, const float *fy_host, const float *fx_host, void *y, int64_t n,
                                      int64_t in_h, int64_t in_w, int64_t c, int kh, int kw, int pad_x0, int pad_x1,
                                      int pad_y0, int pad_y1, const vsp_conv_epilogue *epi, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(kh >= 1 && kw >= 1 && kh <= kK && kw <= kK, "blur_sep_nhwc: filter up to 4x4");
  VSP_REQUIRE(c >= 0 && c % 8 == 0, "blur_sep_nhwc: channels must be a multiple of 8");
  VSP_REQUIRE(in_h < (1 << 30) && in_w < (1 << 30), "blur_sep_nhwc: extent too large");
  const int64_t out_h = vsp_upfirdn2d_out_size(in_h, kh, 1, 1, pad_y0, pad_y1);
  const int64_t out_w = vsp_upfirdn2d_out_size(in_w, kw, 1, 1, pad_x0, pad_x1);
  if (n <= 0 || c == 0 || out_h <= 0 || out_w <= 0) return 0;
  VSP_REQUIRE(x && y && fy_host && fx_host, "blur_sep_nhwc: null pointer");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
              "blur_sep_nhwc: tensors must be 16-byte aligned");
  UfdParams p;
  memset(&p, 0, sizeof(p));
  p.in_h = (int)in_h; p.in_w = (int)in_w; p.out_h = (int)out_h; p.out_w = (int)out_w;
  p.kh = kh; p.kw = kw; p.up_x = p.up_y = p.down_x = p.down_y = 1;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  SepTaps t;
  // true convolution: tap j of the window multiplies the flipped filter entry
  for (int j = 0; j < kK; ++j) {
    t.fy[j] = j < kh ? fy_host[kh - 1 - j] : 0.f;
    t.fx[j] = j < kw ? fx_host[kw - 1 - j] : 0.f;
  }
  NhwcEpi e;
  memset(&e, 0, sizeof(e));
  if (epi) {
    VSP_REQUIRE(epi->act == 0 || epi->act == 3, "blur_sep_nhwc: epilogue act must be 0 or 3");
    VSP_REQUIRE(epi->row_scale == nullptr && epi->pre_act == 0, "blur_sep_nhwc: row_scale / pre_act are conv-only");
    e.noise = epi->noise; e.noise_bstride = epi->noise_bstride; e.noise_weight = epi->noise_weight;
    e.noise_weight_dev = epi->noise_weight_dev; e.bias = epi->bias; e.act = epi->act; e.alpha = epi->alpha;
    e.scale = epi->scale;
    e.residual = static_cast<const uint4 *>(epi->residual);
    e.residual2 = static_cast<const uint4 *>(epi->residual2);
  }
  static const bool strip_on = getenv("VSP_NO_BLUR_STRIP") == nullptr;
  if (strip_on && c % 64 == 0 && in_w >= kStripCols && in_h >= kStageRows) {
    StripParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.in_h = (int)in_h; sp.in_w = (int)in_w; sp.out_h = (int)out_h; sp.out_w = (int)out_w;
    sp.pad_x0 = pad_x0; sp.pad_y0 = pad_y0;
    const bool has_epi = e.noise || e.bias || e.act != 0 || e.residual || e.residual2;
    if (!has_epi) return launch_blur_strip<false>(x, y, sp, e, t, n, c, stream);
    // Epilogue form: at 3 CTAs/SM (168 registers) ptxas spills 124 B per thread INSIDE the row loop — ncu showed more local-
    // memory requests (2.1 M loads + 2.2 M stores, 16 % L1 hits) than global loads (1.6 M) and the kernel at 2.9 TB/s.  With 2
    // CTAs/SM (204 registers, no spills) and the operand loads three rows ahead: [32,129,129,256] 365 -> 312 us (3.5 TB/s).
    // VSP_BLUR_EPI_BLOCKS=3 / VSP_BLUR_EPI_AHEAD=0..3 select the other forms (tools/bench_blur_epi.py).
    // operands of the epilogue on the bulk-copy path (blur_strip_epi_kernel); VSP_BLUR_EPI_TMA=0 keeps the register ring
    static const bool epi_tma = getenv("VSP_BLUR_EPI_TMA") == nullptr || atoi(getenv("VSP_BLUR_EPI_TMA")) != 0;
    if (epi_tma) {
      const void *r1 = e.residual ? (const void *)e.residual : (const void *)e.residual2;
      const void *r2 = e.residual ? (const void *)e.residual2 : nullptr;
      const int nres = (r1 != nullptr) + (r2 != nullptr);
      const bool aligned = (reinterpret_cast<uintptr_t>(r1) & 15) == 0 && (reinterpret_cast<uintptr_t>(r2) & 15) == 0;
      if (aligned) {
        if (nres == 2) return launch_blur_strip_epi<2>(x, y, sp, e, r1, r2, t, n, c, stream);
        if (nres == 1) return launch_blur_strip_epi<1>(x, y, sp, e, r1, nullptr, t, n, c, stream);
        return launch_blur_strip_epi<0>(x, y, sp, e, nullptr, nullptr, t, n, c, stream);
      }
    }
    static const int ahead = getenv("VSP_BLUR_EPI_AHEAD") ? atoi(getenv("VSP_BLUR_EPI_AHEAD")) : 3;
    static const int minb = getenv("VSP_BLUR_EPI_BLOCKS") ? atoi(getenv("VSP_BLUR_EPI_BLOCKS")) : 2;
    if (minb == 3)
      return ahead == 0 ? launch_blur_strip<true, 0, 3>(x, y, sp, e, t, n, c, stream)
                        : launch_blur_strip<true, 2, 3>(x, y, sp, e, t, n, c, stream);
    switch (ahead) {
      case 0: return launch_blur_strip<true, 0, 2>(x, y, sp, e, t, n, c, stream);
      case 1: return launch_blur_strip<true, 1, 2>(x, y, sp, e, t, n, c, stream);
      case 2: return launch_blur_strip<true, 2, 2>(x, y, sp, e, t, n, c, stream);
      default: return launch_blur_strip<true, 3, 2>(x, y, sp, e, t, n, c, stream);
    }
  }
  constexpr int R = 4;
  const int cg = (int)(c / 8);
  const int col_pairs = (int)((out_w + 1) / 2), row_blocks = (int)((out_h + R - 1) / R);
  const long long tot = (long long)n * row_blocks * col_pairs * cg;
  long long nb = (tot + kThreads - 1) / kThreads;
  if (nb > (long long)num_sms() * 64) nb = (long long)num_sms() * 64;
  blur_sep_nhwc_kernel<R><<<(unsigned)nb, kThreads, 0, stream>>>(static_cast<const uint4 *>(x), static_cast<uint4 *>(y), p,
                                                                e, t, cg, col_pairs, row_blocks, tot);
  return check_launch("blur_sep_nhwc_kernel");
}
