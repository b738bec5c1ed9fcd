// conv_sm100.cu — 2-D convolution as a bf16 implicit GEMM on tcgen05 / TMEM, fed by TMA.
//
// What it replaces: the cuDNN grouped convolution the reference reaches through
// conv2d_gradfix.conv2d / conv_transpose2d (op/conv2d_gradfix.py:22-75) from
// ModulatedConv2d (models/RestoreNet.py:522-553) and EqualConv2d (:125-131).
//
// Formulation ("gather convolution"):
//   out[b, oh, ow, n] = sum_{t < ntaps} sum_c x[b, oh*s + dy[t], ow*s + dx[t], c] * wq[g, tw[t], n, c]
// with g = b (per-sample modulated weights) or 0 (shared weights).  The host expresses
// stride-1/2 convs with dilation, their input gradients, and the four parity classes of a
// stride-2 transposed conv (no zero-stuffing) as tap lists over the same kernel.
//
// GEMM view: M = 128 output pixels of ONE sample (a th x tw patch), N = BLOCK_N output
// channels, K = ntaps * Cin walked in 64-channel blocks.  A (activations, NHWC bf16) and
// B (weights, [g][tap][n][c] bf16) tiles are K-major 128B-swizzled boxes written by TMA;
// the im2col gather is a 4-D box whose spatial coordinates are shifted by the tap offset —
// out-of-bounds rows/cols/channels are zero-filled by the TMA unit, which is exactly the
// convolution's zero padding.  Accumulators live in TMEM (double-buffered so the epilogue of
// tile i overlaps the MMAs of tile i+1).  Warp roles: warp0 = TMA producer, warp1 = MMA
// issuer (one elected lane), warps2-5 = epilogue (tcgen05.ld -> demod/noise/bias/lrelu/
// residual -> global).  Persistent: one CTA per SM walks the tile list.
#include "conv_common.cuh"

#include <stdlib.h>

namespace vsp {
namespace {

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int B_BYTES = BLOCK_N * kBlockK * 2;
  static constexpr int STAGE_BYTES = kABytes + B_BYTES;
  static constexpr int STAGES = (BLOCK_N >= 256) ? 4 : ((BLOCK_N >= 128) ? 5 : (BLOCK_N >= 64 ? 7 : 8));
  static constexpr int TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
  static constexpr int CHUNK = BLOCK_N < 32 ? BLOCK_N : 32;
  static constexpr int SBUF_BYTES = 32 * CHUNK * 2;          // staged epilogue: 32 pixels x CHUNK channels bf16
  static constexpr int NBUF = BLOCK_N >= 256 ? 1 : 2;         // staging buffers per epilogue warp (smem budget)
  static constexpr int OSTAGE_BYTES = 8 * NBUF * SBUF_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 8 * BLOCK_N * 4 /*epilogue vectors*/;
  static constexpr int SMEM_BYTES_STAGED = SMEM_BYTES + OSTAGE_BYTES + 1024;
  static_assert(SMEM_BYTES_STAGED <= 232448, "conv_fprop_kernel exceeds shared memory");
};

// Epilogue role (warps 2..5 of either kernel): TMEM -> registers -> fused epilogue -> global.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_role(const ConvParams &p, float *epi_vec, uint64_t *tmem_full,
                                              uint64_t *tmem_empty, uint32_t tmem_base, int warp, int lane) {
  using C = ConvCfg<BLOCK_N>;
  // Per-channel vectors (demod, biases) are staged in shared memory once per tile by the 128
  // epilogue threads (coalesced, issued BEFORE waiting for the accumulator so the latency
  // hides behind the MMAs) and then read back as 128-bit broadcasts; accumulators leave TMEM
  // 32 columns at a time.
  const int quad = warp & 3;              // TMEM lane quadrant this warp may access
  const int wg = (warp - 2) >> 2;         // the two warps of a quadrant alternate accumulator chunks
  const int row = quad * 32 + lane;       // pixel row inside the tile
  const int et = threadIdx.x - 64;        // 0..255 among the epilogue threads
  float *vec_rs = epi_vec;                // [2][BLOCK_N] demod
  float *vec_b1 = epi_vec + 2 * BLOCK_N;  // [2][BLOCK_N] pre-activation bias (stage 1)
  float *vec_b2 = epi_vec + 4 * BLOCK_N;  // [2][BLOCK_N] bias (stage 2)
  const float nw = p.noise ? (p.noise_weight_dev ? __ldg(p.noise_weight_dev) : p.noise_weight) : 0.f;
  int acc = 0;
  uint32_t acc_phase = 0;
  const int creal = p.shuffle_cout ? p.shuffle_cout : p.cout;
  const long long plane = (long long)p.full_h * p.full_w;
  for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    long long t = tile;
    const int n_i = tile_n_index(p, tile); t /= p.tiles_n;
    const int w_i = (int)(t % p.tiles_w); t /= p.tiles_w;
    const int h_i = (int)(t % p.tiles_h); t /= p.tiles_h;
    const int sp = p.th * p.tw;                                    // pixels of one sample inside the tile
    const int b = (int)t * p.tb + row / sp;                        // stacked tiles: rows [k*sp, (k+1)*sp) = sample k
    const int rem = row % sp;
    const int oh = h_i * p.th + rem / p.tw;
    const int ow = w_i * p.tw + rem % p.tw;
    const bool pix_ok = oh < p.out_h && ow < p.out_w && b < p.batch;
    const int nbase = n_i * BLOCK_N;

    // stage the channel vectors of this tile (demod only when the whole tile belongs to one sample)
    for (int c = et; c < BLOCK_N; c += 256) {
      const int n = nbase + c;
      const int cr = p.shuffle_cout ? n % p.shuffle_cout : n;
      const bool ok = n < p.cout;
      vec_rs[acc * BLOCK_N + c] = (ok && p.row_scale && p.tb == 1) ? __ldg(p.row_scale + (long long)b * creal + cr) : 1.f;
      vec_b1[acc * BLOCK_N + c] = (ok && p.pre_bias) ? __ldg(p.pre_bias + cr) : 0.f;
      vec_b2[acc * BLOCK_N + c] = (ok && p.bias) ? __ldg(p.bias + cr) : 0.f;
    }
    float nz0 = 0.f;
    if (p.noise != nullptr && pix_ok && !p.shuffle_cout)
      nz0 = __ldg(p.noise + b * p.noise_bstride + (long long)(oh * p.os + p.oo_h) * p.full_w + ow * p.os + p.oo_w);   // raw: first use after the wait
    asm volatile("bar.sync 1, 256;" ::: "memory");   // epilogue warps only

    mbar_wait(&tmem_full[acc], acc_phase);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
    for (int ch = wg; ch < BLOCK_N / C::CHUNK; ch += 2) {
      const int n0 = nbase + ch * C::CHUNK;
      const int cls = p.shuffle_cout ? n0 / p.shuffle_cout : 0;
      const int c0 = p.shuffle_cout ? n0 % p.shuffle_cout : n0;
      const int ctot = n0 < p.cout ? creal : 0;                   // padded tail of the channel tile: nothing to write
      const long long pix = (long long)(oh * p.os + p.oo_h + (cls >> 1)) * p.full_w + ow * p.os + p.oo_w + (cls & 1);
      float nz = nw * nz0;
      if (p.shuffle_cout && p.noise != nullptr && pix_ok) nz = nw * __ldg(p.noise + b * p.noise_bstride + pix);
      const float *rs_row = (p.tb > 1 && p.row_scale && pix_ok) ? p.row_scale + (long long)b * creal + c0 : nullptr;
      epi_chunk<C::CHUNK>(p, taddr + ch * C::CHUNK, c0, ctot, b, pix, plane, pix_ok, nz,
                          vec_rs + acc * BLOCK_N + ch * C::CHUNK, vec_b1 + acc * BLOCK_N + ch * C::CHUNK,
                          vec_b2 + acc * BLOCK_N + ch * C::CHUNK, rs_row);
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}

// Staged epilogue role: same arithmetic (branch-free form), but each warp stages its 32 pixels x CHUNK channels in
// swizzled shared memory and writes them with one TMA store (full lines instead of 16-byte pieces at pixel
// pitch).  Handles the parity-class output mappings (one tensor map per class) and the pixel-shuffle mapping of
// the fused up-convolution.  Requires tw >= 32 (a warp's pixels lie in one output row).
// PAIR: the CTA-pair kernel — tiles come in pairs (spatial tiles 2*mp + cluster rank of one sample and channel tile) and
// the accumulator is handed back to the LEADER's barrier (a remote arrive for the peer CTA).
template <int BLOCK_N, bool PAIR = false>
__device__ __forceinline__ void epilogue_role_staged(const ConvParams &p, const OutMaps &maps, float *epi_vec,
                                                     unsigned char *o_buf, uint64_t *tmem_full, uint64_t *tmem_empty,
                                                     uint32_t tmem_base, int warp, int lane, int rank = 0) {
  using C = ConvCfg<BLOCK_N>;
  constexpr int CHUNK = C::CHUNK;
  constexpr int ROW_BYTES = CHUNK * 2;
  const int quad = warp & 3;
  const int wg = (warp - 2) >> 2;         // the two warps of a quadrant alternate accumulator chunks
  const int row = quad * 32 + lane;
  const int et = threadIdx.x - 64;
  float *vec_rs = epi_vec, *vec_b1 = epi_vec + 2 * BLOCK_N, *vec_b2 = epi_vec + 4 * BLOCK_N, *vec_a = epi_vec + 6 * BLOCK_N;
  const float nw = p.noise ? (p.noise_weight_dev ? __ldg(p.noise_weight_dev) : p.noise_weight) : 0.f;
  const LeanK lk = lean_consts(p);
  const bool prelu = p.alpha_vec != nullptr && p.act != 0 && p.pre_act == 0;     // per-channel slope (host guarantees the form)
  const int sw = ROW_BYTES == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1);
  unsigned char *stage = o_buf + (warp - 2) * C::NBUF * C::SBUF_BYTES;
  const long long plane = (long long)p.full_h * p.full_w;
  const int creal = p.shuffle_cout ? p.shuffle_cout : p.cout;
  const __nv_bfloat16 *res1 = static_cast<const __nv_bfloat16 *>(p.residual);
  const __nv_bfloat16 *res2 = static_cast<const __nv_bfloat16 *>(p.residual2);
  int acc = 0;
  uint32_t acc_phase = 0, sbuf = 0;
  auto release_acc = [&](int a) {               // hand accumulator stage `a` back to the MMA warp
    if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[a]), 0));
    else mbar_arrive(&tmem_empty[a]);
  };
  const long long t_begin = PAIR ? (blockIdx.x >> 1) : blockIdx.x, t_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const long long t_end = PAIR ? p.total_pairs : p.total_tiles;
  for (long long tile = t_begin; tile < t_end; tile += t_step) {
    long long t = tile;
    const int n_i = tile_n_index(p, tile); t /= p.tiles_n;
    int w_i, h_i;
    if (PAIR) {
      const int m = 2 * (int)(t % p.mpairs) + rank; t /= p.mpairs;
      w_i = m % p.tiles_w; h_i = m / p.tiles_w;          // h_i == tiles_h for the odd tail: every pixel out of range
    } else {
      w_i = (int)(t % p.tiles_w); t /= p.tiles_w;
      h_i = (int)(t % p.tiles_h); t /= p.tiles_h;
    }
    const int b = (int)t;
    const int oh = h_i * p.th + row / p.tw;
    const int ow = w_i * p.tw + row % p.tw;
    const bool pix_ok = oh < p.out_h && ow < p.out_w;
    const int oh_w = h_i * p.th + (quad * 32) / p.tw, ow_w = w_i * p.tw + (quad * 32) % p.tw;   // warp's first pixel
    const int nbase = n_i * BLOCK_N;
    for (int c = et; c < BLOCK_N; c += 256) {
      const int n = nbase + c;
      const int cr = p.shuffle_cout ? n % p.shuffle_cout : n;
      const bool ok = n < p.cout;
      vec_rs[acc * BLOCK_N + c] = lean_scale_rs(lk, (ok && p.row_scale) ? __ldg(p.row_scale + (long long)b * creal + cr) : 1.f);
      vec_b1[acc * BLOCK_N + c] = lean_scale_b1(lk, (ok && p.pre_bias) ? __ldg(p.pre_bias + cr) : 0.f);
      vec_b2[acc * BLOCK_N + c] = lean_scale_b2(lk, (ok && p.bias) ? __ldg(p.bias + cr) : 0.f);
      if (prelu) vec_a[acc * BLOCK_N + c] = ok ? __ldg(p.alpha_vec + cr) : 1.f;
    }
    // noise of this thread's pixel in every output class of the tile, fetched before the accumulator wait
    float nzc[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.noise != nullptr && pix_ok) {
      const float *nb = p.noise + b * p.noise_bstride;
      if (p.shuffle_cout) {
        const int c_lo = nbase / p.shuffle_cout, c_hi = min(3, (nbase + BLOCK_N - 1) / p.shuffle_cout);
#pragma unroll
        for (int cls = 0; cls < 4; ++cls)
          if (cls >= c_lo && cls <= c_hi)
            nzc[cls] = __ldg(nb + (long long)(oh * 2 + (cls >> 1)) * p.full_w + ow * 2 + (cls & 1));   // raw, scaled at first use
      } else {
        nzc[0] = __ldg(nb + (long long)(oh * p.os + p.oo_h) * p.full_w + ow * p.os + p.oo_w);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // epilogue warps only

    constexpr int NCH = BLOCK_N / CHUNK;
    constexpr int PIECES = CHUNK / 8;                 // 16-byte pieces per pixel of a chunk
    constexpr int PPR = 32 / PIECES;                  // pixels covered by one warp-wide 128-bit request
    const bool has_res = res1 != nullptr || res2 != nullptr;
    // residuals: warp-cooperative loads (lane -> piece lane % PIECES of pixel lane / PIECES: whole sectors per
    // request), prefetched one chunk ahead and transposed to pixel-per-lane through the staging buffer
    uint4 pr1[PIECES], pr2[PIECES];
    auto load_res = [&](int ch) {
      const int n0 = nbase + ch * CHUNK;
      const int cls = p.shuffle_cout ? n0 / p.shuffle_cout : 0;
      const int c0 = p.shuffle_cout ? n0 % p.shuffle_cout : n0;
      const int fh = oh_w * p.os + p.oo_h + (cls >> 1);
#pragma unroll
      for (int i = 0; i < PIECES; ++i) {
        const int px = i * PPR + lane / PIECES;
        const int fw = (ow_w + px) * p.os + p.oo_w + (cls & 1);
        const bool ok = n0 < p.cout && oh_w < p.out_h && ow_w + px < p.out_w;
        const long long off = (((long long)b * p.full_h + fh) * p.full_w + fw) * p.ldo + p.co_off + c0 + (lane % PIECES) * 8;
        pr1[i] = (res1 && ok) ? __ldg(reinterpret_cast<const uint4 *>(res1 + off)) : make_uint4(0u, 0u, 0u, 0u);
        pr2[i] = (res2 && ok) ? __ldg(reinterpret_cast<const uint4 *>(res2 + off)) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    if (has_res && wg < NCH) load_res(wg);

    mbar_wait(&tmem_full[acc], acc_phase);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
    if (wg >= NCH) {                 // single-chunk tiles: the second warp of the quadrant has nothing to read
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(acc);
    }
#pragma unroll 1
    for (int ch = wg; ch < NCH; ch += 2) {
      const int n0 = nbase + ch * CHUNK;
      const bool last_ch = ch + 2 >= NCH;
      if (n0 >= p.cout) {            // warp-uniform: nothing to write for the padded tail of the channel tile
        if (last_ch) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
        }
        continue;
      }
      const int cls = p.shuffle_cout ? n0 / p.shuffle_cout : 0;
      const int c0 = p.shuffle_cout ? n0 % p.shuffle_cout : n0;
      float nz = nzc[0];
#pragma unroll
      for (int q = 1; q < 4; ++q) nz = cls == q ? nzc[q] : nz;
      uint32_t r[CHUNK];
      if constexpr (CHUNK == 32) tmem_ld_32x32b_x32(taddr + ch * CHUNK, r);
      else tmem_ld_32x32b_x16(taddr + ch * CHUNK, reinterpret_cast<uint32_t(&)[16]>(r));
      tmem_ld_wait();
      if (last_ch) {                 // accumulators are in registers: the MMA warp may reuse this stage
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
      }
      unsigned char *buf = stage + (C::NBUF == 2 ? (sbuf & 1) : 0) * C::SBUF_BYTES;
      ++sbuf;
      if (lane == 0) bulk_wait_group_read<C::NBUF - 1>();
      __syncwarp();
      uint4 own1[PIECES], own2[PIECES];
      if (has_res) {
        // transpose the prefetched pieces: [pixel][piece] in the (swizzled) staging tile, then read the own row
#pragma unroll
        for (int i = 0; i < PIECES; ++i) {
          const int px = i * PPR + lane / PIECES;
          const int swp = ROW_BYTES == 64 ? ((px >> 1) & 3) : ((px >> 2) & 1);
          *reinterpret_cast<uint4 *>(buf + px * ROW_BYTES + (((lane % PIECES) ^ swp) << 4)) = pr1[i];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PIECES; ++i) own1[i] = *reinterpret_cast<const uint4 *>(buf + lane * ROW_BYTES + ((i ^ sw) << 4));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PIECES; ++i) {
          const int px = i * PPR + lane / PIECES;
          const int swp = ROW_BYTES == 64 ? ((px >> 1) & 3) : ((px >> 2) & 1);
          *reinterpret_cast<uint4 *>(buf + px * ROW_BYTES + (((lane % PIECES) ^ swp) << 4)) = pr2[i];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PIECES; ++i) own2[i] = *reinterpret_cast<const uint4 *>(buf + lane * ROW_BYTES + ((i ^ sw) << 4));
        __syncwarp();
        if (ch + 2 < NCH) load_res(ch + 2);
      }
      const float *vrs = vec_rs + acc * BLOCK_N + ch * CHUNK, *vb1 = vec_b1 + acc * BLOCK_N + ch * CHUNK,
                  *vb2 = vec_b2 + acc * BLOCK_N + ch * CHUNK;
#pragma unroll
      for (int i = 0; i < PIECES; ++i) {
        float v[8];
        if (prelu) epi_lean8f_prelu(&r[8 * i], vrs + 8 * i, vb2 + 8 * i, vec_a + acc * BLOCK_N + ch * CHUNK + 8 * i, nw * nz * lk.m2, v);
        else epi_lean8f(&r[8 * i], vrs + 8 * i, vb1 + 8 * i, vb2 + 8 * i, nw * nz * lk.m2, lk, v);
        if (has_res) add2_bf16x8(v, own1[i], own2[i]);
        *reinterpret_cast<uint4 *>(buf + lane * ROW_BYTES + ((i ^ sw) << 4)) = pack8_bf16(v);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&maps.m[cls], buf, (int)p.co_off + c0, ow_w, oh_w, b);
        bulk_commit_group();
      }
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
  if (lane == 0) bulk_wait_group<0>();
}

template <int BLOCK_N, bool STAGED>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_fprop_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ OutMaps omaps) {
  using C = ConvCfg<BLOCK_N>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t *empty_bar = full_bar + C::STAGES;
  uint64_t *tmem_full = empty_bar + C::STAGES;
  uint64_t *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  float *epi_vec = reinterpret_cast<float *>(smem + C::STAGES * C::STAGE_BYTES + 256);  // 8 * BLOCK_N floats
  unsigned char *o_buf = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(epi_vec + 8 * BLOCK_N) + 1023) & ~uintptr_t(1023));   // STAGED only

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.ntaps * p.kc, num_kb_all = num_kb;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long t = tile;
      const int n_i = tile_n_index(p, tile); t /= p.tiles_n;
      const int w_i = (int)(t % p.tiles_w); t /= p.tiles_w;
      const int h_i = (int)(t % p.tiles_h); t /= p.tiles_h;
      const int b = (int)t * p.tb;                       // first sample of the (possibly stacked) tile
      const int g = p.groups == 1 ? 0 : b;
      const int iw0 = w_i * p.tw * p.stride, ih0 = h_i * p.th * p.stride;
      const int dmul = p.branch_mode ? p.n_dil[n_i] : 1;
      const int cls = p.cls_mode ? n_i / p.cls_tpc : 0;
      const int tile_kb = p.cls_mode ? p.cls_ntaps[cls] * p.kc : num_kb;
      const int brow = (p.cls_mode ? n_i % p.cls_tpc : n_i) * BLOCK_N;
      for (int kb = 0; kb < tile_kb; ++kb) {
        const int tap = kb / p.kc;
        const int c0 = (kb - tap * p.kc) * kBlockK;
        const int ts = p.cls_mode ? p.cls_shift[cls][tap] : tap;      // activation shift of this tap
        const int tw = p.cls_mode ? p.cls_w[cls][tap] : p.tap_w[tap];  // weight tap
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          unsigned char *sa = smem + stage * C::STAGE_BYTES;
          unsigned char *sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_4d(sa, &tmap_a, &full_bar[stage], c0, iw0 + p.tap_dx[ts] * dmul, ih0 + p.tap_dy[ts] * dmul, b);
          tma_load_4d(sb, &tmap_b, &full_bar[stage], c0, brow, tw, g);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (same scheme) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int num_kb = p.cls_mode ? p.cls_ntaps[tile_n_index(p, tile) / p.cls_tpc] * p.kc : num_kb_all;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = umma_smem_desc(sa, 128);
          const uint64_t bdesc = umma_smem_desc(sa + kABytes, 128);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    if constexpr (STAGED)
      epilogue_role_staged<BLOCK_N>(p, omaps, epi_vec, o_buf, tmem_full, tmem_empty, tmem_base, warp, lane);
    else
      epilogue_role<BLOCK_N>(p, epi_vec, tmem_full, tmem_empty, tmem_base, warp, lane);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair form of the N = 256 kernel (tcgen05 cta_group::2).  The single-CTA mainloop above is bound by load LATENCY, not by
// the tensor pipe: a stage is 48 KB (16 KB of A + 32 KB of B), four stages fill the shared memory and cover only ~2000
// cycles of L2 latency at the rate the MMAs consume them (ncu: tensor pipe 69 % active, shared-memory operand path 52 %,
// L2 39 %).  Two CTAs on the SMs of one TPC that work on neighbouring spatial tiles of the same sample and channel tile need
// the same weights: as a pair they issue ONE M = 256 MMA per k-step, each CTA supplying its own 128 activation rows and HALF
// of the weight rows.  A stage shrinks to 32 KB, six stages fit, and the same shared memory covers 1.5x the latency; the
// weight tile crosses L2 -> SM once per pair instead of once per CTA.
// Protocol: both producers wait for their OWN empty barrier (the leader's tcgen05.commit multicasts its arrive to both
// CTAs) and issue TMA loads that complete on the LEADER's full barrier (one expect_tx of both CTAs' bytes by the leader);
// only the leader's warp 1 issues MMAs; accumulator-ready is multicast to both CTAs' epilogues; accumulator-free is 16
// arrives (8 epilogue warps x 2 CTAs, the peer's remote) on the leader's barrier.
template <int BLOCK_N>
struct PairCfg {
  static constexpr int B_HALF_BYTES = (BLOCK_N / 2) * kBlockK * 2;
  static constexpr int STAGE_BYTES = kABytes + B_HALF_BYTES;
  static constexpr int STAGES = BLOCK_N >= 256 ? 6 : 7;
  static constexpr int OSTAGE_BYTES = ConvCfg<BLOCK_N>::OSTAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 8 * BLOCK_N * 4 + OSTAGE_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "conv_fprop_pair_kernel exceeds shared memory");
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
conv_fprop_pair_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap_a,
                       const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ OutMaps omaps) {
  using P = PairCfg<BLOCK_N>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + P::STAGES * P::STAGE_BYTES);
  uint64_t *empty_bar = full_bar + P::STAGES;
  uint64_t *tmem_full = empty_bar + P::STAGES;
  uint64_t *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  float *epi_vec = reinterpret_cast<float *>(smem + P::STAGES * P::STAGE_BYTES + 256);
  unsigned char *o_buf = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(epi_vec + 8 * BLOCK_N) + 1023) & ~uintptr_t(1023));

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < P::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 16);   // 8 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, 2 * BLOCK_N);
    tmem_relinquish_2cta();
  }
  tcgen05_fence_before();
  cluster_sync_all();                  // barriers of BOTH CTAs are initialised before any remote arrive / TMA completion
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.ntaps * p.kc, num_kb_all = num_kb;
  const long long t_begin = blockIdx.x >> 1, t_step = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = t_begin; tile < p.total_pairs; tile += t_step) {
      long long t = tile;
      const int n_i = tile_n_index(p, tile); t /= p.tiles_n;
      const int m = 2 * (int)(t % p.mpairs) + rank; t /= p.mpairs;
      const int w_i = m % p.tiles_w, h_i = m / p.tiles_w;
      const int b = (int)t;
      const int g = p.groups == 1 ? 0 : b;
      const int iw0 = w_i * p.tw * p.stride, ih0 = h_i * p.th * p.stride;
      const int dmul = p.branch_mode ? p.n_dil[n_i] : 1;     // channel tile = SMART branch: its own dilation
      const int cls = p.cls_mode ? n_i / p.cls_tpc : 0;
      const int tile_kb = p.cls_mode ? p.cls_ntaps[cls] * p.kc : num_kb;
      const int brow = (p.cls_mode ? n_i % p.cls_tpc : n_i) * BLOCK_N + rank * (BLOCK_N / 2);
      for (int kb = 0; kb < tile_kb; ++kb) {
        const int tap = kb / p.kc;
        const int c0 = (kb - tap * p.kc) * kBlockK;
        const int ts = p.cls_mode ? p.cls_shift[cls][tap] : tap;      // activation shift of this tap
        const int tw = p.cls_mode ? p.cls_w[cls][tap] : p.tap_w[tap];  // weight tap
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          unsigned char *sa = smem + stage * P::STAGE_BYTES;
          unsigned char *sb = sa + kABytes;
          const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * P::STAGE_BYTES);
          tma_load_4d_2cta(sa, &tmap_a, lead_full, c0, iw0 + p.tap_dx[ts] * dmul, ih0 + p.tap_dy[ts] * dmul, b);
          tma_load_4d_2cta(sb, &tmap_b, lead_full, c0, brow, tw, g);
        }
        __syncwarp();
        if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader only): M = 256 across the pair =====================
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long tile = t_begin; tile < p.total_pairs; tile += t_step) {
        const int num_kb = p.cls_mode ? p.cls_ntaps[tile_n_index(p, tile) / p.cls_tpc] * p.kc : num_kb_all;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * P::STAGE_BYTES);
            const uint64_t adesc = umma_smem_desc(sa, 128);
            const uint64_t bdesc = umma_smem_desc(sa + kABytes, 128);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_bf16_ss_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_2cta(&empty_bar[stage]);                         // frees the slot in BOTH CTAs
            if (kb == num_kb - 1) umma_commit_2cta(&tmem_full[acc]);     // accumulators complete in both CTAs
          }
          __syncwarp();
          if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    epilogue_role_staged<BLOCK_N, true>(p, omaps, epi_vec, o_buf, tmem_full, tmem_empty, tmem_base, warp, lane, rank);
  }

  tcgen05_fence_before();
  cluster_sync_all();                  // both CTAs are done with the pair's tensor memory and barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2cta(tmem_base, 2 * BLOCK_N);
  }
}

// ---------------------------------------------------------------------------------------------
// Row-halo variant for stride-1 3x3 (dilated) convolutions on wide images (out_w >= 128).
// The plain implicit GEMM re-reads every input pixel once per tap (9x through L2), which makes the
// C <= 128 layers at 128^2..1024^2 L2-bound.  Here a tile is 128 consecutive pixels of ONE output row
// and the three input rows it needs (128 + 2*dil pixels each, x 64 channels) are TMA-loaded ONCE per
// channel block; the nine taps are nine row-shifted UMMA descriptors into that halo (start address
// + (kh*halo_w + kw*dil) * 128 B; the swizzle XOR is absolute-address based, base offset stays 0).
// Weights: RESIDENT_B keeps all nine [N x 64] tap tiles in smem across the tiles of a sample (Cin <= 64);
// otherwise they stream through their own mbarrier ring, decoupled from the halo ring.
template <int BLOCK_N, bool RESIDENT_B>
struct HaloCfg {
  static constexpr int HALO_W_MAX = 144;                           // 128 + 2*8
  static constexpr int A_STAGE_BYTES = 3 * HALO_W_MAX * 128;       // 55296, multiple of 1024
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int B_SLOTS = RESIDENT_B ? 9 : (BLOCK_N >= 256 ? 3 : (BLOCK_N >= 64 ? 6 : 9));
  // three halo stages wherever shared memory allows (the loads are latency bound: bytes in flight matter)
  static constexpr int A_STAGES = (3 * A_STAGE_BYTES + B_SLOTS * B_BYTES + 1024 + 512 + 8 * BLOCK_N * 4 <= 232448) ? 3 : 2;
  static constexpr int B_TOTAL = B_SLOTS * B_BYTES;
  static constexpr int DATA_BYTES = A_STAGES * A_STAGE_BYTES + B_TOTAL;
  static constexpr int SMEM_BYTES = DATA_BYTES + 1024 + 512 + 8 * BLOCK_N * 4;
  static_assert(SMEM_BYTES <= 232448, "halo kernel exceeds shared memory");
};

template <int BLOCK_N, bool RESIDENT_B>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_rowhalo_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b) {
  using C = ConvCfg<BLOCK_N>;
  using H = HaloCfg<BLOCK_N, RESIDENT_B>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *a_buf = smem;
  unsigned char *b_buf = smem + H::A_STAGES * H::A_STAGE_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + H::DATA_BYTES);
  uint64_t *a_full = bars, *a_empty = bars + 3, *b_full = bars + 6, *b_empty = bars + 15;
  uint64_t *tmem_full = bars + 24, *tmem_empty = bars + 26;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 28);
  float *epi_vec = reinterpret_cast<float *>(smem + H::DATA_BYTES + 512);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    for (int i = 0; i < 9; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int d = p.halo_d, hw = p.halo_w;
  const uint32_t a_row_bytes = (uint32_t)hw * 128u;

  if (warp == 0) {
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0, res_ph = 0;
    long long cur_key = -1;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long t = tile;
      const int n_i = tile_n_index(p, tile); t /= p.tiles_n;
      const int w_i = (int)(t % p.tiles_w); t /= p.tiles_w;
      const int oh = (int)(t % p.tiles_h); t /= p.tiles_h;
      const int b = (int)t;
      const int g = p.groups == 1 ? 0 : b;
      if (RESIDENT_B) {
        const long long key = (long long)g * p.tiles_n + n_i;
        if (key != cur_key) {
          mbar_wait(&b_empty[0], res_ph ^ 1);          // every MMA that read the old weights has retired
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[0], 9 * H::B_BYTES);
            for (int tap = 0; tap < 9; ++tap)
              tma_load_4d(b_buf + tap * H::B_BYTES, &tmap_b, &b_full[0], 0, n_i * BLOCK_N, p.tap_w[tap], g);
          }
          __syncwarp();
          cur_key = key;
          res_ph ^= 1;
        }
      }
      for (int c = 0; c < p.kc; ++c) {
        mbar_wait(&a_empty[as], aph ^ 1);
        if (elect_one()) {
          unsigned char *sa = a_buf + as * H::A_STAGE_BYTES;
          mbar_arrive_expect_tx(&a_full[as], 3 * a_row_bytes);
          for (int kh = 0; kh < 3; ++kh)
            tma_load_4d(sa + kh * a_row_bytes, &tmap_a, &a_full[as], c * kBlockK, w_i * kBlockM - d, oh + (kh - 1) * d, b);
        }
        __syncwarp();
        if (++as == H::A_STAGES) { as = 0; aph ^= 1; }
        if (!RESIDENT_B) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&b_full[bs], H::B_BYTES);
              tma_load_4d(b_buf + bs * H::B_BYTES, &tmap_b, &b_full[bs], c * kBlockK, n_i * BLOCK_N, p.tap_w[tap], g);
            }
            __syncwarp();
            if (++bs == H::B_SLOTS) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, acc_phase = 0, res_ph = 0;
    long long cur_key = -1;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      if (RESIDENT_B) {
        const int n_i = (int)(tile % p.tiles_n);
        const int b = (int)(tile / ((long long)p.tiles_n * p.tiles_w * p.tiles_h));
        const long long key = (long long)(p.groups == 1 ? 0 : b) * p.tiles_n + n_i;
        if (key != cur_key) {
          mbar_wait(&b_full[0], res_ph);
          res_ph ^= 1;
          cur_key = key;
        }
      }
      bool release = false;
      if (RESIDENT_B) {
        const long long nt = tile + gridDim.x;
        release = nt >= p.total_tiles;
        if (!release) {
          const int n_i2 = (int)(nt % p.tiles_n);
          const int b2 = (int)(nt / ((long long)p.tiles_n * p.tiles_w * p.tiles_h));
          release = ((long long)(p.groups == 1 ? 0 : b2) * p.tiles_n + n_i2) != cur_key;
        }
      }
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int c = 0; c < p.kc; ++c) {
        mbar_wait(&a_full[as], aph);
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(a_buf + as * H::A_STAGE_BYTES);
        if (RESIDENT_B) {
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int kh = tap / 3, kw = tap - kh * 3;
              const uint64_t adesc = umma_smem_desc(a_base + (uint32_t)kh * a_row_bytes + (uint32_t)(kw * d) * 128u, 128);
              const uint64_t bdesc = umma_smem_desc(smem_u32(b_buf + tap * H::B_BYTES), 128);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                             (c > 0 || tap > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&a_empty[as]);
            if (c == p.kc - 1) {
              umma_commit(&tmem_full[acc]);
              if (release) umma_commit(&b_empty[0]);   // weights may be overwritten once these MMAs retire
            }
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int kh = tap / 3, kw = tap - kh * 3;
            mbar_wait(&b_full[bs], bph);
            tcgen05_fence_after();
            if (elect_one()) {
              const uint64_t adesc = umma_smem_desc(a_base + (uint32_t)kh * a_row_bytes + (uint32_t)(kw * d) * 128u, 128);
              const uint64_t bdesc = umma_smem_desc(smem_u32(b_buf + bs * H::B_BYTES), 128);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                             (c > 0 || tap > 0 || k > 0) ? 1u : 0u);
              umma_commit(&b_empty[bs]);
              if (tap == 8) {
                umma_commit(&a_empty[as]);
                if (c == p.kc - 1) umma_commit(&tmem_full[acc]);
              }
            }
            __syncwarp();
            if (++bs == H::B_SLOTS) { bs = 0; bph ^= 1; }
          }
        }
        if (++as == H::A_STAGES) { as = 0; aph ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    epilogue_role<BLOCK_N>(p, epi_vec, tmem_full, tmem_empty, tmem_base, warp, lane);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BLOCK_N, bool RESIDENT_B>
int launch_halo(ConvParams &p, const void *x, int64_t batch, int64_t in_h, int64_t in_w, const void *wq,
                int64_t cout_pad, int taps_total, cudaStream_t stream) {
  using H = HaloCfg<BLOCK_N, RESIDENT_B>;
  auto kern = conv_rowhalo_kernel<BLOCK_N, RESIDENT_B>;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, H::SMEM_BYTES));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)batch};
    uint64_t strides[4] = {0, (uint64_t)p.cin * 2, (uint64_t)p.cin * in_w * 2, (uint64_t)p.cin * in_w * in_h * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.halo_w, 1, 1};
    if (int rc = encode_tma(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)p.cout, (uint64_t)taps_total, (uint64_t)p.groups};
    uint64_t strides[4] = {0, (uint64_t)p.cin * 2, (uint64_t)p.cin * cout_pad * 2,
                           (uint64_t)p.cin * cout_pad * taps_total * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)BLOCK_N, 1, 1};
    if (int rc = encode_tma(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, wq, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  p.tiles_n = (p.cout + BLOCK_N - 1) / BLOCK_N;
  p.total_tiles = (long long)p.batch * p.tiles_h * p.tiles_w * p.tiles_n;
  long long grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kern<<<(unsigned)grid, kNumThreads, H::SMEM_BYTES, stream>>>(p, ta, tb);
  return check_launch("conv_rowhalo_kernel");
}

template <int BLOCK_N, bool STAGED>
int launch_conv_impl(ConvParams &p, const CUtensorMap &ta, const void *wq, int64_t cout_pad, int taps_total,
                     cudaStream_t stream) {
  using C = ConvCfg<BLOCK_N>;
  auto kern = conv_fprop_kernel<BLOCK_N, STAGED>;
  constexpr int SMEM = STAGED ? C::SMEM_BYTES_STAGED : C::SMEM_BYTES;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  CUtensorMap tb;
  {
    // visible rows = cout (a channel slice of a wider packed tensor is legal); strides use cout_pad
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)(p.cls_mode ? p.shuffle_cout : p.cout), (uint64_t)taps_total, (uint64_t)p.groups};
    uint64_t strides[4] = {0, (uint64_t)p.cin * 2, (uint64_t)p.cin * cout_pad * 2,
                           (uint64_t)p.cin * cout_pad * taps_total * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)BLOCK_N, 1, 1};
    if (int rc = encode_tma(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, wq, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  OutMaps om;
  memset(&om, 0, sizeof(om));
  if (STAGED) {
    // one map per output parity class: base shifted to the class origin, pixel strides scaled by `os`
    const int ncls = p.shuffle_cout ? 4 : 1;
    const int os = p.os;
    const int creal = p.shuffle_cout ? p.shuffle_cout : p.cout;
    for (int cls = 0; cls < ncls; ++cls) {
      const int oh0 = p.oo_h + (cls >> 1), ow0 = p.oo_w + (cls & 1);
      const __nv_bfloat16 *base = static_cast<const __nv_bfloat16 *>(p.out) + ((long long)oh0 * p.full_w + ow0) * p.ldo;
      uint64_t dims[4] = {(uint64_t)(p.co_off + creal), (uint64_t)((p.full_w - ow0 + os - 1) / os),
                          (uint64_t)((p.full_h - oh0 + os - 1) / os), (uint64_t)p.batch};
      uint64_t strides[4] = {0, (uint64_t)p.ldo * 2 * os, (uint64_t)p.ldo * p.full_w * 2 * os,
                             (uint64_t)p.ldo * p.full_w * p.full_h * 2};
      uint32_t box[4] = {(uint32_t)C::CHUNK, 32, 1, 1};
      if (int rc = encode_tma(&om.m[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, nullptr,
                              C::CHUNK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B))
        return rc;
    }
  }
  p.tiles_n = (p.cout + BLOCK_N - 1) / BLOCK_N;
  p.total_tiles = (long long)p.tiles_b * p.tiles_h * p.tiles_w * p.tiles_n;
  long long grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kern<<<(unsigned)grid, kNumThreads, SMEM, stream>>>(p, ta, tb, om);
  return check_launch("conv_fprop_kernel");
}

// CTA-pair launch (N = 256, staged epilogue, one sample per tile).  Returns -1 when the device cannot co-schedule a pair.
template <int BLOCK_N>
int launch_conv_pair(ConvParams &p, const CUtensorMap &ta, const void *wq, int64_t cout_pad, int taps_total,
                     cudaStream_t stream) {
  using P = PairCfg<BLOCK_N>;
  auto kern = conv_fprop_pair_kernel<BLOCK_N>;
  static int state[64] = {0};            // 0 = unknown, 1 = usable, -1 = not
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return -1;
  if (state[dev] == 0) {
    state[dev] = -1;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(2, 1, 1);
      cfg.blockDim = dim3(kNumThreads, 1, 1);
      cfg.dynamicSmemBytes = P::SMEM_BYTES;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters >= 1) state[dev] = nclusters;
    }
    cudaGetLastError();
  }
  if (state[dev] < 1) return -1;
  CUtensorMap tb;
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)(p.cls_mode ? p.shuffle_cout : p.cout), (uint64_t)taps_total, (uint64_t)p.groups};
    uint64_t strides[4] = {0, (uint64_t)p.cin * 2, (uint64_t)p.cin * cout_pad * 2,
                           (uint64_t)p.cin * cout_pad * taps_total * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(BLOCK_N / 2), 1, 1};
    if (int rc = encode_tma(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, wq, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  OutMaps om;
  memset(&om, 0, sizeof(om));
  {
    const int ncls = p.shuffle_cout ? 4 : 1;
    const int os = p.os;
    const int creal = p.shuffle_cout ? p.shuffle_cout : p.cout;
    for (int cls = 0; cls < ncls; ++cls) {
      const int oh0 = p.oo_h + (cls >> 1), ow0 = p.oo_w + (cls & 1);
      const __nv_bfloat16 *base = static_cast<const __nv_bfloat16 *>(p.out) + ((long long)oh0 * p.full_w + ow0) * p.ldo;
      uint64_t dims[4] = {(uint64_t)(p.co_off + creal), (uint64_t)((p.full_w - ow0 + os - 1) / os),
                          (uint64_t)((p.full_h - oh0 + os - 1) / os), (uint64_t)p.batch};
      uint64_t strides[4] = {0, (uint64_t)p.ldo * 2 * os, (uint64_t)p.ldo * p.full_w * 2 * os,
                             (uint64_t)p.ldo * p.full_w * p.full_h * 2};
      uint32_t box[4] = {(uint32_t)ConvCfg<BLOCK_N>::CHUNK, 32, 1, 1};
      if (int rc = encode_tma(&om.m[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, nullptr,
                              CU_TENSOR_MAP_SWIZZLE_64B))
        return rc;
    }
  }
  p.tiles_n = (p.cout + BLOCK_N - 1) / BLOCK_N;
  p.mpairs = (p.tiles_h * p.tiles_w + 1) / 2;
  p.total_pairs = (long long)p.batch * p.mpairs * p.tiles_n;
  long long pairs = p.total_pairs < state[dev] ? p.total_pairs : state[dev];
  kern<<<(unsigned)(2 * pairs), kNumThreads, P::SMEM_BYTES, stream>>>(p, ta, tb, om);
  return check_launch("conv_fprop_pair_kernel");
}

template <int BLOCK_N>
int launch_conv(ConvParams &p, const CUtensorMap &ta, const void *wq, int64_t cout_pad, int taps_total,
                cudaStream_t stream) {
  return p.staged ? launch_conv_impl<BLOCK_N, true>(p, ta, wq, cout_pad, taps_total, stream)
                  : launch_conv_impl<BLOCK_N, false>(p, ta, wq, cout_pad, taps_total, stream);
}

inline int next_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace

// class-mode tap table (ConvParams::cls_*): per output parity class, its taps as (shift index, weight tap)
struct ClassTaps {
  int ntaps[4], shift[4][4], w[4][4];
};

// Shared by the public entry points below and by the transposed-conv helper.
int conv_gather_launch(const void *x, const void *wq, int64_t batch, int64_t groups, int64_t in_h,
                       int64_t in_w, int64_t cin, int64_t cout, int64_t cout_pad, int taps_total,
                       int ntaps, const int *tap_w, const int *tap_dy, const int *tap_dx, int stride,
                       int64_t out_h, int64_t out_w, void *out, int out_nhwc, int64_t full_h,
                       int64_t full_w, int os, int oo_h, int oo_w, int64_t ldo, int64_t co_off,
                       const vsp_conv_epilogue *epi, cudaStream_t stream, int shuffle_cout = 0, int n_branches = 0,
                       const int *branch_dils = nullptr, const ClassTaps *cls = nullptr) {
  VSP_REQUIRE(batch >= 1 && (groups == 1 || groups == batch), "conv: groups must be 1 or batch");
  VSP_REQUIRE(cin >= 8 && cin % 8 == 0, "conv: cin must be a multiple of 8 (pad NHWC channels), got %lld", (long long)cin);
  VSP_REQUIRE(cout >= 1 && cout_pad >= (cls != nullptr ? shuffle_cout : cout), "conv: bad cout");
  VSP_REQUIRE(ntaps >= 1 && ntaps <= kMaxTaps, "conv: 1..16 taps supported");
  VSP_REQUIRE(stride == 1 || stride == 2, "conv: stride must be 1 or 2");
  VSP_REQUIRE(x && wq && out, "conv: null pointer");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(wq) & 15) == 0,
              "conv: operands must be 16-byte aligned");
  VSP_REQUIRE(out_h >= 1 && out_w >= 1 && in_h >= 1 && in_w >= 1, "conv: empty extent");
  VSP_REQUIRE(batch < 65536 && in_h < 65536 && in_w < 65536 && full_h < 65536 && full_w < 65536, "conv: extent too large");

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.batch = (int)batch; p.groups = (int)groups; p.cin = (int)cin; p.cout = (int)cout;
  p.out_h = (int)out_h; p.out_w = (int)out_w; p.stride = stride; p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    VSP_REQUIRE(tap_w[t] >= 0 && tap_w[t] < taps_total, "conv: tap index out of range");
    p.tap_w[t] = tap_w[t]; p.tap_dy[t] = tap_dy[t]; p.tap_dx[t] = tap_dx[t];
  }
  p.tw = next_pow2((int)out_w) < kBlockM ? next_pow2((int)out_w) : kBlockM;
  p.th = kBlockM / p.tw;
  p.tb = 1;
  if (groups == 1 && next_pow2((int)out_h) < p.th) {
    // shared weights and an image smaller than one tile: stack several samples into the 128 rows
    p.th = next_pow2((int)out_h);
    p.tb = kBlockM / (p.tw * p.th);
  }
  p.tiles_b = ((int)batch + p.tb - 1) / p.tb;
  p.tiles_w = ((int)out_w + p.tw - 1) / p.tw;
  p.tiles_h = ((int)out_h + p.th - 1) / p.th;
  if (n_branches > 0) {
    VSP_REQUIRE(n_branches <= 4 && branch_dils != nullptr && cout % n_branches == 0, "conv: bad branch description");
    const int cq = (int)cout / n_branches;
    VSP_REQUIRE(cq >= 16 && cq <= 256 && (cq & (cq - 1)) == 0, "conv: branch width must be a power of two in 16..256");
    p.branch_mode = 1;
    for (int j = 0; j < n_branches; ++j) p.n_dil[j] = branch_dils[j];
  }
  p.kc = ((int)cin + kBlockK - 1) / kBlockK;
  p.out = out; p.out_nhwc = out_nhwc;
  p.shuffle_cout = shuffle_cout;
  if (cls != nullptr) {
    VSP_REQUIRE(shuffle_cout > 0 && cout == 4 * shuffle_cout && shuffle_cout % 128 == 0 && n_branches == 0,
                "conv: class mode needs the pixel-shuffle mapping with Cout %% 128 == 0");
    p.cls_mode = 1;
    for (int c = 0; c < 4; ++c) {
      p.cls_ntaps[c] = cls->ntaps[c];
      for (int j = 0; j < 4; ++j) { p.cls_shift[c][j] = cls->shift[c][j]; p.cls_w[c][j] = cls->w[c][j]; }
    }
  }
  p.full_h = (int)full_h; p.full_w = (int)full_w; p.os = os; p.oo_h = oo_h; p.oo_w = oo_w;
  p.ldo = ldo; p.co_off = co_off;
  if (epi) {
    p.row_scale = epi->row_scale; p.noise_weight = epi->noise_weight; p.bias = epi->bias;
    p.act = epi->act; p.alpha = epi->alpha; p.scale = epi->scale; p.residual = epi->residual;
    p.noise = epi->noise; p.noise_bstride = epi->noise_bstride; p.residual2 = epi->residual2;
    p.noise_weight_dev = epi->noise_weight_dev; p.pre_bias = epi->pre_bias; p.pre_act = epi->pre_act;
    p.alpha_vec = epi->act != 0 ? epi->alpha_vec : nullptr;
    VSP_REQUIRE(p.alpha_vec == nullptr || p.pre_act == 0, "conv: a per-channel slope needs a single activation stage");
    VSP_REQUIRE(p.pre_act == 0 || p.pre_act == 3, "conv: epilogue pre_act must be 0 or 3");
    VSP_REQUIRE(p.act == 0 || p.act == 3, "conv: epilogue act must be 0 or 3");
  }

  // staged epilogue (swizzled shared memory + TMA stores) for NHWC bf16 outputs whose warps cover one output row
  {
    static const bool no_staged = getenv("VSP_NO_STAGED") != nullptr;
    const float al = p.alpha, sc = p.scale;
    p.staged = (!no_staged && out_nhwc && (ldo % 8) == 0 && (co_off % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                p.tw >= 32 && p.tb == 1 && ((al >= 0.f && al <= 1.f) || p.alpha_vec != nullptr) &&
                (sc > 0.f || (p.act == 0 && p.pre_act == 0))) ? 1 : 0;
    if (shuffle_cout) {
      VSP_REQUIRE(os == 2 && oo_h == 0 && oo_w == 0 && shuffle_cout % 32 == 0 && cout == 4 * shuffle_cout,
                  "conv: pixel-shuffle epilogue needs Cout %% 32 == 0");
    }
  }

  // Row-ring path (conv_ring_sm100.cu): wide, shallow stride-1 3x3 (dilated) / 1x1 layers (scalar slopes only)
  if (p.cls_mode) VSP_REQUIRE(p.staged, "conv: class mode needs the staged (NHWC bf16, TMA store) epilogue");
  if (p.alpha_vec == nullptr && !p.cls_mode) {
    const int rc = conv_ring_try_launch(p, x, wq, in_h, in_w, cout_pad, taps_total, ntaps == 9 ? tap_dx[8] : 1, stream);
    if (rc >= 0) return rc;
  }

  // Row-halo path: stride-1 3x3 (dilated) "same" convolution on a wide image, plain output mapping
  {
    static const bool no_halo = getenv("VSP_NO_HALO") != nullptr;
    // (Cout > 128 goes to the generic kernel: as a CTA pair it runs 256->256 @128^2 in 385 us vs 468 us here)
    static const bool halo_wide = getenv("VSP_HALO_WIDE_N") != nullptr;
    bool grid3 = !no_halo && p.alpha_vec == nullptr && !p.branch_mode && stride == 1 && ntaps == 9 && os == 1 && out_w >= kBlockM && out_w == in_w && out_h == in_h &&
                 (cout <= 128 || halo_wide || !p.staged);
    int dd = grid3 ? tap_dx[8] : 0;   // tap (kh,kw) offset must be ((kh-1)*d, (kw-1)*d)
    grid3 = grid3 && dd >= 1 && dd <= 8;
    for (int t = 0; grid3 && t < 9; ++t)
      grid3 = tap_dy[t] == (t / 3 - 1) * dd && tap_dx[t] == (t % 3 - 1) * dd;
    if (grid3) {
      p.tw = kBlockM; p.th = 1;
      p.tiles_w = ((int)out_w + kBlockM - 1) / kBlockM;
      p.tiles_h = (int)out_h;
      p.halo_d = dd;
      p.halo_w = (kBlockM + 2 * dd + 7) & ~7;
      const bool resident = p.kc == 1;
      if (cout > 128) return launch_halo<256, false>(p, x, batch, in_h, in_w, wq, cout_pad, taps_total, stream);
      if (cout > 64) return launch_halo<128, false>(p, x, batch, in_h, in_w, wq, cout_pad, taps_total, stream);
      if (cout > 32) return resident ? launch_halo<64, true>(p, x, batch, in_h, in_w, wq, cout_pad, taps_total, stream)
                                     : launch_halo<64, false>(p, x, batch, in_h, in_w, wq, cout_pad, taps_total, stream);
      if (cout > 16) return resident ? launch_halo<32, true>(p, x, batch, in_h, in_w, wq, cout_pad, taps_total, stream)
                                     : launch_halo<32, false>(p, x, batch, in_h, in_w, wq, cout_pad, taps_total, stream);
      return resident ? launch_halo<16, true>(p, x, batch, in_h, in_w, wq, cout_pad, taps_total, stream)
                      : launch_halo<16, false>(p, x, batch, in_h, in_w, wq, cout_pad, taps_total, stream);
    }
  }

  // A operand: NHWC bf16 activations as a 4-D tensor (c, w, h, b); the box is one tap-shifted
  // patch of th x tw pixels x 64 channels; stride-2 convs use the TMA traversal stride.
  CUtensorMap ta;
  {
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)batch};
    uint64_t strides[4] = {0, (uint64_t)cin * 2, (uint64_t)cin * in_w * 2, (uint64_t)cin * in_w * in_h * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(p.tw * stride), (uint32_t)(p.th * stride), (uint32_t)p.tb};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (int rc = encode_tma(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  // BLOCK_N: widest tile that still fills the machine. Cost model per k-block ~ (256 + 2*BLOCK_N) cycles
  // (measured: the mainloop is bound by ~64 B/clk/SM of L2->smem traffic: 16 KB of A + 128*BLOCK_N B of B),
  // times k-blocks, times waves over the SMs.
  int best_bn = 16;
  {
    double best = 1e300;
    const long long m_tiles = (long long)p.tiles_b * p.tiles_h * p.tiles_w;
    for (int bn = 256; bn >= 16; bn >>= 1) {
      if (p.branch_mode && bn != (int)cout / n_branches) continue;   // one channel tile per branch
      if (p.cls_mode && bn != (shuffle_cout % 256 == 0 ? 256 : 128)) continue;   // whole channel tiles per parity class
      if (bn > 16 && bn >= 2 * cout) continue;               // more than half of the tile would be padding
      const long long tiles = m_tiles * ((cout + bn - 1) / bn);
      const long long waves = (tiles + num_sms() - 1) / num_sms();
      const double cost = (double)waves * ((double)p.ntaps * p.kc * (256.0 + 2.0 * bn) + 600.0 + 8.0 * bn);
      if (cost < best) { best = cost; best_bn = bn; }
    }
  }
  if (p.cls_mode) p.cls_tpc = shuffle_cout / best_bn;
  if ((best_bn == 256 || best_bn == 128) && p.staged && p.tb == 1) {
    // CTA pairs (tcgen05 cta_group::2) for the N = 256 / 128 layers; VSP_CONV_PAIR=0 keeps the single-CTA kernel
    static const int pair_on = getenv("VSP_CONV_PAIR") == nullptr ? 3 : atoi(getenv("VSP_CONV_PAIR"));   // bit 0: N = 256, bit 1: N = 128
    if (best_bn == 256 && (pair_on & 1)) {
      const int rc = launch_conv_pair<256>(p, ta, wq, cout_pad, taps_total, stream);
      if (rc >= 0) return rc;
    }
    if (best_bn == 128 && (pair_on & 2)) {
      const int rc = launch_conv_pair<128>(p, ta, wq, cout_pad, taps_total, stream);
      if (rc >= 0) return rc;
    }
  }
  switch (best_bn) {
    case 256: return launch_conv<256>(p, ta, wq, cout_pad, taps_total, stream);
    case 128: return launch_conv<128>(p, ta, wq, cout_pad, taps_total, stream);
    case 64: return launch_conv<64>(p, ta, wq, cout_pad, taps_total, stream);
    case 32: return launch_conv<32>(p, ta, wq, cout_pad, taps_total, stream);
    default: return launch_conv<16>(p, ta, wq, cout_pad, taps_total, stream);
  }
}

}  // namespace vsp

extern "C" int vsp_conv2d_fprop_bf16(const void *x, const void *wq, void *out, int64_t batch,
                                     int64_t groups, int64_t in_h, int64_t in_w, int64_t cin,
                                     int64_t cout, int64_t cout_pad, int kh, int kw, int stride,
                                     int pad, int dil, int out_nhwc_bf16, int64_t ldo, int64_t co_off,
                                     const vsp_conv_epilogue *epi, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= kMaxTaps, "conv2d_fprop: kernel up to 16 taps");
  VSP_REQUIRE(dil >= 1 && pad >= 0, "conv2d_fprop: bad dilation/padding");
  const int64_t out_h = (in_h + 2 * pad - dil * (kh - 1) - 1) / stride + 1;
  const int64_t out_w = (in_w + 2 * pad - dil * (kw - 1) - 1) / stride + 1;
  VSP_REQUIRE(out_h >= 1 && out_w >= 1, "conv2d_fprop: empty output");
  int tw_[kMaxTaps], dy[kMaxTaps], dx[kMaxTaps];
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) {
      tw_[i * kw + j] = i * kw + j;
      dy[i * kw + j] = i * dil - pad;
      dx[i * kw + j] = j * dil - pad;
    }
  if (!out_nhwc_bf16) { ldo = cout; co_off = 0; }
  return conv_gather_launch(x, wq, batch, groups, in_h, in_w, cin, cout, cout_pad, kh * kw, kh * kw, tw_, dy, dx,
                            stride, out_h, out_w, out, out_nhwc_bf16, out_h, out_w, 1, 0, 0, ldo, co_off, epi, stream);
}

extern "C" int vsp_conv2d_gather_bf16(const void *x, const void *wq, void *out, int64_t batch, int64_t groups,
                                      int64_t in_h, int64_t in_w, int64_t cin, int64_t cout, int64_t cout_pad,
                                      int taps_total, int ntaps, const int *tap_w, const int *tap_dy,
                                      const int *tap_dx, int stride, int64_t out_h, int64_t out_w,
                                      int out_nhwc_bf16, int64_t full_h, int64_t full_w, int os, int oo_h,
                                      int oo_w, int64_t ldo, int64_t co_off, const vsp_conv_epilogue *epi,
                                      void *stream_) {
  using namespace vsp;
  VSP_REQUIRE(tap_w && tap_dy && tap_dx, "conv2d_gather: null tap list");
  VSP_REQUIRE(os >= 1 && oo_h >= 0 && oo_w >= 0, "conv2d_gather: bad output mapping");
  VSP_REQUIRE((out_h - 1) * os + oo_h < full_h && (out_w - 1) * os + oo_w < full_w,
              "conv2d_gather: output mapping exceeds the full extent");
  if (!out_nhwc_bf16) { ldo = cout; co_off = 0; }
  return conv_gather_launch(x, wq, batch, groups, in_h, in_w, cin, cout, cout_pad, taps_total, ntaps, tap_w, tap_dy,
                            tap_dx, stride, out_h, out_w, out, out_nhwc_bf16, full_h, full_w, os, oo_h, oo_w, ldo,
                            co_off, epi, static_cast<cudaStream_t>(stream_));
}

extern "C" int vsp_conv_transpose2d_s2_bf16(const void *x, const void *wq, void *out, int64_t batch,
                                            int64_t groups, int64_t in_h, int64_t in_w, int64_t cin,
                                            int64_t cout, int64_t cout_pad, int kh, int kw, int out_nhwc_bf16,
                                            int64_t ldo, int64_t co_off, const vsp_conv_epilogue *epi,
                                            void *stream_) {
  using namespace vsp;
  VSP_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= kMaxTaps, "conv_transpose2d_s2: kernel up to 16 taps");
  const int64_t full_h = (in_h - 1) * 2 + kh, full_w = (in_w - 1) * 2 + kw;
  if (!out_nhwc_bf16) { ldo = cout; co_off = 0; }
  // out[2a+pa, 2b+pb] = sum_{kh_i = pa (mod 2), kw_i = pb (mod 2)} x[a - (kh_i-pa)/2, b - (kw_i-pb)/2] * w[kh_i, kw_i]
  //
  // 3x3, NHWC bf16 output, Cout % 128 == 0: the 2H x 2W block of the output as ONE launch in class mode.  The four parity
  // classes are the four column groups of a pixel-shuffle GEMM over the H x W low-resolution grid; each channel tile runs
  // only its class's 4 / 2 / 2 / 1 taps straight from the layer's packed weights.  As four launches every class re-read x
  // and was bound by its own ramp, tail and epilogue (256->128 @128^2 x 32: 670 us = 461 TFLOP/s).  The last output row and
  // column (index 2H, 2W: the even classes have one more row / column than the grid) are four thin gather launches.
  // VSP_TCONV_CLASSES=0 keeps the four-launch form.
  {
    static const bool one_launch = getenv("VSP_TCONV_CLASSES") == nullptr || atoi(getenv("VSP_TCONV_CLASSES")) != 0;
    const bool epi_ok = epi == nullptr || (epi->noise == nullptr && epi->residual == nullptr && epi->residual2 == nullptr &&
                                           epi->alpha_vec == nullptr && epi->alpha >= 0.f && epi->alpha <= 1.f &&
                                           (epi->scale > 0.f || (epi->act == 0 && epi->pre_act == 0)));
    const int tw = next_pow2((int)in_w) < kBlockM ? next_pow2((int)in_w) : kBlockM;
    const bool stacked = groups == 1 && next_pow2((int)in_h) < kBlockM / tw;
    if (one_launch && kh == 3 && kw == 3 && out_nhwc_bf16 && cout % 128 == 0 && epi_ok && tw >= 32 && !stacked &&
        (ldo % 8) == 0 && (co_off % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && getenv("VSP_NO_STAGED") == nullptr) {
      cudaStream_t st = static_cast<cudaStream_t>(stream_);
      ClassTaps ct;
      memset(&ct, 0, sizeof(ct));
      for (int pa = 0; pa < 2; ++pa)
        for (int pb = 0; pb < 2; ++pb) {
          const int c = pa * 2 + pb;
          for (int i = pa; i < 3; i += 2)
            for (int j = pb; j < 3; j += 2) {
              const int n = ct.ntaps[c]++;
              ct.shift[c][n] = ((i - pa) / 2) * 2 + (j - pb) / 2;     // shift index: bit 1 = row shift -1, bit 0 = column shift -1
              ct.w[c][n] = i * 3 + j;
            }
        }
      const int tw4[4] = {0, 0, 0, 0}, dy4[4] = {0, 0, -1, -1}, dx4[4] = {0, -1, 0, -1};
      if (int rc = conv_gather_launch(x, wq, batch, groups, in_h, in_w, cin, 4 * cout, cout_pad, 9, 4, tw4, dy4, dx4, 1, in_h,
                                      in_w, out, 1, full_h, full_w, 2, 0, 0, ldo, co_off, epi, st, (int)cout, 0, nullptr, &ct))
        return rc;
      for (int pb = 0; pb < 2; ++pb) {          // output row 2H (class row parity 0, a = H): columns of parity pb
        int tw_[4], dy[4], dx[4], nt = 0;
        for (int i = 0; i < 3; i += 2)
          for (int j = pb; j < 3; j += 2) { tw_[nt] = i * 3 + j; dy[nt] = (int)in_h - i / 2; dx[nt] = -(j - pb) / 2; ++nt; }
        if (int rc = conv_gather_launch(x, wq, batch, groups, in_h, in_w, cin, cout, cout_pad, 9, nt, tw_, dy, dx, 1, 1,
                                        (full_w - pb + 1) / 2, out, 1, full_h, full_w, 2, 2 * (int)in_h, pb, ldo, co_off, epi, st))
          return rc;
      }
      for (int pa = 0; pa < 2; ++pa) {          // output column 2W (column parity 0, b = W): rows 2a + pa, a < H
        int tw_[4], dy[4], dx[4], nt = 0;
        for (int i = pa; i < 3; i += 2)
          for (int j = 0; j < 3; j += 2) { tw_[nt] = i * 3 + j; dy[nt] = -(i - pa) / 2; dx[nt] = (int)in_w - j / 2; ++nt; }
        if (int rc = conv_gather_launch(x, wq, batch, groups, in_h, in_w, cin, cout, cout_pad, 9, nt, tw_, dy, dx, 1, in_h, 1,
                                        out, 1, full_h, full_w, 2, pa, 2 * (int)in_w, ldo, co_off, epi, st))
          return rc;
      }
      return 0;
    }
  }
  for (int pa = 0; pa < 2; ++pa)
    for (int pb = 0; pb < 2; ++pb) {
      int tw_[kMaxTaps], dy[kMaxTaps], dx[kMaxTaps], nt = 0;
      for (int i = pa; i < kh; i += 2)
        for (int j = pb; j < kw; j += 2) {
          tw_[nt] = i * kw + j;
          dy[nt] = -(i - pa) / 2;
          dx[nt] = -(j - pb) / 2;
          ++nt;
        }
      const int64_t oh = (full_h - pa + 1) / 2, ow = (full_w - pb + 1) / 2;
      if (nt == 0 || oh <= 0 || ow <= 0) continue;  // (a class with no taps stays zero: caller pre-zeroes if kh or kw == 1)
      if (int rc = conv_gather_launch(x, wq, batch, groups, in_h, in_w, cin, cout, cout_pad, kh * kw, nt, tw_, dy, dx, 1,
                                      oh, ow, out, out_nhwc_bf16, full_h, full_w, 2, pa, pb, ldo, co_off, epi,
                                      static_cast<cudaStream_t>(stream_)))
        return rc;
    }
  return 0;
}

extern "C" int vsp_conv2d_up2_fused_bf16(const void *x, const void *wq, void *out, int64_t batch, int64_t groups,
                                         int64_t in_h, int64_t in_w, int64_t cin, int64_t cout, int64_t ldo,
                                         int64_t co_off, const vsp_conv_epilogue *epi, void *stream_) {
  using namespace vsp;
  VSP_REQUIRE(cout >= 32 && cout % 32 == 0, "conv2d_up2_fused: Cout must be a multiple of 32");
  int tw_[9], dy[9], dx[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      tw_[i * 3 + j] = i * 3 + j;
      dy[i * 3 + j] = i - 1;
      dx[i * 3 + j] = j - 1;
    }
  return conv_gather_launch(x, wq, batch, groups, in_h, in_w, cin, 4 * cout, 4 * cout, 9, 9, tw_, dy, dx, 1, in_h, in_w,
                            out, 1, 2 * in_h, 2 * in_w, 2, 0, 0, ldo, co_off, epi, static_cast<cudaStream_t>(stream_),
                            (int)cout);
}

extern "C" int vsp_conv2d_branches_bf16(const void *x, const void *wq, void *out, int64_t batch, int64_t groups,
                                        int64_t in_h, int64_t in_w, int64_t cin, int64_t cout, int n_branches,
                                        const int *dils, int out_nhwc_bf16, int64_t ldo, int64_t co_off,
                                        const vsp_conv_epilogue *epi, void *stream_) {
  using namespace vsp;
  VSP_REQUIRE(n_branches >= 1 && n_branches <= 4 && dils != nullptr, "conv2d_branches: 1..4 branches");
  int tw_[9], dy[9], dx[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      tw_[i * 3 + j] = i * 3 + j;
      dy[i * 3 + j] = i - 1;
      dx[i * 3 + j] = j - 1;
    }
  for (int j = 0; j < n_branches; ++j) VSP_REQUIRE(dils[j] >= 1 && dils[j] <= 64, "conv2d_branches: bad dilation");
  if (!out_nhwc_bf16) { ldo = cout; co_off = 0; }
  return conv_gather_launch(x, wq, batch, groups, in_h, in_w, cin, cout, cout, 9, 9, tw_, dy, dx, 1, in_h, in_w, out,
                            out_nhwc_bf16, in_h, in_w, 1, 0, 0, ldo, co_off, epi, static_cast<cudaStream_t>(stream_), 0,
                            n_branches, dils);
}
