// conv_ring_sm100.cu — "row-ring" tcgen05 convolution for the wide, shallow layers (Cin <= 64, Cout <= 64,
// width >= 128: the 512^2 and 1024^2 levels of both networks, SURVEY.md Appendix A "ridge point").
//
// Those layers are bandwidth- and latency-bound, not tensor-bound: with the plain implicit GEMM
// (conv_sm100.cu) every 128-pixel tile re-reads its input once per tap, pays a full epilogue round trip
// for 128 x N outputs and scatters 16-byte stores over 128-byte pixel pitches.  Here a CTA owns a vertical
// strip (128 output columns) of one sample and walks down it:
//  * every INPUT ROW (128 + 2*dil pixels x 64 channels, one TMA box) is loaded exactly once into a ring of
//    row slots and serves the three output rows that need it (the nine taps are row-/column-shifted UMMA
//    descriptors into the ring); a row is waited for / released individually, so the window rolls smoothly;
//    dilated convs walk the strip as `dil` interleaved chains (rows c, c+d, c+2d, ...);
//  * the nine [N x 64] weight tiles stay resident in shared memory per (sample, strip run);
//  * R = 4 output rows are accumulated per TMEM hand-off (4 x N columns per stage, two stages), which
//    amortises the epilogue round trip;
//  * the epilogue (8 warps) stages each 32-pixel x N-channel result in swizzled shared memory and writes it
//    with one TMA store (full 128-byte lines) when the output is NHWC bf16; NCHW fp32 outputs are written
//    directly (already coalesced across pixels);
//  * channel blocks shorter than 64 skip the zero-padded K steps.
// 1x1 convolutions (TAPS = 1) use the same machinery without halo rows.
//
// Warp roles (12 warps): 0 = activation-row producer, 1 = MMA issuer (+ TMEM alloc), 2 = weight producer,
// 3 = idle, 4..11 = epilogue (two warps per TMEM lane quadrant, alternating output rows).
#include "conv_common.cuh"

#include <stdlib.h>

namespace vsp {
namespace {

constexpr int kRingThreads = 384;
constexpr int kMaxSlots = 16;
constexpr int kRingR = 4;

struct RingUnit {
  int b, strip, o0, L, slice, d;
};

// unit -> (sample, strip, chain, segment[, slice]); the host guarantees every unit has L >= 1 rows.
// Branch slices (the dilated branches of a SMART layer in one launch): slice j has its own dilation n_dil[j]; the
// rr_segs sub-units of a strip are split into n_dil[j] chains x rr_segs / n_dil[j] segments of rr_L rows, so every
// slice has the same number of equally long units and a CTA (grid = multiple of the slice count) keeps its branch.
__device__ __forceinline__ RingUnit ring_decode(const ConvParams &p, long long u, int d_eff) {
  RingUnit r;
  r.slice = (int)(u % p.rr_nslices); u /= p.rr_nslices;   // fastest: the slices of one strip segment run side by side (L2 reuse)
  if (p.branch_mode) {
    const int sub = (int)(u % p.rr_segs); u /= p.rr_segs;
    r.strip = (int)(u % p.rr_strips); u /= p.rr_strips;
    r.b = (int)u;
    r.d = p.n_dil[r.slice];
    const int chain = sub % r.d, seg = sub / r.d;
    r.L = p.rr_L;
    r.o0 = chain + seg * p.rr_L * r.d;
    return r;
  }
  const int seg = (int)(u % p.rr_segs); u /= p.rr_segs;
  const int chain = (int)(u % p.rr_chains); u /= p.rr_chains;
  r.strip = (int)(u % p.rr_strips); u /= p.rr_strips;
  r.b = (int)u;
  r.d = d_eff;
  const int rows_in_chain = (p.out_h - chain + d_eff - 1) / d_eff;
  const int start = seg * p.rr_L;
  r.L = min(p.rr_L, rows_in_chain - start);
  r.o0 = chain + start * d_eff;
  return r;
}

template <int BLOCK_N, int TAPS, bool STAGED>
__global__ void __launch_bounds__(kRingThreads, 1)
conv_ring_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_o) {
  constexpr int R = kRingR;
  constexpr int HALO = TAPS == 9 ? 1 : 0;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int CHUNK = BLOCK_N < 32 ? BLOCK_N : 32;
  constexpr int NCH = BLOCK_N / CHUNK;
  constexpr int ACC_COLS = R * BLOCK_N;
  constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  constexpr int ROW_BYTES = CHUNK * 2;                // one pixel of a staged output chunk (32 / 64 B)
  constexpr int SBUF_BYTES = 32 * ROW_BYTES;          // one staging buffer: 32 pixels x CHUNK channels
  constexpr int STAGE_BYTES = STAGED ? 2 * SBUF_BYTES : 0;   // two buffers per epilogue warp

  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.rr_S;
  const int d = TAPS == 9 ? p.halo_d : 1;       // tap spacing == chain stride
  const uint32_t slot_bytes = (uint32_t)p.halo_w * 128u;
  constexpr uint32_t b_total = (uint32_t)TAPS * B_BYTES;
  unsigned char *a_buf = smem;
  unsigned char *b_buf = smem + (size_t)S * slot_bytes;
  unsigned char *o_buf = b_buf + ((b_total + 1023u) & ~1023u);        // 8 x STAGE_BYTES, 1024-aligned
  uint64_t *bars = reinterpret_cast<uint64_t *>(o_buf + 8 * STAGE_BYTES);
  uint64_t *a_full = bars, *a_empty = bars + kMaxSlots;
  uint64_t *b_full = bars + 2 * kMaxSlots, *b_empty = b_full + 1;
  uint64_t *tmem_full = b_empty + 1, *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  float *epi_vec = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + 512);  // [3][BLOCK_N]

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (STAGED) tma_prefetch_desc(&tmap_o);
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long units = (long long)p.batch * p.rr_strips * p.rr_chains * p.rr_segs * p.rr_nslices;

  if (warp == 0) {
    // ===================== activation rows: each input row of the strip is loaded once =====================
    int slot = 0;
    uint32_t ph = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      const int nrows = un.L + 2 * HALO;
      const int w0 = un.strip * kBlockM - HALO * d;
      for (int k = 0; k < nrows; ++k) {
        const int ih = un.o0 + (k - HALO) * d;
        mbar_wait(&a_empty[slot], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&a_full[slot], slot_bytes);
          tma_load_4d(a_buf + (size_t)slot * slot_bytes, &tmap_a, &a_full[slot], 0, w0, ih, un.b);
        }
        __syncwarp();
        if (++slot == S) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================== weights: resident per sample (per-sample modulated) or for the whole launch =====================
    int cur_g = -1;
    uint32_t res_ph = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      const int g = p.groups == 1 ? 0 : un.b;
      if (g == cur_g) continue;
      mbar_wait(b_empty, res_ph ^ 1);          // every MMA that read the old weights has retired
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full, b_total);
        for (int tap = 0; tap < TAPS; ++tap)
          tma_load_4d(b_buf + (size_t)tap * B_BYTES, &tmap_b, b_full, 0, 0, p.tap_w[tap], g);
      }
      __syncwarp();
      cur_g = g;
      res_ph ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one output row (9 taps) at a time over the rolling window =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
    const uint32_t a_base = smem_u32(a_buf), b_base = smem_u32(b_buf);
    const int kk = min(kBlockK / kUmmaK, (p.cin + kUmmaK - 1) / kUmmaK);   // skip zero-padded K steps
    int wslot = 0, base = 0, acc = 0;
    uint32_t wph = 0, acc_phase = 0, res_ph = 0;
    int cur_g = -1;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      const int g = p.groups == 1 ? 0 : un.b;
      if (g != cur_g) {
        mbar_wait(b_full, res_ph);
        res_ph ^= 1;
        cur_g = g;
      }
      bool release_b = u + gridDim.x >= units;
      if (!release_b && p.groups != 1) release_b = ring_decode(p, u + gridDim.x, d).b != un.b;
      int rows_waited = 0;
      for (int j = 0; j < un.L; ++j) {
        const int rr = j % R;
        const bool last = j == un.L - 1;
        if (rr == 0) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tcgen05_fence_after();
        }
        while (rows_waited < j + 1 + 2 * HALO) {
          mbar_wait(&a_full[wslot], wph);
          if (++wslot == S) { wslot = 0; wph ^= 1; }
          ++rows_waited;
        }
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS + rr * BLOCK_N);
          uint32_t ra[3];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            int s = base + q;
            if (s >= S) s -= S;
            ra[q] = a_base + (uint32_t)s * slot_bytes;
          }
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap) {
            const int kh = TAPS == 9 ? tap / 3 : 0, kw = TAPS == 9 ? tap % 3 : 0;
            const uint64_t adesc = umma_smem_desc(ra[kh] + (uint32_t)(kw * d) * 128u, 128);
            const uint64_t bdesc = umma_smem_desc(b_base + (uint32_t)tap * B_BYTES, 128);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              if (k < kk)
                umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                             (tap > 0 || k > 0) ? 1u : 0u);
          }
          // input row j is not needed by any later output row; the unit's two trailing halo rows go with the last one
          umma_commit(&a_empty[base]);
          if (last && HALO) {
            int s = base + 1;
            if (s >= S) s -= S;
            umma_commit(&a_empty[s]);
            if (++s >= S) s -= S;
            umma_commit(&a_empty[s]);
          }
          if (rr == R - 1 || last) umma_commit(&tmem_full[acc]);
          if (last && release_b) umma_commit(b_empty);
        }
        __syncwarp();
        base += (last ? 1 + 2 * HALO : 1);
        while (base >= S) base -= S;
        if (rr == R - 1 || last) {
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int et = threadIdx.x - 128;
    float *vec_rs = epi_vec, *vec_b1 = epi_vec + BLOCK_N, *vec_b2 = epi_vec + 2 * BLOCK_N;
    const float nw = p.noise ? (p.noise_weight_dev ? __ldg(p.noise_weight_dev) : p.noise_weight) : 0.f;
    const long long plane = (long long)p.full_h * p.full_w;
    unsigned char *stage = o_buf + (warp - 4) * STAGE_BYTES;
    // 16-byte chunk swizzle of a staged row (must match the TMA store's swizzle mode: 64B or 32B rows)
    const int sw = ROW_BYTES == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1);
    const LeanK lk = lean_consts(p);
    int acc = 0;
    uint32_t acc_phase = 0, sbuf = 0;
    int cur_b = -1;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      if (un.b != cur_b) {
        // per-channel vectors change only with the sample: restage them once per run of units
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = et; c < BLOCK_N; c += 256) {
          const bool ok = c < p.cout;
          const float rs = (ok && p.row_scale) ? __ldg(p.row_scale + (long long)un.b * p.cout + c) : 1.f;
          const float b1 = (ok && p.pre_bias) ? __ldg(p.pre_bias + c) : 0.f;
          const float b2 = (ok && p.bias) ? __ldg(p.bias + c) : 0.f;
          vec_rs[c] = STAGED ? lean_scale_rs(lk, rs) : rs;      // staged path: activation gains folded in
          vec_b1[c] = STAGED ? lean_scale_b1(lk, b1) : b1;
          vec_b2[c] = STAGED ? lean_scale_b2(lk, b2) : b2;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_b = un.b;
      }
      const int ow = un.strip * kBlockM + row;
      const bool pix_ok = ow < p.out_w;
      const int nsteps = (un.L + R - 1) / R;
      for (int st = 0; st < nsteps; ++st) {
        const int reff = min(R, un.L - st * R);
        // noise of this warp's rows is fetched before waiting for the accumulators
        float nz[R / 2];
        const int oh0 = un.o0 + st * R * d;
#pragma unroll
        for (int q = 0; q < R / 2; ++q) {
          const int rr = 2 * q + wg;
          nz[q] = 0.f;
          if (rr < reff && p.noise != nullptr && pix_ok)
            nz[q] = (STAGED ? 1.f : nw) *       // staged form: raw value, scaled at its first use after the accumulator wait
                    __ldg(p.noise + un.b * p.noise_bstride + (long long)(oh0 + rr * d) * p.full_w + ow);
        }
        // output rows alternate between the two epilogue warps of this lane quadrant: rows wg, wg + 2
        const int n_items = ((reff - wg + 1) >> 1) * NCH;
        mbar_wait(&tmem_full[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_COLS);
        if (n_items <= 0) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
#pragma unroll 1
        for (int it = 0; it < n_items; ++it) {
          const int q = it / NCH, ch = it % NCH;
          const int rr = 2 * q + wg;
          const int oh = oh0 + rr * d;
          float nzv = nz[0];
#pragma unroll
          for (int i = 1; i < R / 2; ++i) nzv = q == i ? nz[i] : nzv;
          if constexpr (STAGED) {
            uint32_t r[CHUNK];
            if constexpr (CHUNK == 32) tmem_ld_32x32b_x32(taddr + rr * BLOCK_N + ch * CHUNK, r);
            else tmem_ld_32x32b_x16(taddr + rr * BLOCK_N + ch * CHUNK, reinterpret_cast<uint32_t(&)[16]>(r));
            tmem_ld_wait();
            if (it == n_items - 1) {          // accumulators are in registers: the MMA warp may reuse this stage
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
            unsigned char *buf = stage + (sbuf & 1) * SBUF_BYTES;
            ++sbuf;
            if (lane == 0) bulk_wait_group_read<1>();   // the TMA store issued two chunks ago has drained this buffer
            __syncwarp();
#pragma unroll
            for (int i = 0; i < CHUNK / 8; ++i)
              *reinterpret_cast<uint4 *>(buf + lane * ROW_BYTES + ((i ^ sw) << 4)) =
                  epi_lean8(&r[8 * i], vec_rs + ch * CHUNK + 8 * i, vec_b1 + ch * CHUNK + 8 * i,
                            vec_b2 + ch * CHUNK + 8 * i, nw * nzv * lk.m2, lk);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmap_o, buf, (int)p.co_off + ch * CHUNK, un.strip * kBlockM + quad * 32, oh, un.b);
              bulk_commit_group();
            }
          } else {
            const long long pix = (long long)oh * p.full_w + ow;
            float v[CHUNK];
            epi_compute<CHUNK>(p, taddr + rr * BLOCK_N + ch * CHUNK, ch * CHUNK, p.cout, un.b, pix, plane, pix_ok, nzv,
                               vec_rs + ch * CHUNK, vec_b1 + ch * CHUNK, vec_b2 + ch * CHUNK, nullptr, v);
            if (it == n_items - 1) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
            epi_store_direct<CHUNK>(p, ch * CHUNK, p.cout, un.b, pix, plane, pix_ok, v);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    if (STAGED && lane == 0) bulk_wait_group<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// kh-folded variant for very narrow layers (Cout <= 32: the dilated SMART branches at 256^2 / 512^2 and the 32-channel
// 1024^2 layer).  With N = Cout <= 32 the nine-tap formulation above is bound by the tensor core's shared-memory
// OPERAND bandwidth, not by math or HBM: every MMA re-reads its 128 x 16 activation slab (32 wavefronts) to do only
// 128 x N x 16 MACs (ncu: l1tex__data_pipe_tc_wavefronts at 55 % with the tensor pipe 12 % busy).  Here the three
// vertical taps are folded into the GEMM's N dimension instead: for each INPUT row and each horizontal tap kw
//     Z_k[p, (kh, c)] += X_k[p + (kw-1) d, :] . W[kh, kw, c, :]        (one MMA chain with N = 3 C)
// so an input row is read three times (once per kw) instead of nine, and the output row r is assembled in the epilogue
// from the TMEM blocks of three consecutive input rows:  out_r = Z_{r-1}[kh=0] + Z_r[kh=1] + Z_{r+1}[kh=2]  (lane-wise
// adds, no cross-lane traffic).  TMEM is a ring of NR input-row slots of 3 C columns; an input row's shared-memory
// slot is released as soon as its own MMAs retire (each row is consumed once), its TMEM slot after the three output
// rows that read it.  Cin up to 128 (two 64-channel blocks accumulated into the same slots).
template <int C, bool STAGED>
__global__ void __launch_bounds__(kRingThreads, 1)
conv_ringfold_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap_a,
                     const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ OutMaps omaps) {
  constexpr int NR = (512 / (3 * C));                // TMEM row slots: 10 (C = 16) or 5 (C = 32)
  constexpr int B_BYTES = C * 128;                    // one (kh, kw) weight tile of a 64-channel block
  constexpr int ROW_BYTES = C * 2;
  constexpr int SBUF_BYTES = 32 * ROW_BYTES;
  constexpr int STAGE_BYTES = STAGED ? 2 * SBUF_BYTES : 0;

  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.rr_S, kc = p.kc;
  const int d = p.halo_d;                 // (branch slices: per-unit dilation un.d, halo box sized for the largest)
  const uint32_t slot_bytes = (uint32_t)p.halo_w * 128u;
  const uint32_t b_total = (uint32_t)(kc * 9) * B_BYTES;
  unsigned char *a_buf = smem;
  unsigned char *b_buf = smem + (size_t)S * slot_bytes;
  unsigned char *o_buf = b_buf + ((b_total + 1023u) & ~1023u);
  uint64_t *bars = reinterpret_cast<uint64_t *>(o_buf + 8 * STAGE_BYTES);
  uint64_t *a_full = bars, *a_empty = bars + kMaxSlots;
  uint64_t *b_full = bars + 2 * kMaxSlots, *b_empty = b_full + 1;
  uint64_t *acc_full = b_empty + 1, *acc_empty = acc_full + NR;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + NR);
  float *epi_vec = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + 512);  // [3][C]

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (STAGED) tma_prefetch_desc(&omaps.m[0]);
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    for (int i = 0; i < NR; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 12);   // three reading output rows x four lane-quadrant warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long units = (long long)p.batch * p.rr_strips * p.rr_chains * p.rr_segs * p.rr_nslices;

  if (warp == 0) {
    // ===================== activation rows (kc slots per row) =====================
    int slot = 0;
    uint32_t ph = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      const int w0 = un.strip * kBlockM - un.d;
      for (int k = 0; k < un.L + 2; ++k) {
        const int ih = un.o0 + (k - 1) * un.d;
        for (int cb = 0; cb < kc; ++cb) {
          mbar_wait(&a_empty[slot], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[slot], slot_bytes);
            tma_load_4d(a_buf + (size_t)slot * slot_bytes, &tmap_a, &a_full[slot], cb * kBlockK, w0, ih, un.b);
          }
          __syncwarp();
          if (++slot == S) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weights, resident: smem tile order [cb][kw][kh] so that one kw is an [3C x 64] operand =====
    int cur_key = -1;
    uint32_t res_ph = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      const int g = p.groups == 1 ? 0 : un.b;
      const int key = g * p.rr_nslices + un.slice;
      if (key == cur_key) continue;
      mbar_wait(b_empty, res_ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full, b_total);
        for (int cb = 0; cb < kc; ++cb)
          for (int kw = 0; kw < 3; ++kw)
            for (int kh = 0; kh < 3; ++kh)
              tma_load_4d(b_buf + (size_t)((cb * 3 + kw) * 3 + kh) * B_BYTES, &tmap_b, b_full, cb * kBlockK,
                          un.slice * C, p.tap_w[kh * 3 + kw], g);
      }
      __syncwarp();
      cur_key = key;
      res_ph ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one input row = 3 (kw) x kk MMAs with N = 3C =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, 3 * C);
    const uint32_t a_base = smem_u32(a_buf), b_base = smem_u32(b_buf);
    int aslot = 0, rslot = 0;
    uint32_t aph = 0, rph = 0, res_ph = 0;
    int cur_key = -1;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      const int key = (p.groups == 1 ? 0 : un.b) * p.rr_nslices + un.slice;
      if (key != cur_key) {
        mbar_wait(b_full, res_ph);
        res_ph ^= 1;
        cur_key = key;
      }
      bool release_b = u + gridDim.x >= units;
      if (!release_b) {
        const RingUnit nx = ring_decode(p, u + gridDim.x, d);
        release_b = ((p.groups == 1 ? 0 : nx.b) * p.rr_nslices + nx.slice) != key;
      }
      for (int k = 0; k < un.L + 2; ++k) {
        mbar_wait(&acc_empty[rslot], rph ^ 1);        // the output rows that read this TMEM slot last time are done
        int s0 = aslot;
        for (int cb = 0; cb < kc; ++cb) {
          mbar_wait(&a_full[aslot], aph);
          if (++aslot == S) { aslot = 0; aph ^= 1; }
        }
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(rslot * 3 * C);
          int s = s0;
          for (int cb = 0; cb < kc; ++cb) {
            const int kk = min(kBlockK / kUmmaK, (p.cin - cb * kBlockK + kUmmaK - 1) / kUmmaK);
            const uint32_t arow = a_base + (uint32_t)s * slot_bytes;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint64_t adesc = umma_smem_desc(arow + (uint32_t)(kw * un.d) * 128u, 128);
              const uint64_t bdesc = umma_smem_desc(b_base + (uint32_t)((cb * 3 + kw) * 3) * B_BYTES, 128);
#pragma unroll
              for (int ks = 0; ks < kBlockK / kUmmaK; ++ks)
                if (ks < kk)
                  umma_bf16_ss(d_tmem, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc,
                               (cb > 0 || kw > 0 || ks > 0) ? 1u : 0u);
            }
            umma_commit(&a_empty[s]);                  // this row block is consumed exactly once
            if (++s == S) s = 0;
          }
          umma_commit(&acc_full[rslot]);
          if (k == un.L + 1 && release_b) umma_commit(b_empty);
        }
        __syncwarp();
        if (++rslot == NR) { rslot = 0; rph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps): out_j = Z_j[kh=0] + Z_{j+1}[kh=1] + Z_{j+2}[kh=2] =====================
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int et = threadIdx.x - 128;
    float *vec_rs = epi_vec, *vec_b1 = epi_vec + C, *vec_b2 = epi_vec + 2 * C;
    const float nw = p.noise ? (p.noise_weight_dev ? __ldg(p.noise_weight_dev) : p.noise_weight) : 0.f;
    const long long plane = (long long)p.full_h * p.full_w;
    unsigned char *stage = o_buf + (warp - 4) * STAGE_BYTES;
    const int sw = ROW_BYTES == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1);
    const LeanK lk = lean_consts(p);
    // channel slices: GEMM columns [slice*C, slice*C + C); with the pixel-shuffle mapping the column is
    // class * shuffle_cout + channel, so a slice is (class, channel part)
    const int creal = p.shuffle_cout ? p.shuffle_cout : p.cout;
    const int parts = p.shuffle_cout ? p.shuffle_cout / C : 1;
    const __nv_bfloat16 *res1 = static_cast<const __nv_bfloat16 *>(p.residual);
    const __nv_bfloat16 *res2 = static_cast<const __nv_bfloat16 *>(p.residual2);
    const bool has_res = STAGED && (res1 != nullptr || res2 != nullptr);
    constexpr int PIECES = C / 8, PPR = 32 / PIECES;
    uint32_t sbuf = 0;
    long long rg0 = 0;                                  // running index of the unit's first input row
    int cur_key = -1;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      const int cls = p.shuffle_cout ? un.slice / parts : 0;
      const int c0 = p.shuffle_cout ? (un.slice % parts) * C : un.slice * C;
      const int key = un.b * p.rr_nslices + un.slice;
      if (key != cur_key) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = et; c < C; c += 256) {
          const bool ok = c0 + c < creal;
          const float rs = (ok && p.row_scale) ? __ldg(p.row_scale + (long long)un.b * creal + c0 + c) : 1.f;
          const float b1 = (ok && p.pre_bias) ? __ldg(p.pre_bias + c0 + c) : 0.f;
          const float b2 = (ok && p.bias) ? __ldg(p.bias + c0 + c) : 0.f;
          vec_rs[c] = STAGED ? lean_scale_rs(lk, rs) : rs;      // staged path: activation gains folded in
          vec_b1[c] = STAGED ? lean_scale_b1(lk, b1) : b1;
          vec_b2[c] = STAGED ? lean_scale_b2(lk, b2) : b2;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_key = key;
      }
      const int ow = un.strip * kBlockM + row;
      const bool pix_ok = ow < p.out_w;
      const int fw = ow * p.os + (cls & 1);
      const int L = un.L;
      // residuals (staged path): sector-coalesced warp loads of the 32-pixel x C-channel tile, one row ahead
      uint4 pr1[PIECES], pr2[PIECES];
      auto load_res = [&](int j) {
        const int fh = (un.o0 + j * un.d) * p.os + (cls >> 1);
#pragma unroll
        for (int i = 0; i < PIECES; ++i) {
          const int px = un.strip * kBlockM + quad * 32 + i * PPR + lane / PIECES;
          const bool ok = px < p.out_w;
          const long long off = (((long long)un.b * p.full_h + fh) * p.full_w + px * p.os + (cls & 1)) * p.ldo + p.co_off +
                                c0 + (lane % PIECES) * 8;
          pr1[i] = (res1 && ok) ? __ldg(reinterpret_cast<const uint4 *>(res1 + off)) : make_uint4(0u, 0u, 0u, 0u);
          pr2[i] = (res2 && ok) ? __ldg(reinterpret_cast<const uint4 *>(res2 + off)) : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      if (has_res && wg < L) load_res(wg);
      // noise one row ahead as well (an HBM round trip per row would otherwise sit on the epilogue's critical path)
      auto load_noise = [&](int j) -> float {
        if (p.noise == nullptr || !pix_ok || j >= L) return 0.f;
        return __ldg(p.noise + un.b * p.noise_bstride + (long long)((un.o0 + j * un.d) * p.os + (cls >> 1)) * p.full_w + fw);
      };
      float nz_next = load_noise(wg);
      for (int j = wg; j < L; j += 2) {
        const int oh = un.o0 + j * un.d;
        const long long pix = (long long)(oh * p.os + (cls >> 1)) * p.full_w + fw;
        const float nz = nw * nz_next;
        nz_next = load_noise(j + 2);
        const long long r0 = rg0 + j;
        const int s0 = (int)(r0 % NR), s1 = (int)((r0 + 1) % NR), s2 = (int)((r0 + 2) % NR);
        mbar_wait(&acc_full[s2], (uint32_t)(((r0 + 2) / NR) & 1));      // MMAs retire in order: rows r0, r0+1 are done too
        tcgen05_fence_after();
        const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t r[C], t[C];
        if constexpr (C == 32) {
          tmem_ld_32x32b_x32(tq + s0 * 3 * C, r);
          tmem_ld_32x32b_x32(tq + s1 * 3 * C + C, t);
        } else {
          tmem_ld_32x32b_x16(tq + s0 * 3 * C, reinterpret_cast<uint32_t(&)[16]>(r));
          tmem_ld_32x32b_x16(tq + s1 * 3 * C + C, reinterpret_cast<uint32_t(&)[16]>(t));
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < C; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(t[i]));
        if constexpr (C == 32) tmem_ld_32x32b_x32(tq + s2 * 3 * C + 2 * C, t);
        else tmem_ld_32x32b_x16(tq + s2 * 3 * C + 2 * C, reinterpret_cast<uint32_t(&)[16]>(t));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < C; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(t[i]));
        // release the three TMEM slots; output rows that do not exist at the unit's ends are accounted for by the
        // row min(k, L-1), which reads slot k in any case
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int k = j + q;                                   // unit-local input row
            const int readers = min(k, L - 1) - max(k - 2, 0) + 1;
            const int mine = 1 + ((min(k, L - 1) == j) ? 3 - readers : 0);
            mbar_arrive_n(&acc_empty[q == 0 ? s0 : (q == 1 ? s1 : s2)], (uint32_t)mine);
          }
        }
        if constexpr (STAGED) {
          unsigned char *buf = stage + (sbuf & 1) * SBUF_BYTES;
          ++sbuf;
          if (lane == 0) bulk_wait_group_read<1>();
          __syncwarp();
          // activation in place (r holds fp32 bits), then the residual tiles are transposed through the staging
          // buffer one after the other and added, so only one transposed tile is live at a time
#pragma unroll
          for (int i = 0; i < PIECES; ++i) {
            float v[8];
            epi_lean8f(&r[8 * i], vec_rs + 8 * i, vec_b1 + 8 * i, vec_b2 + 8 * i, nz * lk.m2, lk, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[8 * i + e] = __float_as_uint(v[e]);
          }
          if (has_res) {
#pragma unroll
            for (int which = 0; which < 2; ++which) {
#pragma unroll
              for (int i = 0; i < PIECES; ++i) {
                const int px = i * PPR + lane / PIECES;
                const int swp = ROW_BYTES == 64 ? ((px >> 1) & 3) : ((px >> 2) & 1);
                *reinterpret_cast<uint4 *>(buf + px * ROW_BYTES + (((lane % PIECES) ^ swp) << 4)) = which == 0 ? pr1[i] : pr2[i];
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < PIECES; ++i) {
                const uint4 own = *reinterpret_cast<const uint4 *>(buf + lane * ROW_BYTES + ((i ^ sw) << 4));
                const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&own);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  r[8 * i + 2 * e] = __float_as_uint(__uint_as_float(r[8 * i + 2 * e]) + f.x);
                  r[8 * i + 2 * e + 1] = __float_as_uint(__uint_as_float(r[8 * i + 2 * e + 1]) + f.y);
                }
              }
              __syncwarp();
            }
            if (j + 2 < L) load_res(j + 2);
          }
#pragma unroll
          for (int i = 0; i < PIECES; ++i) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * i + e]);
            *reinterpret_cast<uint4 *>(buf + lane * ROW_BYTES + ((i ^ sw) << 4)) = pack8_bf16(v);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&omaps.m[cls], buf, (int)p.co_off + c0, un.strip * kBlockM + quad * 32, oh, un.b);
            bulk_commit_group();
          }
        } else {
          // direct stores (fp32 NCHW outputs / unaligned layouts): same arithmetic on the summed accumulator
          float v[C];
#pragma unroll
          for (int i = 0; i < C; ++i) {
            float xv = __uint_as_float(r[i]) * vec_rs[i];
            if (p.pre_act) xv = epi_act(xv + vec_b1[i], p.pre_act, p.alpha, p.scale);
            xv = epi_act(xv + nz + vec_b2[i], p.act, p.alpha, p.scale);
            v[i] = xv;
          }
          if (pix_ok && (p.residual || p.residual2)) {
#pragma unroll
            for (int i = 0; i < C; ++i) {
              if (c0 + i >= creal) break;
              if (!p.out_nhwc) {
                const long long off = ((long long)un.b * creal + c0 + i) * plane + pix;
                if (p.residual) v[i] += __ldg(static_cast<const float *>(p.residual) + off);
                if (p.residual2) v[i] += __ldg(static_cast<const float *>(p.residual2) + off);
              } else {
                const long long off = ((long long)un.b * plane + pix) * p.ldo + p.co_off + c0 + i;
                if (p.residual) v[i] += __bfloat162float(static_cast<const __nv_bfloat16 *>(p.residual)[off]);
                if (p.residual2) v[i] += __bfloat162float(static_cast<const __nv_bfloat16 *>(p.residual2)[off]);
              }
            }
          }
          epi_store_direct<C>(p, c0, creal, un.b, pix, plane, pix_ok, v);
        }
      }
      rg0 += L + 2;
    }
    if (STAGED && lane == 0) bulk_wait_group<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int C, bool STAGED>
int launch_ringfold(const ConvParams &p, const void *x, const void *wq, int64_t in_h, int64_t in_w, int64_t cout_pad,
                    int taps_total, size_t smem_bytes, cudaStream_t stream) {
  auto kern = conv_ringfold_kernel<C, STAGED>;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  CUtensorMap ta, tb;
  OutMaps om;
  memset(&om, 0, sizeof(om));
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)p.batch};
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
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)C, 1, 1};
    if (int rc = encode_tma(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, wq, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  if (STAGED) {
    // one map per output parity class (pixel-shuffle: class origin in the base pointer, pixel strides x os)
    const int ncls = p.shuffle_cout ? 4 : 1;
    const int os = p.os;
    const int creal = p.shuffle_cout ? p.shuffle_cout : p.cout;
    for (int cls = 0; cls < ncls; ++cls) {
      const int oh0 = cls >> 1, ow0 = cls & 1;
      const __nv_bfloat16 *base = static_cast<const __nv_bfloat16 *>(p.out) + ((long long)oh0 * p.full_w + ow0) * p.ldo;
      uint64_t dims[4] = {(uint64_t)(p.co_off + creal), (uint64_t)((p.full_w - ow0 + os - 1) / os),
                          (uint64_t)((p.full_h - oh0 + os - 1) / os), (uint64_t)p.batch};
      uint64_t strides[4] = {0, (uint64_t)p.ldo * 2 * os, (uint64_t)p.ldo * p.full_w * 2 * os,
                             (uint64_t)p.ldo * p.full_w * p.full_h * 2};
      uint32_t box[4] = {(uint32_t)C, 32, 1, 1};
      if (int rc = encode_tma(&om.m[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, nullptr,
                              C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B))
        return rc;
    }
  }
  const long long units = (long long)p.batch * p.rr_strips * p.rr_chains * p.rr_segs * p.rr_nslices;
  // a CTA keeps its channel slice (weights stay resident): grid = a multiple of the slice count
  long long grid = (num_sms() / p.rr_nslices) * p.rr_nslices;
  if (grid < p.rr_nslices) grid = p.rr_nslices;
  if (units < grid) grid = units;
  kern<<<(unsigned)grid, kRingThreads, smem_bytes, stream>>>(p, ta, tb, om);
  return check_launch("conv_ringfold_kernel");
}

template <int BLOCK_N, int TAPS, bool STAGED>
int launch_ring(const ConvParams &p, const void *x, const void *wq, int64_t in_h, int64_t in_w, int64_t cout_pad,
                int taps_total, size_t smem_bytes, cudaStream_t stream) {
  auto kern = conv_ring_kernel<BLOCK_N, TAPS, STAGED>;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  CUtensorMap ta, tb, to;
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)p.batch};
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
  if (STAGED) {
    // output viewed as (channels visible to this launch, w, h, b); the box is one warp's 32 pixels x CHUNK
    // channels; channels beyond co_off + cout and pixels beyond the width are clipped by TMA
    uint64_t dims[4] = {(uint64_t)(p.co_off + p.cout), (uint64_t)p.full_w, (uint64_t)p.full_h, (uint64_t)p.batch};
    uint64_t strides[4] = {0, (uint64_t)p.ldo * 2, (uint64_t)p.ldo * p.full_w * 2,
                           (uint64_t)p.ldo * p.full_w * p.full_h * 2};
    constexpr int CHUNK = BLOCK_N < 32 ? BLOCK_N : 32;
    uint32_t box[4] = {(uint32_t)CHUNK, 32, 1, 1};
    const CUtensorMapSwizzle swz = CHUNK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    if (int rc = encode_tma(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p.out, dims, strides, box, nullptr, swz))
      return rc;
  } else {
    to = ta;
  }
  const long long units = (long long)p.batch * p.rr_strips * p.rr_chains * p.rr_segs * p.rr_nslices;
  const long long grid = units < num_sms() ? units : num_sms();
  kern<<<(unsigned)grid, kRingThreads, smem_bytes, stream>>>(p, ta, tb, to);
  return check_launch("conv_ring_kernel");
}

template <int BLOCK_N>
int dispatch_ring(const ConvParams &p, int taps, const void *x, const void *wq, int64_t in_h, int64_t in_w,
                  int64_t cout_pad, int taps_total, size_t smem_bytes, cudaStream_t stream) {
  if (p.rr_staged) {
    if (taps == 9) return launch_ring<BLOCK_N, 9, true>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
    return launch_ring<BLOCK_N, 1, true>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
  }
  if (taps == 9) return launch_ring<BLOCK_N, 9, false>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
  return launch_ring<BLOCK_N, 1, false>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
}

}  // namespace

int conv_ring_try_launch(ConvParams &p, const void *x, const void *wq, int64_t in_h, int64_t in_w,
                         int64_t cout_pad, int taps_total, int dil, cudaStream_t stream) {
  static const bool disabled = getenv("VSP_NO_RING") != nullptr;
  static const bool no_stage = getenv("VSP_RING_NO_TMA_STORE") != nullptr;
  static const bool no_fold = getenv("VSP_NO_FOLD") != nullptr;
  static const bool no_slices = getenv("VSP_NO_FOLD_SLICES") != nullptr;
  if (disabled) return -1;
  if (p.stride != 1 || p.oo_h != 0 || p.oo_w != 0) return -1;
  if (p.out_w < kBlockM || p.out_w != in_w || p.out_h != in_h) return -1;
  if (p.ntaps != 9 && p.ntaps != 1) return -1;
  // branch slices: the dilated branches of a SMART layer (cout / n_branches <= 32 channels each) in one launch
  int nbr = 0;
  if (p.branch_mode) {
    for (int j = 0; j < 4 && p.n_dil[j] > 0; ++j) nbr = j + 1;
    const int cq = nbr ? p.cout / nbr : 0;
    bool ok = !no_fold && p.ntaps == 9 && p.kc <= 2 && nbr >= 1 && (cq == 16 || cq == 32) && p.os == 1 &&
              p.shuffle_cout == 0 && p.full_w == p.out_w && p.full_h == p.out_h && p.residual == nullptr &&
              p.residual2 == nullptr;
    int dmax = 1;
    for (int j = 0; ok && j < nbr; ++j) {
      const int dj = p.n_dil[j];
      ok = dj >= 1 && dj <= 8 && (dj & (dj - 1)) == 0;
      dmax = dj > dmax ? dj : dmax;
    }
    if (!ok) return -1;
    int S = 16;
    while (S > dmax && p.out_h / S < 8) S >>= 1;
    if (S < dmax || p.out_h % S != 0 || p.out_h / S < 2) return -1;
    for (int t = 0; t < 9; ++t)
      if (p.tap_dy[t] != t / 3 - 1 || p.tap_dx[t] != t % 3 - 1) return -1;
    p.halo_d = dmax;
    p.halo_w = (kBlockM + 2 * dmax + 7) & ~7;
    p.rr_staged = (!no_stage && p.out_nhwc && (p.ldo % 8) == 0 && (p.co_off % 8) == 0 &&
                   (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && p.alpha >= 0.f && p.alpha <= 1.f &&
                   (p.scale > 0.f || (p.act == 0 && p.pre_act == 0))) ? 1 : 0;
    const int slot = p.halo_w * 128;
    const int b_bytes = (9 * p.kc * cq * 128 + 1023) & ~1023;
    const int o_bytes = p.rr_staged ? 8 * 2 * 32 * cq * 2 : 0;
    const int fixed = 1024 + b_bytes + o_bytes + 512 + 3 * cq * 4;
    int slots = (232448 - fixed) / slot;
    if (slots > kMaxSlots) slots = kMaxSlots;
    if (slots < 2 * p.kc) return -1;
    p.rr_R = kRingR; p.rr_S = slots; p.rr_nb = 0;
    p.tiles_n = 1;
    p.rr_nslices = nbr;
    p.rr_strips = (p.out_w + kBlockM - 1) / kBlockM;
    p.rr_chains = 1;
    p.rr_segs = S;
    p.rr_L = p.out_h / S;
    const size_t smem_bytes = (size_t)fixed + (size_t)slots * slot;
    if (cq == 16)
      return p.rr_staged ? launch_ringfold<16, true>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream)
                         : launch_ringfold<16, false>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
    return p.rr_staged ? launch_ringfold<32, true>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream)
                       : launch_ringfold<32, false>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
  }
  // kh-folded kernel: Cout <= 32 directly, wider layers (and the pixel-shuffle up-convolution) as 32-channel slices
  // (measured: four slices win clearly — 128->128 @256^2 172 -> 148 us, fused 64->32 up-conv 428 -> 328 us; two slices
  // of a 64-channel block only tie with the resident nine-tap ring, eight slices lose to the generic kernel)
  const bool fold = !no_fold && p.ntaps == 9 && p.kc <= 2 &&
                    (p.cout <= 32 || (!no_slices && p.cout % 32 == 0 && p.cout <= 128 && !(p.kc == 1 && p.cout == 64) &&
                                      (p.shuffle_cout == 0 || p.shuffle_cout % 32 == 0)));
  if (p.shuffle_cout ? !(fold && p.os == 2 && p.full_w == 2 * p.out_w && p.full_h == 2 * p.out_h)
                     : !(p.os == 1 && p.full_w == p.out_w && p.full_h == p.out_h))
    return -1;
  if (!fold && (p.kc != 1 || p.cout > 64)) return -1;
  const int taps = p.ntaps;
  const int halo = taps == 9 ? 1 : 0;
  int d = 1;
  if (taps == 9) {
    d = dil;
    if (d < 1 || d > 8) return -1;
    for (int t = 0; t < 9; ++t)
      if (p.tap_dy[t] != (t / 3 - 1) * d || p.tap_dx[t] != (t % 3 - 1) * d) return -1;
    if (p.out_h % d != 0 || p.out_h < 2 * d) return -1;     // equal-length chains keep every unit non-empty
  } else if (p.tap_dy[0] != 0 || p.tap_dx[0] != 0) {
    return -1;
  }
  int bn = 16;
  while (bn < p.cout) bn <<= 1;
  p.rr_nslices = 1;
  if (fold && p.cout > 32) {
    bn = 32;
    p.rr_nslices = p.cout / 32;
  }
  p.halo_d = d;
  p.halo_w = taps == 9 ? ((kBlockM + 2 * d + 7) & ~7) : kBlockM;
  p.rr_staged = (!no_stage && p.out_nhwc && (p.ldo % 8) == 0 && (p.co_off % 8) == 0 &&
                 (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
                 ((p.residual == nullptr && p.residual2 == nullptr) || fold) && p.alpha >= 0.f && p.alpha <= 1.f &&
                 (p.scale > 0.f || (p.act == 0 && p.pre_act == 0))) ? 1 : 0;
  if (p.shuffle_cout && !p.rr_staged) return -1;
  const int slot = p.halo_w * 128;
  const int b_bytes = (taps * p.kc * bn * 128 + 1023) & ~1023;
  const int chunk = bn < 32 ? bn : 32;
  const int o_bytes = p.rr_staged ? 8 * 2 * 32 * chunk * 2 : 0;
  const int fixed = 1024 + b_bytes + o_bytes + 512 + 3 * bn * 4;
  int S = (232448 - fixed) / slot;
  if (S > kMaxSlots) S = kMaxSlots;
  if (S < (fold ? 2 * p.kc : 1 + 2 * halo + 2)) return -1;
  p.rr_R = kRingR; p.rr_S = S; p.rr_nb = 0;
  p.tiles_n = 1;
  p.rr_strips = (p.out_w + kBlockM - 1) / kBlockM;
  p.rr_chains = d;
  // segments: minimise waves x (rows per unit + per-unit overhead)
  static const int unit_ovh = getenv("VSP_RING_UNIT_OVH") ? atoi(getenv("VSP_RING_UNIT_OVH")) : 8;   // pipeline fill/drain, in rows
  const int rows_chain = p.out_h / d;
  const long long base_units = (long long)p.batch * p.rr_strips * p.rr_chains * p.rr_nslices;
  double best = 1e300;
  int best_L = rows_chain;
  for (int segs = 1; segs <= rows_chain; ++segs) {
    int L = (rows_chain + segs - 1) / segs;
    L = (L + kRingR - 1) / kRingR * kRingR;
    const int se = (rows_chain + L - 1) / L;
    const long long units = base_units * se;
    const long long waves = (units + num_sms() - 1) / num_sms();
    const double cost = (double)waves * (L + 2 * halo + unit_ovh);
    if (cost < best) { best = cost; best_L = L; }
    if (L <= kRingR) break;
  }
  p.rr_L = best_L;
  p.rr_segs = (rows_chain + best_L - 1) / best_L;
  const size_t smem_bytes = (size_t)fixed + (size_t)S * slot;
  if (fold) {
    if (bn == 16)
      return p.rr_staged ? launch_ringfold<16, true>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream)
                         : launch_ringfold<16, false>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
    return p.rr_staged ? launch_ringfold<32, true>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream)
                       : launch_ringfold<32, false>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
  }
  switch (bn) {
    case 16: return dispatch_ring<16>(p, taps, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
    case 32: return dispatch_ring<32>(p, taps, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
    default: return dispatch_ring<64>(p, taps, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
  }
}

}  // namespace vsp
