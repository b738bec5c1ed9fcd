// conv_up2h_sm100.cu — up-sampling modulated convolution (conv_transpose2d stride 2, 3x3  ->  Blur 4x4 pad 1,
// models/RestoreNet.py:522-535) for the wide levels (W >= 128, Cin <= 128) at HALF the dense form's tensor work.
//
// The dense form (vsp_conv2d_up2_fused_bf16) composes the whole 4x4 blur into the weights: a 3x3 convolution with
// 4*Cout outputs = 4x the layer's algorithmic FLOPs, which makes the 128->64 @256^2 and 64->32 @512^2 layers
// tensor-bound although their data would stream in half the time.  Here only the HORIZONTAL half of the (separable) blur
// is composed into the weights; the vertical half runs in the epilogue on fp32 accumulators:
//
//   transposed conv     z[2a+pa][X]  = sum_{kh = pa (mod 2)} x[a - (kh-pa)/2] . W[kh]        (rows; columns alike)
//   horizontal blur     hz[Y][2j+q]  = sum_v kx[v] z[Y][2j+q+v-1]   ->  composed weights Wc[kh][q][dx], dx = -1..1
//   per INPUT row r     T_kh[r][j][q] = sum_dx x[r][j+dx] . Wc[kh][q][dx]                     (tensor core, 18 C Cin MACs/pixel)
//   hz rows             E[a] = T_0[a] + T_2[a-1]  (= hz[2a]),      O[a] = T_1[a]  (= hz[2a+1])
//   vertical blur       out[2a]   = ky0 O[a-1] + ky1 E[a] + ky2 O[a] + ky3 E[a+1]
//                       out[2a+1] = ky0 E[a]   + ky1 O[a] + ky2 E[a+1] + ky3 O[a+1]
//
// A CTA owns (sample, 128-column strip, 32-channel slice, row segment) and walks down the strip like the row-ring
// kernels: every input row (one TMA box with a one-pixel halo, out-of-bounds zero fill = both the transposed conv's and
// the blur's zero padding) is loaded once and consumed once; its MMAs write T_0 | T_1 into TMEM slot r (N = 128) and T_2
// into the E half of slot r+1 (N = 64), so E rows are completed by the accumulator itself.  TMEM is a ring of four
// 128-column slots; each slot is read exactly once: the epilogue keeps the three pending partial sums of the 4-tap
// vertical filter in registers (thread = pixel column, 32 channels), so nothing but accumulators ever waits in TMEM and
// the MMA warp runs two rows ahead.  The weights of the slice (kc x 3 x 192 rows) stay resident in shared memory.
// Epilogue per output row pair: demodulation, noise, bias, leaky ReLU, the two skip residuals, bf16, swizzled staging
// tile, TMA store through one tensor map per output parity class.
//
// Warp roles (12 warps): 0 = input-row producer, 1 = MMA issuer (+ TMEM alloc), 2 = weight producer, 3 = idle,
// 4..11 = epilogue (two per TMEM lane quadrant: output column parity q = 0 / 1).
#include "conv_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace vsp {
namespace {

constexpr int kUpThreads = 384;
constexpr int kUpC = 32;                         // channels per slice
constexpr int kUpSlotCols = 4 * kUpC;            // [E q0 | E q1 | O q0 | O q1]
constexpr int kUpNR = 512 / kUpSlotCols;         // 4 TMEM slots
constexpr int kUpHaloW = 136;                    // 128 + 2 halo pixels, padded to a multiple of 8 (1024-byte slots)
constexpr int kUpSlotBytes = kUpHaloW * 128;     // one (row, 64-channel block) of activations
constexpr int kUpBoxW = 130;                     // pixels the TMA box actually brings in
constexpr int kUpBoxBytes = kUpBoxW * 128;
constexpr int kUpTileBytes = kUpC * 128;         // one (kh, q) weight tile of a 64-channel block
constexpr int kUpGroupBytes = 6 * kUpTileBytes;  // the six tiles of one (channel block, dx)
constexpr int kUpMaxSlots = 12;
constexpr int kUpSbuf = 32 * kUpC * 2;           // staging tile: 32 pixels x 32 channels bf16

struct Up2hParams {
  int batch, groups, in_h, in_w, cin, cout, kc;
  int S;                      // activation ring slots
  int strips, nslices, segs, L;
  long long ldo, co_off;
  const float *row_scale, *noise, *noise_weight_dev, *bias;
  long long noise_bstride;
  float noise_weight;
  int act;
  float alpha, scale;
  const __nv_bfloat16 *res1, *res2;
  float ky[4];                // vertical taps as applied: out[y] = sum_u ky[u] hz[y + u - 1]
};

struct UpUnit {
  int b, strip, slice, a0, L;
};

__device__ __forceinline__ UpUnit up_decode(const Up2hParams &p, long long u) {
  UpUnit r;
  r.slice = (int)(u % p.nslices); u /= p.nslices;
  const int seg = (int)(u % p.segs); u /= p.segs;
  r.strip = (int)(u % p.strips); u /= p.strips;
  r.b = (int)u;
  r.a0 = seg * p.L;
  r.L = min(p.L, p.in_h - r.a0);
  return r;
}

template <int NSET>
__global__ void __launch_bounds__(kUpThreads, 1)
conv_up2h_kernel(const Up2hParams p, const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ OutMaps omaps) {
  // Shared memory (every byte of the 227 KB is spoken for at Cin = 128, so there is no alignment slack: the dynamic
  // region is declared 1024-byte aligned — the swizzle atom — and checked): [S activation slots][resident weights]
  // [staging tiles]; barriers and the per-channel vectors live in the unused tail of activation slot 0 (the TMA box
  // writes 130 of its 136 pixels).
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int S = p.S, kc = p.kc;
  const uint32_t b_total = (uint32_t)(kc * 3) * kUpGroupBytes;
  unsigned char *a_buf = smem;
  unsigned char *b_buf = smem + (size_t)S * kUpSlotBytes;
  unsigned char *o_buf = b_buf + b_total;                               // 8 warps x NSET x 2 tiles
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kUpBoxBytes);
  uint64_t *a_full = bars, *a_empty = bars + kUpMaxSlots;
  uint64_t *b_full = bars + 2 * kUpMaxSlots, *b_empty = b_full + 1;
  uint64_t *acc_full = b_empty + 1, *acc_empty = acc_full + kUpNR;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kUpNR);
  float *epi_vec = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + 384);   // [2][32], ends at byte 640 of 768

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&omaps.m[0]);
    for (int i = 0; i < kUpMaxSlots; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(b_full, 1);
    mbar_init(b_empty, 1);
    for (int i = 0; i < kUpNR; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);    // one arrive per epilogue warp
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

  const long long units = (long long)p.batch * p.strips * p.segs * p.nslices;

  if (warp == 0) {
    // ===================== input rows a0-1 .. a0+L (kc slots per row), each loaded once =====================
    int slot = 0;
    uint32_t ph = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const UpUnit un = up_decode(p, u);
      const int w0 = un.strip * kBlockM - 1;
      for (int k = 0; k < un.L + 2; ++k) {
        const int ih = un.a0 - 1 + k;
        for (int cb = 0; cb < kc; ++cb) {
          mbar_wait(&a_empty[slot], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[slot], kUpBoxBytes);
            tma_load_4d(a_buf + (size_t)slot * kUpSlotBytes, &tmap_a, &a_full[slot], cb * kBlockK, w0, ih, un.b);
          }
          __syncwarp();
          if (++slot == S) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== weights of the slice, resident: [cb][dx][kh][q] tiles of 32 rows x 64 channels =====================
    int cur_key = -1;
    uint32_t res_ph = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const UpUnit un = up_decode(p, u);
      const int g = p.groups == 1 ? 0 : un.b;
      const int key = g * p.nslices + un.slice;
      if (key == cur_key) continue;
      mbar_wait(b_empty, res_ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full, b_total);
        for (int cb = 0; cb < kc; ++cb)
          for (int dx = 0; dx < 3; ++dx)
            for (int kh = 0; kh < 3; ++kh)
              for (int q = 0; q < 2; ++q)
                tma_load_4d(b_buf + (size_t)(cb * 3 + dx) * kUpGroupBytes + (size_t)(kh * 2 + q) * kUpTileBytes, &tmap_b,
                            b_full, cb * kBlockK, q * p.cout + un.slice * kUpC, kh * 3 + dx, g);
      }
      __syncwarp();
      cur_key = key;
      res_ph ^= 1;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc128 = umma_idesc_bf16(kBlockM, 4 * kUpC);
    constexpr uint32_t idesc64 = umma_idesc_bf16(kBlockM, 2 * kUpC);
    constexpr uint32_t idesc192 = umma_idesc_bf16(kBlockM, 6 * kUpC);
    const uint32_t a_base = smem_u32(a_buf), b_base = smem_u32(b_buf);
    int aslot = 0;
    uint32_t aph = 0, res_ph = 0;
    long long acq = 0;                       // TMEM slots acquired so far (ring order)
    long long row0 = 0;                      // running index of the unit's first slot
    int cur_key = -1;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const UpUnit un = up_decode(p, u);
      const int key = (p.groups == 1 ? 0 : un.b) * p.nslices + un.slice;
      if (key != cur_key) {
        mbar_wait(b_full, res_ph);
        res_ph ^= 1;
        cur_key = key;
      }
      bool release_b = u + gridDim.x >= units;
      if (!release_b) {
        const UpUnit nx = up_decode(p, u + gridDim.x);
        release_b = ((p.groups == 1 ? 0 : nx.b) * p.nslices + nx.slice) != key;
      }
      const int rows = un.L + 2;
      for (int k = 0; k < rows; ++k) {
        // slots row0 .. row0 + min(k + 1, rows - 1) must be ours before this row's MMAs write them
        const long long need = row0 + min(k + 1, rows - 1) + 1;
        while (acq < need) {
          mbar_wait(&acc_empty[acq % kUpNR], (uint32_t)(((acq / kUpNR) & 1) ^ 1));
          ++acq;
        }
        const int s0 = aslot;
        for (int cb = 0; cb < kc; ++cb) {
          mbar_wait(&a_full[aslot], aph);
          if (++aslot == S) { aslot = 0; aph ^= 1; }
        }
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t d_own = tmem_base + (uint32_t)(((row0 + k) % kUpNR) * kUpSlotCols);
          const uint32_t d_next = tmem_base + (uint32_t)(((row0 + k + 1) % kUpNR) * kUpSlotCols);
          const bool has_next = k + 1 < rows;
          const bool joined = has_next && (int)((row0 + k) % kUpNR) != kUpNR - 1;   // slot r+1 is the next 128 columns
          int s = s0;
          for (int cb = 0; cb < kc; ++cb) {
            const uint32_t arow = a_base + (uint32_t)s * kUpSlotBytes;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const uint64_t adesc = umma_smem_desc(arow + (uint32_t)dx * 128u, 128);
              const uint32_t grp = b_base + (uint32_t)(cb * 3 + dx) * kUpGroupBytes;
              const uint64_t b01 = umma_smem_desc(grp, 128);                         // T0 | T1 (128 rows)
              const uint64_t b1 = umma_smem_desc(grp + 2 * kUpTileBytes, 128);       // T1 (64 rows)
              const uint64_t b2 = umma_smem_desc(grp + 4 * kUpTileBytes, 128);       // T2 (64 rows)
#pragma unroll
              for (int ks = 0; ks < kBlockK / kUmmaK; ++ks) {
                const bool first = cb == 0 && dx == 0 && ks == 0;
                const uint64_t ko = (uint64_t)(2 * ks);
                // Slot r+1 starts right after slot r in TMEM, so [E_r | O_r | E_{r+1}] = [T0 | T1 | T2] is ONE N = 192
                // MMA (the activation slab is fetched once per k-step) except across the ring wrap; only the very first
                // k-step of a row is split, because E_r accumulates while O_r and E_{r+1} are written for the first time.
                if (k == 0) {                   // unit's first row: every target is a first write
                  if (joined) {
                    umma_bf16_ss(d_own, adesc + ko, b01 + ko, idesc192, first ? 0u : 1u);
                  } else {
                    umma_bf16_ss(d_own, adesc + ko, b01 + ko, idesc128, first ? 0u : 1u);
                    if (has_next) umma_bf16_ss(d_next, adesc + ko, b2 + ko, idesc64, first ? 0u : 1u);
                  }
                } else if (first) {
                  umma_bf16_ss(d_own, adesc + ko, b01 + ko, idesc64, 1u);                     // E[r] += T0[r]
                  if (joined) {
                    umma_bf16_ss(d_own + 2 * kUpC, adesc + ko, b1 + ko, idesc128, 0u);        // O[r] | E[r+1] = T1 | T2
                  } else {
                    umma_bf16_ss(d_own + 2 * kUpC, adesc + ko, b1 + ko, idesc64, 0u);         // O[r]  = T1[r]
                    if (has_next) umma_bf16_ss(d_next, adesc + ko, b2 + ko, idesc64, 0u);     // E[r+1] = T2[r]
                  }
                } else if (joined) {
                  umma_bf16_ss(d_own, adesc + ko, b01 + ko, idesc192, 1u);
                } else {
                  umma_bf16_ss(d_own, adesc + ko, b01 + ko, idesc128, 1u);
                  if (has_next) umma_bf16_ss(d_next, adesc + ko, b2 + ko, idesc64, 1u);
                }
              }
            }
            umma_commit(&a_empty[s]);                  // this row block is consumed exactly once
            if (++s == S) s = 0;
          }
          umma_commit(&acc_full[(row0 + k) % kUpNR]);
          if (k == rows - 1 && release_b) umma_commit(b_empty);
        }
        __syncwarp();
      }
      row0 += rows;
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps): vertical 4-tap filter over the arriving hz rows =====================
    // Step k (slot k = input row a0-1+k has arrived) completes output row pair a = a0+k-2 from
    //   R = ky0 O[a-1] (registers),  E[a], O[a] (slot k-1, second and last read),  E[a+1], O[a+1] (slot k, first read),
    // then R <- ky0 O[a] and slot k-1 is handed back: a slot lives for two steps, the MMA warp runs one row ahead.
    const int q = (warp - 4) >> 2;             // output column parity handled by this warp
    const int quad = warp & 3;
    const int et = threadIdx.x - 128;
    float *vec_rs = epi_vec, *vec_b2 = epi_vec + kUpC;
    const float nw = p.noise ? (p.noise_weight_dev ? __ldg(p.noise_weight_dev) : p.noise_weight) : 0.f;
    const float m2 = p.act ? p.scale : 1.f, a2 = p.act ? p.alpha : 1.f;
    unsigned char *stage = o_buf + (size_t)(warp - 4) * NSET * 2 * kUpSbuf;
    const int sw = (lane >> 1) & 3;
    const float ky0 = p.ky[0], ky1 = p.ky[1], ky2 = p.ky[2], ky3 = p.ky[3];
    const int full_h = 2 * p.in_h, full_w = 2 * p.in_w;
    const bool has_res = p.res1 != nullptr || p.res2 != nullptr;
    const long long rstride = (long long)full_w * p.ldo;       // one output row, in elements
    unsigned long long R2[kUpC / 2];           // ky0 * O[a-1], packed (channel 2i, 2i+1): all epilogue arithmetic is FFMA2
#pragma unroll
    for (int i = 0; i < kUpC / 2; ++i) R2[i] = 0ull;
    const unsigned long long ky0p = f2pack(ky0, ky0), ky1p = f2pack(ky1, ky1), ky2p = f2pack(ky2, ky2), ky3p = f2pack(ky3, ky3);
    const unsigned long long a2p = f2pack(a2, a2);
    uint32_t rr[2][2][16];                      // [residual][row parity][64 bytes of this thread's pixel], one row pair ahead
#pragma unroll
    for (int i = 0; i < 64; ++i) (&rr[0][0][0])[i] = 0u;
    long long rd = 0;                           // slots seen so far
    uint32_t set = 0;
    int cur_key = -1;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const UpUnit un = up_decode(p, u);
      const int c0 = un.slice * kUpC;
      const int key = un.b * p.nslices + un.slice;
      if (key != cur_key) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = et; c < kUpC; c += 256) {
          const float rs = p.row_scale ? __ldg(p.row_scale + (long long)un.b * p.cout + c0 + c) : 1.f;
          const float b2 = p.bias ? __ldg(p.bias + c0 + c) : 0.f;
          vec_rs[c] = rs * m2;                  // lrelu(t) * s == max(t s, t s a): gains folded into the vectors
          vec_b2[c] = b2 * m2;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_key = key;
      }
      const int j = un.strip * kBlockM + quad * 32 + lane;       // low-resolution column of this thread
      const bool pix_ok = j < p.in_w;
      const int fw = 2 * j + q;
      const int rows = un.L + 2;
      const long long pix0 = (((long long)un.b * full_h) * full_w + fw) * p.ldo + p.co_off + c0;   // row 0 of this thread's column
      auto load_noise = [&](int a, int pp) -> float {
        if (p.noise == nullptr || !pix_ok || a >= un.a0 + un.L) return 0.f;
        return __ldg(p.noise + un.b * p.noise_bstride + (long long)(2 * a + pp) * full_w + fw);
      };
      // residuals of output row pair `a`: the thread's own pixel, 64 contiguous bytes per row as two 256-bit loads
      auto load_res = [&](int a) {
        if (!has_res || !pix_ok || a >= un.a0 + un.L) return;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const __nv_bfloat16 *base = w ? p.res2 : p.res1;
          if (base == nullptr) continue;
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            const __nv_bfloat16 *src = base + pix0 + (long long)(2 * a + pp) * rstride;
            ldg256(src, &rr[w][pp][0]);
            ldg256(src + 16, &rr[w][pp][8]);
          }
        }
        if (a + 1 < un.a0 + un.L) {              // and the pair after it towards L2
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            if (p.res1) prefetch_l2(p.res1 + pix0 + (long long)(2 * a + 2 + pp) * rstride);
            if (p.res2) prefetch_l2(p.res2 + pix0 + (long long)(2 * a + 2 + pp) * rstride);
          }
        }
      };
      float nz0 = load_noise(un.a0, 0), nz1 = load_noise(un.a0, 1);
      for (int k = 0; k < rows; ++k, ++rd) {
        if (k == 0) {                            // slot 0 (row a0-1) is read together with slot 1
          load_res(un.a0);
          continue;
        }
        const int slot = (int)(rd % kUpNR), prev = (int)((rd + kUpNR - 1) % kUpNR);
        const bool emit = k >= 2;
        const int a = un.a0 + k - 2;             // output row pair completed by this step
        const float n0 = nw * nz0 * m2, n1 = nw * nz1 * m2;
        unsigned char *buf0 = stage + (size_t)((NSET == 2 ? (set & 1) : 0) * 2) * kUpSbuf, *buf1 = buf0 + kUpSbuf;
        if (emit) {
          nz0 = load_noise(a + 1, 0);
          nz1 = load_noise(a + 1, 1);
          ++set;
          if (lane == 0) bulk_wait_group_read<NSET - 1>();
          __syncwarp();
        }
        mbar_wait(&acc_full[slot], (uint32_t)((rd / kUpNR) & 1));        // MMAs retire in order: slot k-1 is complete too
        tcgen05_fence_after();
        const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
        const uint32_t tp = tmem_base + lane_sel + (uint32_t)(prev * kUpSlotCols + q * kUpC);
        const uint32_t tn = tmem_base + lane_sel + (uint32_t)(slot * kUpSlotCols + q * kUpC);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint32_t e[8], o[8], en[8], on[8];
          tmem_ld_32x32b_x8(tp + 2 * kUpC + h * 8, o);
          if (emit) {
            tmem_ld_32x32b_x8(tp + h * 8, e);
            tmem_ld_32x32b_x8(tn + h * 8, en);
            tmem_ld_32x32b_x8(tn + 2 * kUpC + h * 8, on);
          }
          tmem_ld_wait();
          if (h == 3) {                          // hand slot k-1 (and the unit's last slot) back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&acc_empty[prev]);
              if (k == rows - 1) mbar_arrive(&acc_empty[slot]);
            }
          }
          unsigned long long o2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) o2[i] = f2pack(__uint_as_float(o[2 * i]), __uint_as_float(o[2 * i + 1]));
          if (emit) {
            const unsigned long long n0p = f2pack(n0, n0), n1p = f2pack(n1, n1);
            float v0[8], v1[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c2 = h * 4 + i;                                        // channel pair index
              const unsigned long long e2 = f2pack(__uint_as_float(e[2 * i]), __uint_as_float(e[2 * i + 1]));
              const unsigned long long en2 = f2pack(__uint_as_float(en[2 * i]), __uint_as_float(en[2 * i + 1]));
              const unsigned long long on2 = f2pack(__uint_as_float(on[2 * i]), __uint_as_float(on[2 * i + 1]));
              const unsigned long long f0 = f2fma(ky3p, en2, f2fma(ky2p, o2[i], f2fma(ky1p, e2, R2[c2])));
              const unsigned long long f1 = f2fma(ky3p, on2, f2fma(ky2p, en2, f2fma(ky1p, o2[i], f2mul(ky0p, e2))));
              const unsigned long long rs2 = reinterpret_cast<const unsigned long long *>(vec_rs)[c2];
              const unsigned long long b22 = reinterpret_cast<const unsigned long long *>(vec_b2)[c2];
              const unsigned long long t0 = f2fma(f0, rs2, f2add(b22, n0p));
              const unsigned long long t1 = f2fma(f1, rs2, f2add(b22, n1p));
              const unsigned long long s0 = f2mul(t0, a2p), s1 = f2mul(t1, a2p);
              float t0l, t0h, t1l, t1h, s0l, s0h, s1l, s1h;
              f2unpack(t0, t0l, t0h); f2unpack(t1, t1l, t1h); f2unpack(s0, s0l, s0h); f2unpack(s1, s1l, s1h);
              v0[2 * i] = fmaxf(t0l, s0l); v0[2 * i + 1] = fmaxf(t0h, s0h);
              v1[2 * i] = fmaxf(t1l, s1l); v1[2 * i + 1] = fmaxf(t1h, s1h);
            }
            if (has_res) {
              const uint4 a0v = make_uint4(rr[0][0][4 * h], rr[0][0][4 * h + 1], rr[0][0][4 * h + 2], rr[0][0][4 * h + 3]);
              const uint4 b0v = make_uint4(rr[1][0][4 * h], rr[1][0][4 * h + 1], rr[1][0][4 * h + 2], rr[1][0][4 * h + 3]);
              const uint4 a1v = make_uint4(rr[0][1][4 * h], rr[0][1][4 * h + 1], rr[0][1][4 * h + 2], rr[0][1][4 * h + 3]);
              const uint4 b1v = make_uint4(rr[1][1][4 * h], rr[1][1][4 * h + 1], rr[1][1][4 * h + 2], rr[1][1][4 * h + 3]);
              add2_bf16x8(v0, a0v, b0v);
              add2_bf16x8(v1, a1v, b1v);
            }
            *reinterpret_cast<uint4 *>(buf0 + lane * 64 + ((h ^ sw) << 4)) = pack8_bf16(v0);
            *reinterpret_cast<uint4 *>(buf1 + lane * 64 + ((h ^ sw) << 4)) = pack8_bf16(v1);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) R2[h * 4 + i] = f2mul(ky0p, o2[i]);
        }
        if (emit) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const int cx = un.strip * kBlockM + quad * 32;
            tma_store_4d(&omaps.m[q], buf0, (int)p.co_off + c0, cx, a, un.b);        // row 2a,   columns 2j+q
            tma_store_4d(&omaps.m[2 + q], buf1, (int)p.co_off + c0, cx, a, un.b);    // row 2a+1
            bulk_commit_group();
          }
          load_res(a + 1);                       // next pair's residuals: in flight during the next accumulator wait
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NSET>
int launch_up2h(const Up2hParams &p, const void *x, const void *wq, void *out, size_t smem_bytes, cudaStream_t stream) {
  auto kern = conv_up2h_kernel<NSET>;
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
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)p.in_w, (uint64_t)p.in_h, (uint64_t)p.batch};
    uint64_t strides[4] = {0, (uint64_t)p.cin * 2, (uint64_t)p.cin * p.in_w * 2, (uint64_t)p.cin * p.in_w * p.in_h * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)kUpBoxW, 1, 1};
    if (int rc = encode_tma(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)(2 * p.cout), 9, (uint64_t)p.groups};
    uint64_t strides[4] = {0, (uint64_t)p.cin * 2, (uint64_t)p.cin * 2 * p.cout * 2, (uint64_t)p.cin * 2 * p.cout * 9 * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)kUpC, 1, 1};
    if (int rc = encode_tma(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, wq, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  const long long full_h = 2LL * p.in_h, full_w = 2LL * p.in_w;
  for (int cls = 0; cls < 4; ++cls) {       // class = (row parity, column parity): origin in the base, pixel strides x 2
    const int oh0 = cls >> 1, ow0 = cls & 1;
    const __nv_bfloat16 *base = static_cast<const __nv_bfloat16 *>(out) + ((long long)oh0 * full_w + ow0) * p.ldo;
    uint64_t dims[4] = {(uint64_t)(p.co_off + p.cout), (uint64_t)p.in_w, (uint64_t)p.in_h, (uint64_t)p.batch};
    uint64_t strides[4] = {0, (uint64_t)p.ldo * 4, (uint64_t)p.ldo * full_w * 4, (uint64_t)p.ldo * full_w * full_h * 2};
    uint32_t box[4] = {(uint32_t)kUpC, 32, 1, 1};
    if (int rc = encode_tma(&om.m[cls], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_64B))
      return rc;
  }
  const long long units = (long long)p.batch * p.strips * p.segs * p.nslices;
  long long grid = (num_sms() / p.nslices) * p.nslices;      // a CTA keeps its slice: the weights stay resident
  if (grid < p.nslices) grid = p.nslices;
  if (units < grid) grid = units;
  kern<<<(unsigned)grid, kUpThreads, smem_bytes, stream>>>(p, ta, tb, om);
  return check_launch("conv_up2h_kernel");
}

}  // namespace
}  // namespace vsp

extern "C" int vsp_conv2d_up2h_bf16(const void *x, const void *wq, void *out, int64_t batch, int64_t groups,
                                    int64_t in_h, int64_t in_w, int64_t cin, int64_t cout, int64_t ldo, int64_t co_off,
                                    const float *ky_host, const vsp_conv_epilogue *epi, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(x && wq && out && ky_host, "conv2d_up2h: null pointer");
  VSP_REQUIRE(batch >= 1 && (groups == 1 || groups == batch), "conv2d_up2h: groups must be 1 or batch");
  VSP_REQUIRE(cin >= 64 && cin <= 128 && cin % 64 == 0, "conv2d_up2h: Cin must be 64 or 128, got %lld", (long long)cin);
  VSP_REQUIRE(cout >= 32 && cout % 32 == 0, "conv2d_up2h: Cout must be a multiple of 32");
  VSP_REQUIRE(in_w >= 32 && in_w % 32 == 0 && in_h >= 1, "conv2d_up2h: width must be a multiple of 32");
  VSP_REQUIRE(batch < 65536 && in_h < 32768 && in_w < 32768, "conv2d_up2h: extent too large");
  VSP_REQUIRE(ldo % 8 == 0 && co_off % 8 == 0 && ldo >= co_off + cout, "conv2d_up2h: output channels must be 16-byte aligned");
  VSP_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wq) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "conv2d_up2h: operands must be 16-byte aligned");
  Up2hParams p;
  memset(&p, 0, sizeof(p));
  p.batch = (int)batch; p.groups = (int)groups; p.in_h = (int)in_h; p.in_w = (int)in_w; p.cin = (int)cin; p.cout = (int)cout;
  p.kc = (int)cin / kBlockK;
  p.ldo = ldo; p.co_off = co_off;
  for (int i = 0; i < 4; ++i) p.ky[i] = ky_host[i];
  p.scale = 1.f; p.alpha = 1.f;
  if (epi) {
    VSP_REQUIRE(epi->pre_act == 0 && epi->pre_bias == nullptr, "conv2d_up2h: no first activation stage");
    VSP_REQUIRE(epi->act == 0 || epi->act == 3, "conv2d_up2h: epilogue act must be 0 or 3");
    VSP_REQUIRE(epi->act == 0 || (epi->alpha >= 0.f && epi->alpha <= 1.f && epi->scale > 0.f),
                "conv2d_up2h: leaky ReLU needs 0 <= alpha <= 1 and scale > 0");
    p.row_scale = epi->row_scale; p.noise = epi->noise; p.noise_bstride = epi->noise_bstride;
    p.noise_weight = epi->noise_weight; p.noise_weight_dev = epi->noise_weight_dev; p.bias = epi->bias;
    p.act = epi->act; p.alpha = epi->alpha; p.scale = epi->scale;
    p.res1 = static_cast<const __nv_bfloat16 *>(epi->residual);
    p.res2 = static_cast<const __nv_bfloat16 *>(epi->residual2);
    if (p.res1 || p.res2)      // the residual path reads each pixel's 64 bytes as two 256-bit loads
      VSP_REQUIRE(ldo % 16 == 0 && co_off % 16 == 0 &&
                      ((reinterpret_cast<uintptr_t>(p.res1) | reinterpret_cast<uintptr_t>(p.res2)) & 31) == 0,
                  "conv2d_up2h: residuals need 32-byte aligned pixels (ldo, co_off multiples of 16)");
  }
  p.strips = (int)((in_w + kBlockM - 1) / kBlockM);
  p.nslices = (int)cout / kUpC;
  // shared memory: resident weights, staging tiles (double-buffered when they fit), the rest is the input-row ring
  const int b_bytes = p.kc * 3 * kUpGroupBytes;
  int nset = 2;
  int S = (232448 - b_bytes - 8 * nset * 2 * kUpSbuf) / kUpSlotBytes;
  if (S < 2 * p.kc + 1) {
    nset = 1;
    S = (232448 - b_bytes - 8 * nset * 2 * kUpSbuf) / kUpSlotBytes;
  }
  if (S > kUpMaxSlots) S = kUpMaxSlots;
  VSP_REQUIRE(S >= p.kc + 1, "conv2d_up2h: shared memory budget");
  p.S = S;
  // row segments: minimise waves x (rows per unit + fill/drain)
  {
    const long long base_units = (long long)p.batch * p.strips * p.nslices;
    const long long cap = (long long)(num_sms() / p.nslices > 0 ? (num_sms() / p.nslices) * p.nslices : p.nslices);
    double best = 1e300;
    int best_L = p.in_h;
    for (int segs = 1; segs <= p.in_h; ++segs) {
      const int L = (p.in_h + segs - 1) / segs;
      const int se = (p.in_h + L - 1) / L;
      const long long un = base_units * se;
      const long long waves = (un + cap - 1) / cap;
      const double cost = (double)waves * (L + 2 + 6);
      if (cost < best) { best = cost; best_L = L; }
      if (L <= 4) break;
    }
    p.L = best_L;
    p.segs = (p.in_h + best_L - 1) / best_L;
  }
  const size_t smem_bytes = (size_t)S * kUpSlotBytes + b_bytes + (size_t)8 * nset * 2 * kUpSbuf;
  return nset == 2 ? launch_up2h<2>(p, x, wq, out, smem_bytes, stream) : launch_up2h<1>(p, x, wq, out, smem_bytes, stream);
}
