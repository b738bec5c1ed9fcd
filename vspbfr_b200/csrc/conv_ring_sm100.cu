// conv_ring_sm100.cu — "row-ring" tcgen05 convolution for the wide, shallow layers (Cin <= 128, Cout <= 128,
// width >= 128: the 256^2 .. 1024^2 levels of both networks, SURVEY.md Appendix A "ridge point").
//
// Those layers are bandwidth- and latency-bound, not tensor-bound: with the plain implicit GEMM
// (conv_sm100.cu) every 128-pixel tile re-reads its input once per tap and pays a full epilogue round trip
// for 128 x N outputs.  Here a CTA owns a vertical strip (128 output columns) of one sample and walks down
// it: every INPUT ROW (128 + 2*dil pixels x 64 channels, one TMA box) is loaded exactly once into a ring
// of row slots and serves the three output rows that need it (the nine taps are row-/column-shifted UMMA
// descriptors into the ring, as in the row-halo kernel); dilated convs walk the strip as `dil` interleaved
// chains (rows c, c+d, c+2d, ...) so the same rolling window applies.  R output rows are accumulated per
// TMEM hand-off (R x N columns per stage, two stages), which amortises the epilogue latency and lets one
// weight tile feed R MMAs.  Weights stay resident in shared memory when they fit (9*kc*N*128 B), else they
// stream through their own ring from a dedicated producer warp.  Channel blocks shorter than 64 skip the
// zero-padded K steps.  1x1 convolutions (TAPS = 1) use the same machinery without halo rows.
//
// Warp roles (12 warps): 0 = activation-row producer, 1 = MMA issuer (+ TMEM alloc), 2 = weight producer,
// 3 = idle, 4..11 = epilogue (two warps per TMEM lane quadrant, alternating accumulator chunks).
#include "conv_common.cuh"

#include <stdlib.h>

namespace vsp {
namespace {

constexpr int kRingThreads = 384;
constexpr int kMaxSlots = 16;
constexpr int kMaxBSlots = 8;

struct RingUnit {
  int b, n_i, strip, o0, L;
};

// unit -> (sample, channel tile, strip, chain, segment); the host guarantees every unit has L >= 1 rows.
__device__ __forceinline__ RingUnit ring_decode(const ConvParams &p, long long u, int d_eff) {
  RingUnit r;
  const int seg = (int)(u % p.rr_segs); u /= p.rr_segs;
  const int chain = (int)(u % p.rr_chains); u /= p.rr_chains;
  r.strip = (int)(u % p.rr_strips); u /= p.rr_strips;
  r.n_i = (int)(u % p.tiles_n); u /= p.tiles_n;
  r.b = (int)u;
  const int rows_in_chain = (p.out_h - chain + d_eff - 1) / d_eff;
  const int start = seg * p.rr_L;
  r.L = min(p.rr_L, rows_in_chain - start);
  r.o0 = chain + start * d_eff;
  return r;
}

template <int BLOCK_N, int TAPS, bool RESIDENT_B>
__global__ void __launch_bounds__(kRingThreads, 1)
conv_ring_kernel(const ConvParams p, const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b) {
  constexpr int RMAX = BLOCK_N <= 64 ? 4 : (BLOCK_N == 128 ? 2 : 1);
  constexpr int HALO = TAPS == 9 ? 1 : 0;
  constexpr int B_BYTES = BLOCK_N * 128;
  constexpr int CHUNK = BLOCK_N < 32 ? BLOCK_N : 32;
  constexpr int NCH = BLOCK_N / CHUNK;
  constexpr int ACC_COLS = RMAX * BLOCK_N;
  constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;

  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.rr_S, R = p.rr_R, kc = p.kc;
  const int d = TAPS == 9 ? p.halo_d : 1;       // tap spacing == chain stride
  const uint32_t slot_bytes = (uint32_t)p.halo_w * 128u;
  const uint32_t b_total = RESIDENT_B ? (uint32_t)(TAPS * kc) * B_BYTES : (uint32_t)p.rr_nb * B_BYTES;
  unsigned char *a_buf = smem;
  unsigned char *b_buf = smem + (size_t)S * slot_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(b_buf + b_total);
  uint64_t *a_full = bars, *a_empty = bars + kMaxSlots;
  uint64_t *b_full = bars + 2 * kMaxSlots, *b_empty = b_full + kMaxBSlots;
  uint64_t *tmem_full = b_empty + kMaxBSlots, *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  float *epi_vec = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + 512);  // [3][BLOCK_N]

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kMaxBSlots; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
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

  const long long units = (long long)p.batch * p.tiles_n * p.rr_strips * p.rr_chains * p.rr_segs;

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
    // ===================== weights =====================
    if (RESIDENT_B) {
      long long cur_key = -1;
      uint32_t res_ph = 0;
      for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const RingUnit un = ring_decode(p, u, d);
        const int g = p.groups == 1 ? 0 : un.b;
        const long long key = (long long)g * p.tiles_n + un.n_i;
        if (key == cur_key) continue;
        mbar_wait(&b_empty[0], res_ph ^ 1);          // every MMA that read the old weights has retired
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[0], b_total);
          for (int cb = 0; cb < kc; ++cb)
            for (int tap = 0; tap < TAPS; ++tap)
              tma_load_4d(b_buf + (size_t)(cb * TAPS + tap) * B_BYTES, &tmap_b, &b_full[0], cb * kBlockK,
                          un.n_i * BLOCK_N, p.tap_w[tap], g);
        }
        __syncwarp();
        cur_key = key;
        res_ph ^= 1;
      }
    } else {
      int bs = 0;
      uint32_t bph = 0;
      for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const RingUnit un = ring_decode(p, u, d);
        const int g = p.groups == 1 ? 0 : un.b;
        const int nsteps = (un.L + R - 1) / R;
        for (int st = 0; st < nsteps; ++st)
          for (int cb = 0; cb < kc; ++cb)
            for (int tap = 0; tap < TAPS; ++tap) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[bs], B_BYTES);
                tma_load_4d(b_buf + (size_t)bs * B_BYTES, &tmap_b, &b_full[bs], cb * kBlockK, un.n_i * BLOCK_N,
                            p.tap_w[tap], g);
              }
              __syncwarp();
              if (++bs == p.rr_nb) { bs = 0; bph ^= 1; }
            }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
    const uint32_t a_base = smem_u32(a_buf), b_base0 = smem_u32(b_buf);
    int wslot = 0, base = 0, acc = 0, bs = 0;
    uint32_t wph = 0, acc_phase = 0, bph = 0, res_ph = 0;
    long long cur_key = -1;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      bool release_b = false;
      if (RESIDENT_B) {
        const long long key = (long long)(p.groups == 1 ? 0 : un.b) * p.tiles_n + un.n_i;
        if (key != cur_key) {
          mbar_wait(&b_full[0], res_ph);
          res_ph ^= 1;
          cur_key = key;
        }
        const long long nu = u + gridDim.x;
        release_b = nu >= units;
        if (!release_b) {
          const RingUnit nx = ring_decode(p, nu, d);
          release_b = ((long long)(p.groups == 1 ? 0 : nx.b) * p.tiles_n + nx.n_i) != cur_key;
        }
      }
      const int nsteps = (un.L + R - 1) / R;
      int rows_waited = 0;
      for (int st = 0; st < nsteps; ++st) {
        const int reff = min(R, un.L - st * R);
        const bool last = st == nsteps - 1;
        const int need = st * R + reff + 2 * HALO;
        while (rows_waited < need) {
          for (int cb = 0; cb < kc; ++cb) {
            mbar_wait(&a_full[wslot], wph);
            if (++wslot == S) { wslot = 0; wph ^= 1; }
          }
          ++rows_waited;
        }
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        for (int cb = 0; cb < kc; ++cb) {
          const int kk = min(kBlockK / kUmmaK, (p.cin - cb * kBlockK + kUmmaK - 1) / kUmmaK);   // skip zero-padded K
          uint32_t ra[RMAX + 2];
#pragma unroll
          for (int j = 0; j < RMAX + 2; ++j) {
            int s = base + j * kc + cb;
            if (s >= S) s -= S;
            if (s >= S) s -= S;
            ra[j] = a_base + (uint32_t)s * slot_bytes;
          }
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap) {
            const int kh = TAPS == 9 ? tap / 3 : 0, kw = TAPS == 9 ? tap % 3 : 0;
            uint32_t bb;
            if (RESIDENT_B) {
              bb = b_base0 + (uint32_t)(cb * TAPS + tap) * B_BYTES;
            } else {
              mbar_wait(&b_full[bs], bph);
              tcgen05_fence_after();
              bb = b_base0 + (uint32_t)bs * B_BYTES;
            }
            if (elect_one()) {
              const uint64_t bdesc = umma_smem_desc(bb, 128);
#pragma unroll
              for (int rr = 0; rr < RMAX; ++rr) {
                if (rr < reff) {
                  const uint64_t adesc = umma_smem_desc(ra[rr + kh] + (uint32_t)(kw * d) * 128u, 128);
#pragma unroll
                  for (int k = 0; k < kBlockK / kUmmaK; ++k)
                    if (k < kk)
                      umma_bf16_ss(d_tmem + (uint32_t)(rr * BLOCK_N), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                                   idesc, (cb > 0 || tap > 0 || k > 0) ? 1u : 0u);
                }
              }
              if (!RESIDENT_B) umma_commit(&b_empty[bs]);
            }
            __syncwarp();
            if (!RESIDENT_B) {
              if (++bs == p.rr_nb) { bs = 0; bph ^= 1; }
            }
          }
        }
        // hand the accumulators over and release the rows no later step needs
        const int nrel = (reff + (last ? 2 * HALO : 0)) * kc;
        if (elect_one()) {
          umma_commit(&tmem_full[acc]);
          int s = base;
          for (int i = 0; i < nrel; ++i) {
            umma_commit(&a_empty[s]);
            if (++s == S) s = 0;
          }
          if (RESIDENT_B && last && release_b) umma_commit(&b_empty[0]);
        }
        __syncwarp();
        base += nrel;
        while (base >= S) base -= S;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
    int acc = 0;
    uint32_t acc_phase = 0;
    long long cur_key = -1;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const RingUnit un = ring_decode(p, u, d);
      const long long key = (long long)un.b * p.tiles_n + un.n_i;
      const int nbase = un.n_i * BLOCK_N;
      if (key != cur_key) {
        // per-channel vectors change only with (sample, channel tile): restage them once per run of units
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = et; c < BLOCK_N; c += 256) {
          const int n = nbase + c;
          const bool ok = n < p.cout;
          vec_rs[c] = (ok && p.row_scale) ? __ldg(p.row_scale + (long long)un.b * p.cout + n) : 1.f;
          vec_b1[c] = (ok && p.pre_bias) ? __ldg(p.pre_bias + n) : 0.f;
          vec_b2[c] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_key = key;
      }
      const int ow = un.strip * kBlockM + row;
      const bool pix_ok = ow < p.out_w;
      const int fw = ow * p.os + p.oo_w;
      const int nsteps = (un.L + R - 1) / R;
      for (int st = 0; st < nsteps; ++st) {
        const int reff = min(R, un.L - st * R);
        // noise of every row of this step is fetched before waiting for the accumulators
        float nz[RMAX];
        const int oh0 = un.o0 + st * R * d;
#pragma unroll
        for (int rr = 0; rr < RMAX; ++rr) {
          nz[rr] = 0.f;
          if (rr < reff && p.noise != nullptr && pix_ok)
            nz[rr] = nw * __ldg(p.noise + un.b * p.noise_bstride + (long long)((oh0 + rr * d) * p.os + p.oo_h) * p.full_w + fw);
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * ACC_COLS);
        // (row, chunk) items alternate between the two epilogue warps of this lane quadrant
#pragma unroll 1
        for (int it = wg; it < reff * NCH; it += 2) {
          const int rr = it / NCH, ch = it % NCH;
          float nzv = nz[0];
#pragma unroll
          for (int q = 1; q < RMAX; ++q) nzv = rr == q ? nz[q] : nzv;
          const long long pix = (long long)((oh0 + rr * d) * p.os + p.oo_h) * p.full_w + fw;
          epi_chunk<CHUNK>(p, taddr + rr * BLOCK_N + ch * CHUNK, nbase + ch * CHUNK, un.b, pix, plane, pix_ok, nzv,
                           vec_rs + ch * CHUNK, vec_b1 + ch * CHUNK, vec_b2 + ch * CHUNK);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int TAPS, bool RESIDENT_B>
int launch_ring(const ConvParams &p, const void *x, const void *wq, int64_t in_h, int64_t in_w, int64_t cout_pad,
                int taps_total, size_t smem_bytes, cudaStream_t stream) {
  auto kern = conv_ring_kernel<BLOCK_N, TAPS, RESIDENT_B>;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  CUtensorMap ta, tb;
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
  const long long units = (long long)p.batch * p.tiles_n * p.rr_strips * p.rr_chains * p.rr_segs;
  const long long grid = units < num_sms() ? units : num_sms();
  kern<<<(unsigned)grid, kRingThreads, smem_bytes, stream>>>(p, ta, tb);
  return check_launch("conv_ring_kernel");
}

template <int BLOCK_N>
int dispatch_ring(const ConvParams &p, int taps, bool resident, const void *x, const void *wq, int64_t in_h,
                  int64_t in_w, int64_t cout_pad, int taps_total, size_t smem_bytes, cudaStream_t stream) {
  if (taps == 9)
    return resident ? launch_ring<BLOCK_N, 9, true>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream)
                    : launch_ring<BLOCK_N, 9, false>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
  return resident ? launch_ring<BLOCK_N, 1, true>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream)
                  : launch_ring<BLOCK_N, 1, false>(p, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
}

}  // namespace

int conv_ring_try_launch(ConvParams &p, const void *x, const void *wq, int64_t in_h, int64_t in_w,
                         int64_t cout_pad, int taps_total, int dil, cudaStream_t stream) {
  static const bool disabled = getenv("VSP_NO_RING") != nullptr;
  if (disabled) return -1;
  if (p.stride != 1 || p.os != 1 || p.oo_h != 0 || p.oo_w != 0) return -1;
  if (p.out_w < kBlockM || p.out_w != in_w || p.out_h != in_h) return -1;
  if (p.ntaps != 9 && p.ntaps != 1) return -1;
  if (p.kc > 2 || p.cout > 128) return -1;
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
  const int rmax = bn <= 64 ? 4 : 2;
  p.halo_d = d;
  p.halo_w = taps == 9 ? ((kBlockM + 2 * d + 7) & ~7) : kBlockM;
  const int slot = p.halo_w * 128;
  const int budget = 232448 - 1024 - 512 - 3 * bn * 4;
  int nb = 49152 / (bn * 128);
  nb = nb < 2 ? 2 : (nb > kMaxBSlots ? kMaxBSlots : nb);
  // resident weights first (R >= 2), then streamed weights, then resident with R = 1
  int best_R = 0, best_S = 0;
  bool resident = false;
  for (int pass = 0; pass < 3 && !best_R; ++pass) {
    const bool res = pass != 1;
    const int bbytes = res ? taps * p.kc * bn * 128 : nb * bn * 128;
    if (bbytes >= budget) continue;
    int S = (budget - bbytes) / slot;
    if (S > kMaxSlots) S = kMaxSlots;
    for (int R = (pass == 2 ? 1 : rmax); R >= (pass == 0 ? 2 : 1); --R) {
      const int live = (R + 2 * halo) * p.kc;
      const int pre = R * p.kc / 2 > p.kc ? R * p.kc / 2 : p.kc;
      if (S >= live + pre) { best_R = R; best_S = S; resident = res; break; }
    }
  }
  if (!best_R) return -1;
  p.rr_R = best_R; p.rr_S = best_S; p.rr_nb = nb;
  p.tiles_n = 1;
  p.rr_strips = (p.out_w + kBlockM - 1) / kBlockM;
  p.rr_chains = d;
  // segments: minimise waves x rows per unit
  const int rows_chain = p.out_h / d;
  const long long base_units = (long long)p.batch * p.rr_strips * p.rr_chains;
  double best = 1e300;
  int best_L = rows_chain;
  for (int segs = 1; segs <= rows_chain; ++segs) {
    int L = (rows_chain + segs - 1) / segs;
    L = (L + best_R - 1) / best_R * best_R;
    const int se = (rows_chain + L - 1) / L;
    const long long units = base_units * se;
    const long long waves = (units + num_sms() - 1) / num_sms();
    const double cost = (double)waves * (L + 2 * halo + 3);
    if (cost < best) { best = cost; best_L = L; }
    if (L <= best_R) break;
  }
  p.rr_L = best_L;
  p.rr_segs = (rows_chain + best_L - 1) / best_L;
  const size_t smem_bytes = 1024 + (size_t)best_S * slot + (resident ? taps * p.kc * bn * 128 : nb * bn * 128) + 512 +
                            3 * bn * 4;
  switch (bn) {
    case 16: return dispatch_ring<16>(p, taps, resident, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
    case 32: return dispatch_ring<32>(p, taps, resident, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
    case 64: return dispatch_ring<64>(p, taps, resident, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
    default: return dispatch_ring<128>(p, taps, resident, x, wq, in_h, in_w, cout_pad, taps_total, smem_bytes, stream);
  }
}

}  // namespace vsp
