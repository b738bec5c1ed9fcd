// wgrad_sm100.cu — convolution weight gradient as a bf16 GEMM on tcgen05 / TMEM (K = pixels).
//
// What it replaces: aten::cudnn_convolution_backward_weight, which the reference calls from
// Conv2dGradWeight (op/conv2d_gradfix.py:177-199) for both the per-sample (groups = B) modulated
// convolution and the plain EqualConv2d.
//
//   gw[g, t, o, i] = sum_{b in group g} sum_{oh, ow} dy[b, oh, ow, o] * x[b, oh*s + dy_t, ow*s + dx_t, i]
//
// GEMM view per (g, tap): M = 128 output channels, N = BLOCK_N input channels, K = pixels walked
// in 64-pixel patches (khb x kwb).  Both operands are channels-last bf16, so channels (M resp. N)
// are the contiguous axis: the smem tiles are [64 pixels][64 channels] 128B-swizzled TMA boxes and
// the UMMA descriptors are MN-major (a_major = b_major = 1; LBO = distance between 64-channel
// boxes, SBO = 1024 B between 8-pixel groups).  The tap shift lives in the spatial TMA coordinates
// of the x box (any alignment is legal there; only the innermost = channel coordinate must be
// 16-byte aligned), so zero padding is again the TMA out-of-bounds fill.  Split-K over pixel blocks
// (fp32 atomics into a zeroed output) keeps all SMs busy when groups*taps*tiles is small.
#include "common.cuh"

namespace vsp {
namespace {

constexpr int kBlockM = 128;
constexpr int kPixK = 64;                 // pixels per k-block
constexpr int kBoxBytes = 64 * 64 * 2;    // one [64 px][64 ch] bf16 box = 8 KB
constexpr int kNumThreads = 192;

struct WgradParams {
  int batch, groups;
  int cin, cout;
  int out_h, out_w;       // dy extent
  int stride;
  int ntaps;
  int tap_dy[16], tap_dx[16];
  int kwb, khb, kb_w, kb_h;   // pixel patch and number of patches per sample
  int tiles_m, tiles_n, ksplit;
  long long total_tiles;
  float *gw;              // [groups, taps, cout, cin]
};

template <int BLOCK_N>
struct WCfg {
  static constexpr int A_BYTES = 2 * kBoxBytes;
  static constexpr int B_BYTES = (BLOCK_N / 64) * kBoxBytes;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N >= 256) ? 4 : ((BLOCK_N >= 128) ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// MN-major, 128B-swizzled operand: see cute/atom/mma_traits_sm100.hpp (make_umma_desc<Major::MN>)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kBoxBytes >> 4) << 16;   // LBO: next 64-channel box
  d |= static_cast<uint64_t>(1024 >> 4) << 32;        // SBO: next group of 8 pixels (8 rows x 128 B)
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;                // SWIZZLE_128B
  return d;
}

__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kNumThreads, 1)
wgrad_kernel(const WgradParams p, const __grid_constant__ CUtensorMap tmap_dy,
             const __grid_constant__ CUtensorMap tmap_x) {
  using C = WCfg<BLOCK_N>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t *empty_bar = full_bar + C::STAGES;
  uint64_t *tmem_full = empty_bar + C::STAGES;
  uint64_t *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
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

  // k-blocks of one tile: samples of the group x pixel patches, divided among ksplit slices
  const int samples = p.groups == 1 ? p.batch : 1;
  const int kb_total = samples * p.kb_h * p.kb_w;
  const int kb_per = (kb_total + p.ksplit - 1) / p.ksplit;

  if (warp == 0) {
    // whole warp walks the loop; one elected lane issues (see elect_one in common.cuh)
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long t = tile;
      const int ks = (int)(t % p.ksplit); t /= p.ksplit;
      const int n_i = (int)(t % p.tiles_n); t /= p.tiles_n;
      const int m_i = (int)(t % p.tiles_m); t /= p.tiles_m;
      const int tap = (int)(t % p.ntaps); t /= p.ntaps;
      const int g = (int)t;
      const int kb0 = ks * kb_per, kb1 = min(kb0 + kb_per, kb_total);
      for (int kb = kb0; kb < kb1; ++kb) {
        int r = kb;
        const int wb = r % p.kb_w; r /= p.kb_w;
        const int hb = r % p.kb_h; r /= p.kb_h;
        const int b = p.groups == 1 ? r : g;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          unsigned char *sa = smem + stage * C::STAGE_BYTES;
          unsigned char *sb = sa + C::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          const int ow0 = wb * p.kwb, oh0 = hb * p.khb;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(sa + j * kBoxBytes, &tmap_dy, &full_bar[stage], m_i * kBlockM + 64 * j, ow0, oh0, b);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_4d(sb + j * kBoxBytes, &tmap_x, &full_bar[stage], n_i * BLOCK_N + 64 * j,
                        ow0 * p.stride + p.tap_dx[tap], oh0 * p.stride + p.tap_dy[tap], b);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16_mn(kBlockM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int ks = (int)(tile % p.ksplit);
      const int kb0 = ks * kb_per, kb1 = min(kb0 + kb_per, kb_total);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = umma_desc_mn(sa);
          const uint64_t bdesc = umma_desc_mn(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < kPixK / 16; ++k) {
            // 16 pixels = two 8-row groups = 2048 bytes: +128 in the (addr >> 4) field
            umma_bf16_ss(d_tmem, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc,
                         (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long t = tile;
      const int ks = (int)(t % p.ksplit); t /= p.ksplit;
      const int n_i = (int)(t % p.tiles_n); t /= p.tiles_n;
      const int m_i = (int)(t % p.tiles_m); t /= p.tiles_m;
      const int tap = (int)(t % p.ntaps); t /= p.ntaps;
      const int g = (int)t;
      const int kb0 = ks * kb_per, kb1 = min(kb0 + kb_per, kb_total);
      const int o = m_i * kBlockM + row;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + ch * 32, r);
        tmem_ld_wait();
        const int i0 = n_i * BLOCK_N + ch * 32;
        if (o < p.cout && i0 < p.cin && kb1 > kb0) {
          float *dst = p.gw + (((long long)g * p.ntaps + tap) * p.cout + o) * p.cin + i0;
          if (p.ksplit == 1) {
            if (i0 + 32 <= p.cin && (p.cin & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                   __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (i0 + j < p.cin) dst[j] = __uint_as_float(r[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (i0 + j < p.cin) atomicAdd(dst + j, __uint_as_float(r[j]));
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BLOCK_N>
int launch_wgrad(WgradParams &p, const CUtensorMap &tdy, const void *x, int64_t in_h, int64_t in_w,
                 cudaStream_t stream) {
  using C = WCfg<BLOCK_N>;
  auto kern = wgrad_kernel<BLOCK_N>;
  static bool attr_done[64] = {false};
  int dev = 0;
  VSP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  CUtensorMap tx;
  {
    uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)p.batch};
    uint64_t strides[4] = {0, (uint64_t)p.cin * 2, (uint64_t)p.cin * in_w * 2, (uint64_t)p.cin * in_w * in_h * 2};
    uint32_t box[4] = {64, (uint32_t)(p.kwb * p.stride), (uint32_t)(p.khb * p.stride), 1};
    uint32_t es[4] = {1, (uint32_t)p.stride, (uint32_t)p.stride, 1};
    if (int rc = encode_tma(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, es,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  p.tiles_m = (p.cout + kBlockM - 1) / kBlockM;
  p.tiles_n = (p.cin + BLOCK_N - 1) / BLOCK_N;
  const long long base_tiles = (long long)p.groups * p.ntaps * p.tiles_m * p.tiles_n;
  const int samples = p.groups == 1 ? p.batch : 1;
  const long long kb_total = (long long)samples * p.kb_h * p.kb_w;
  // split K: pick the slice count that minimises  waves x (k-blocks per slice + epilogue), where a wave is one tile per SM
  // and the epilogue (TMEM -> registers -> fp32 stores / atomics of a 128 x BLOCK_N tile) costs about as much as 12
  // k-blocks of MMAs.  The old rule (~2 tiles per SM) gave e.g. 360 tiles = 2.4 waves for 512 -> 512 at 64x64, batch 8;
  // this one takes 144 tiles = one full wave (181 -> 150 us).  Slices keep >= 8 k-blocks.
  long long ks = 1;
  {
    const long long sms = num_sms();
    const long long max_ks = kb_total / 8 > 0 ? kb_total / 8 : 1;
    double best = 1e30;
    for (long long c = 1; c <= max_ks && c <= 64; ++c) {
      const long long per = (kb_total + c - 1) / c;
      const long long slices = (kb_total + per - 1) / per;        // no empty slices
      const long long waves = (base_tiles * slices + sms - 1) / sms;
      const double cost = (double)waves * ((double)per + (slices > 1 ? 16.0 : 12.0));
      if (cost < best - 1e-9) { best = cost; ks = slices; }
    }
  }
  p.ksplit = (int)ks;
  p.total_tiles = base_tiles * ks;
  if (ks > 1)
    VSP_CUDA(cudaMemsetAsync(p.gw, 0, sizeof(float) * p.groups * p.ntaps * (size_t)p.cout * p.cin, stream));
  long long grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kern<<<(unsigned)grid, kNumThreads, C::SMEM_BYTES, stream>>>(p, tdy, tx);
  return check_launch("wgrad_kernel");
}

inline int next_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace
}  // namespace vsp

extern "C" int vsp_conv2d_wgrad_bf16(const void *dy, const void *x, float *gw, int64_t batch, int64_t groups,
                                     int64_t in_h, int64_t in_w, int64_t cin, int64_t cout, int64_t out_h,
                                     int64_t out_w, int kh, int kw, int stride, int pad_h, int pad_w, int dil_h,
                                     int dil_w, void *stream_) {
  using namespace vsp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VSP_REQUIRE(batch >= 1 && (groups == 1 || groups == batch), "wgrad: groups must be 1 or batch");
  VSP_REQUIRE(cin >= 8 && cin % 8 == 0 && cout >= 8 && cout % 8 == 0,
              "wgrad: channel counts must be multiples of 8 (NHWC padding), got cin=%lld cout=%lld",
              (long long)cin, (long long)cout);
  VSP_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= 16, "wgrad: kernel up to 16 taps");
  VSP_REQUIRE(stride == 1 || stride == 2, "wgrad: stride must be 1 or 2");
  VSP_REQUIRE(dy && x && gw, "wgrad: null pointer");
  VSP_REQUIRE(out_h >= 1 && out_w >= 1 && in_h >= 1 && in_w >= 1, "wgrad: empty extent");
  VSP_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              "wgrad: operands must be 16-byte aligned");

  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.batch = (int)batch; p.groups = (int)groups; p.cin = (int)cin; p.cout = (int)cout;
  p.out_h = (int)out_h; p.out_w = (int)out_w; p.stride = stride; p.ntaps = kh * kw;
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) {
      p.tap_dy[i * kw + j] = i * dil_h - pad_h;
      p.tap_dx[i * kw + j] = j * dil_w - pad_w;
    }
  p.kwb = next_pow2((int)out_w) < kPixK ? next_pow2((int)out_w) : kPixK;
  p.khb = kPixK / p.kwb;
  p.kb_w = ((int)out_w + p.kwb - 1) / p.kwb;
  p.kb_h = ((int)out_h + p.khb - 1) / p.khb;
  p.gw = gw;

  CUtensorMap tdy;
  {
    uint64_t dims[4] = {(uint64_t)cout, (uint64_t)out_w, (uint64_t)out_h, (uint64_t)batch};
    uint64_t strides[4] = {0, (uint64_t)cout * 2, (uint64_t)cout * out_w * 2, (uint64_t)cout * out_w * out_h * 2};
    uint32_t box[4] = {64, (uint32_t)p.kwb, (uint32_t)p.khb, 1};
    if (int rc = encode_tma(&tdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, dims, strides, box, nullptr,
                            CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  if (cin > 128) return launch_wgrad<256>(p, tdy, x, in_h, in_w, stream);
  if (cin > 64) return launch_wgrad<128>(p, tdy, x, in_h, in_w, stream);
  return launch_wgrad<64>(p, tdy, x, in_h, in_w, stream);
}
