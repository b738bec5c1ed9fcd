// Minimal TMA probe: isolates which descriptor / instruction variant the B200 accepts.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float *out, int nfloats, int c0, int c1, int c2, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  float *tile = reinterpret_cast<float *>(smem);
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) tile[i] = -2.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(nfloats * 4) : "memory");
    if (RANK == 2) {
      if (variant == 0)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(s32(tile)), "l"((uint64_t)&tm), "r"(s32(&bar)), "r"(c0), "r"(c1) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(s32(tile)), "l"((uint64_t)&tm), "r"(s32(&bar)), "r"(c0), "r"(c1) : "memory");
    } else {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(tile)), "l"((uint64_t)&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
  }
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char **argv) {
  int test = argc > 1 ? atoi(argv[1]) : 0;
  void *fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int W = 256, H = 256, P = 4;
  float *x, *out;
  cudaMalloc(&x, sizeof(float) * W * H * P);
  cudaMalloc(&out, sizeof(float) * 65536);
  float *hx = (float *)malloc(sizeof(float) * W * H * P);
  for (int i = 0; i < W * H * P; ++i) hx[i] = (float)(i % 100003);
  cudaMemcpy(x, hx, sizeof(float) * W * H * P, cudaMemcpyHostToDevice);
  { float *neg = (float *)malloc(sizeof(float) * 65536); for (int i = 0; i < 65536; ++i) neg[i] = -1.f;
    cudaMemcpy(out, neg, sizeof(float) * 65536, cudaMemcpyHostToDevice); }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  CUresult r;
  int nfl = 0;
  cuuint32_t es[3] = {1, 1, 1};
  if (test == 0 || test == 1) {  // 2D fp32 box 32x8, .tile / no .tile
    cuuint64_t d[2] = {W, H}; cuuint64_t s[1] = {W * 4}; cuuint32_t b[2] = {32, 8};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    nfl = 32 * 8;
    printf("encode rc=%d\n", (int)r);
    probe<2><<<1, 128, 32768>>>(tm, out, nfl, 0, 0, 0, test);
  } else if (test == 2) {  // 3D fp32 box 32x8x1
    cuuint64_t d[3] = {W, H, P}; cuuint64_t s[2] = {W * 4, W * H * 4}; cuuint32_t b[3] = {32, 8, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    nfl = 32 * 8;
    printf("encode rc=%d\n", (int)r);
    probe<3><<<1, 128, 32768>>>(tm, out, nfl, 0, 0, 1, 0);
  } else if (test == 3) {  // 3D fp32 box 132x19x1 (my failing shape), L2 256B
    cuuint64_t d[3] = {W, H, P}; cuuint64_t s[2] = {W * 4, W * H * 4}; cuuint32_t b[3] = {132, 19, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    nfl = 132 * 19;
    printf("encode rc=%d\n", (int)r);
    probe<3><<<1, 128, 32768>>>(tm, out, nfl, -2, -2, 1, 0);
  } else if (test == 4) {  // 3D fp32 box 64x19x1
    cuuint64_t d[3] = {W, H, P}; cuuint64_t s[2] = {W * 4, W * H * 4}; cuuint32_t b[3] = {64, 19, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    nfl = 64 * 19;
    printf("encode rc=%d\n", (int)r);
    probe<3><<<1, 128, 32768>>>(tm, out, nfl, -2, -2, 1, 0);
  } else if (test == 5) {  // 3D fp32 box 12x10x4 , negative coords
    cuuint64_t d[3] = {W, H, P}; cuuint64_t s[2] = {W * 4, W * H * 4}; cuuint32_t b[3] = {12, 10, 4};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    nfl = 12 * 10 * 4;
    printf("encode rc=%d\n", (int)r);
    probe<3><<<1, 128, 32768>>>(tm, out, nfl, -1, -1, 0, 0);
  }
  if (test == 9) {  // ./tma_probe 9 bw bh bz c0 c1 c2 l2
    cuuint32_t b[3] = {(cuuint32_t)atoi(argv[2]), (cuuint32_t)atoi(argv[3]), (cuuint32_t)atoi(argv[4])};
    int l2 = atoi(argv[8]);
    cuuint64_t d[3] = {W, H, P}; cuuint64_t s[2] = {W * 4, W * H * 4};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    nfl = b[0] * b[1] * b[2];
    printf("box %u %u %u coords %s %s %s l2 %d encode rc=%d\n", b[0], b[1], b[2], argv[5], argv[6], argv[7], l2, (int)r);
    probe<3><<<1, 128, 32768>>>(tm, out, nfl, atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), 0);
  }
  cudaError_t le = cudaGetLastError(); printf("launch: %s\n", cudaGetErrorString(le));
  cudaError_t e = cudaDeviceSynchronize();
  printf("test %d: %s\n", test, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    float h[8];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("  out[0..7] = %g %g %g %g %g %g %g %g\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    unsigned char *tb = (unsigned char *)&tm; printf("  desc: "); for (int i = 0; i < 64; ++i) printf("%02x", tb[i]); printf("\n");
  }
  return 0;
}
