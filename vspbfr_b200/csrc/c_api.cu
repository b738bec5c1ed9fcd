// c_api.cu — bookkeeping half of the C ABI (include/vsp_b200.h): error string,
// launch counter, SM count cache and the TMA descriptor encoder.
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <string.h>

namespace vsp {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
  count_launch(1);
  return 0;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tma(CUtensorMap *map, CUtensorMapDataType dtype, int rank, const void *base,
               const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box,
               const uint32_t *elem_strides, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) s[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(map, dtype, (cuuint32_t)rank, const_cast<void *>(base), d, s, b, e,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(
        "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u "
        "%u %u %u]",
        (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
        (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0),
        (unsigned long long)(rank > 4 ? d[4] : 0), b[0], rank > 1 ? b[1] : 0,
        rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0, rank > 4 ? b[4] : 0);
  }
  return 0;
}

}  // namespace vsp

extern "C" {

int vsp_version(void) { return VSP_ABI_VERSION; }

const char *vsp_last_error(void) { return vsp::g_err; }

int64_t vsp_launch_count(void) { return vsp::g_launches.load(std::memory_order_relaxed); }

int64_t vsp_upfirdn2d_out_size(int64_t in, int k, int up, int down, int pad0, int pad1) {
  if (down <= 0) return 0;
  int64_t num = in * up + pad0 + pad1 - k + down;
  // floor division (python //) so that an empty/negative extent stays <= 0
  int64_t q = num / down;
  if ((num % down != 0) && (num < 0)) --q;
  return q;
}

}  // extern "C"

// Debug aid (not part of the public header): encode a 3-D fp32 map and return its 128 bytes.
extern "C" int vsp_debug_tma_3d_f32(const void *base, uint64_t w, uint64_t h, uint64_t planes, uint32_t bw,
                                    uint32_t bh, uint32_t bz, unsigned char *out128) {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  uint64_t dims[3] = {w, h, planes};
  uint64_t strides[3] = {0, w * 4, w * h * 4};
  uint32_t box[3] = {bw, bh, bz};
  int rc = vsp::encode_tma(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, nullptr,
                           CU_TENSOR_MAP_SWIZZLE_NONE);
  memcpy(out128, &m, 128);
  return rc;
}
