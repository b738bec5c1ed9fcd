// placeholder until the tcgen05 kernel lands (same commit series)
#include "common.cuh"
extern "C" int vsp_conv2d_fprop_bf16(const void *, const void *, void *, int64_t, int64_t, int64_t, int64_t,
                                     int64_t, int64_t, int64_t, int, int, int, int, int, int, int64_t, int64_t,
                                     const vsp_conv_epilogue *, void *) {
  return vsp::set_error("vsp_conv2d_fprop_bf16: not implemented yet");
}
extern "C" int vsp_upfirdn2d_nhwc_bf16(const void *, const float *, void *, int64_t, int64_t, int64_t, int64_t, int,
                                       int, int, int, int, int, int, int, int, int, void *) {
  return vsp::set_error("vsp_upfirdn2d_nhwc_bf16: not implemented yet");
}
