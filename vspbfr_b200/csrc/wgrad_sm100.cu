// placeholder until the tcgen05 wgrad kernel lands
#include "common.cuh"
extern "C" int vsp_conv2d_wgrad_bf16(const void *, const void *, float *, int64_t, int64_t, int64_t, int64_t, int64_t,
                                     int64_t, int64_t, int64_t, int, int, int, int, int, void *) {
  return vsp::set_error("vsp_conv2d_wgrad_bf16: not implemented yet");
}
