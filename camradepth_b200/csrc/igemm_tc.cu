// tcgen05 / TMEM / TMA implicit-GEMM convolution (bf16 operands, fp32 accumulation in tensor memory).
// (under construction: entry points report "unsupported" until the kernels land)
#include "common.cuh"
#include "../../include/camradepth_b200.h"

extern "C" int crd_conv_fwd_tc(const crd_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                               float* gn_sums, crd_stream_t stream) {
  return -3;
}
extern "C" int crd_conv_wgrad_tc(const crd_conv_desc* d, const void* x, const void* dy, float* dw,
                                 crd_stream_t stream) {
  return -3;
}
