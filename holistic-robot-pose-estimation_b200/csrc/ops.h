// Internal launchers for the non-GEMM kernels (layout.cu, head.cu); the C ABI in api.cu forwards to these.
#pragma once
#include "hrp_common.cuh"
namespace hrp {
struct FuseAddParams {
  const bf16* pre;
  const bf16* up[3];
  int up_shift[3];
  bf16* out;
  int B, H, W, C, relu;
};
int launch_fuse_add(const FuseAddParams& p, cudaStream_t s);
int launch_pack_input_s2d(const float* x, void* out, int B, int H, int W, cudaStream_t s, int out_pitch = 0,
                          int out_off = 0);  // out_pitch: pixels per output row (0 = dense), out_off: left padding
int launch_pack_input_s2d_u8(const uint8_t* x, void* out, int B, int H, int W, cudaStream_t s, int out_pitch = 0,
                             int out_off = 0);
int launch_maxpool3x3s2(const void* in, void* out, int B, int H, int W, int C, cudaStream_t s);
int launch_nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, int Cpad, cudaStream_t s);
int launch_nhwc_bf16_to_nchw_f32(const void* in, float* out, int B, int C, int H, int W, int Cpad, cudaStream_t s);
}  // namespace hrp
