/* hrp.h -- C ABI of the B200-native HoRoPose inference path (libhrp_b200.so).
 *
 * The reference (Oliverbansk/Holistic-Robot-Pose-Estimation) is pure Python: it has no FFI layer, so the
 * drop-in boundary is its Python call signatures.  Each group of entry points below
 * names the reference interface it replaces (file:line relative to the reference root); the ctypes shim in
 * holistic-robot-pose-estimation_b200/ presents those Python signatures on top of this ABI
 * (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions: extern "C"; plain pointers and sizes; every function returns 0 on success or a negative
 * hrp_status; the message is available from hrp_last_error() (thread-local).  Device pointers are owned by
 * the caller (PyTorch allocates them and passes data_ptr()); handles own their packed weights, workspaces
 * and CUDA graphs.  `stream` is a cudaStream_t passed as void*; launches are asynchronous on it.
 * Handles are bound to the device that was current at creation and are not thread-safe.
 */
#ifndef HRP_H_
#define HRP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum hrp_status {
  HRP_STATUS_OK = 0,
  HRP_STATUS_INVALID = -1,     /* bad argument / shape */
  HRP_STATUS_CUDA = -2,        /* CUDA runtime or driver error (message has the detail) */
  HRP_STATUS_STATE = -3,       /* call order violated (e.g. forward before finalize) */
  HRP_STATUS_UNSUPPORTED = -4  /* configuration outside the hot path */
} hrp_status;

const char* hrp_last_error(void);
/* library / build identification: "hrp_b200 <version> sm_100a" */
const char* hrp_version(void);
/* number of kernels this library has launched since load (all handles, this process) */
int64_t hrp_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Convolution operator (implicit GEMM on tcgen05/TMEM fed by TMA).
 * Replaces the torch.nn.Conv2d / ConvTranspose2d (+ BatchNorm2d eval + ReLU + residual / fuse adds) calls of
 *   lib/models/backbones/HRnet.py:22-98 (BasicBlock, Bottleneck), :197-265 (fuse layers), :341-429
 *   lib/models/backbones/Resnet.py:21-29, :96-135
 *   lib/models/full_net.py:194-216 (deconv head), :78 (final 1x1)
 * Tensors are bf16 NHWC with the channel count padded to 16 / 32 / a multiple of 64.
 * ------------------------------------------------------------------------------------------------ */
typedef struct hrp_conv hrp_conv;

enum { HRP_CONV = 0, HRP_DECONV_K4S2P1 = 1, HRP_STEM_S2D = 2 };
enum { HRP_IMPL_TCGEN05 = 0, HRP_IMPL_SIMT_CHECK = 1 };

typedef struct hrp_conv_desc {
  int32_t kind;               /* HRP_CONV | HRP_DECONV_K4S2P1 | HRP_STEM_S2D */
  int32_t B, Hin, Win, Cin;   /* input (B,Hin,Win,Cin) bf16 NHWC; HRP_STEM_S2D: the (B,H/2,W/2,16) s2d tensor */
  int32_t Cout;
  int32_t kh, kw, stride, pad;
  int32_t relu;
} hrp_conv_desc;

typedef struct hrp_conv_epilogue {
  const float* scale;         /* device [Cout] fp32: gamma/sqrt(var+eps), or ones */
  const float* bias;          /* device [Cout] fp32: beta - mean*scale (+ conv bias*scale) */
  const void* pre[3];         /* device bf16 (B,Hout,Wout,Cout) added before the ReLU, or NULL */
  const void* up[3];          /* device bf16 (B,Hout>>s,Wout>>s,Cout) nearest-upsampled addends, or NULL */
  int32_t up_shift[3];
  const void* post;           /* device bf16 (B,Hout,Wout,Cout) added after the ReLU, or NULL */
  void* out;                  /* device bf16 (B,Hout,Wout,Cout), may be NULL if pool_out is set */
  float* pool_out;            /* device fp32 (B,Cout) += mean over pixels (caller zeroes it), or NULL */
} hrp_conv_epilogue;

int hrp_conv_packed_weight_elems(const hrp_conv_desc* desc, int64_t* elems);
/* w: host fp32 in the reference layout -- Conv2d (Cout,cin_ref,kh,kw); ConvTranspose2d (cin_ref,Cout,4,4);
 * stem (Cout,3,kh,kw).  out: host uint16 (bf16 bits), hrp_conv_packed_weight_elems() entries. */
int hrp_conv_pack_weights(const hrp_conv_desc* desc, int32_t cin_ref, const float* w, uint16_t* out);
int hrp_conv_create(const hrp_conv_desc* desc, const void* in_dev, const void* w_packed_dev,
                    const hrp_conv_epilogue* epi, hrp_conv** out);
int hrp_conv_out_shape(const hrp_conv* conv, int32_t* Hout, int32_t* Wout);
int hrp_conv_run(hrp_conv* conv, int32_t impl, void* stream);
void hrp_conv_destroy(hrp_conv* conv);

/* ------------------------------------------------------------------------------------------------
 * Input / layout kernels.
 *   hrp_pack_input_s2d: (B,3,H,W) fp32 NCHW in [0,1] (scripts/test.py:83-86) -> (B,H/2,W/2,16) bf16, channel
 *     = (hp*2+wp)*3 + c, 4 zero pad channels: the 2x2 space-to-depth view the stride-2 stem convs consume
 *     (HRnet.py:284, Resnet.py:21).
 *   hrp_maxpool3x3s2: nn.MaxPool2d(3,2,1) of Resnet.py:25 on bf16 NHWC.
 *   hrp_nchw_f32_to_nhwc_bf16 / hrp_nhwc_bf16_to_nchw_f32: layout bridges for the operator-level shims.
 * ------------------------------------------------------------------------------------------------ */
int hrp_pack_input_s2d(const float* x_nchw, void* out_s2d, int32_t B, int32_t H, int32_t W, void* stream);
int hrp_maxpool3x3s2(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
int hrp_nchw_f32_to_nhwc_bf16(const float* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W,
                              int32_t Cpad, void* stream);
int hrp_nhwc_bf16_to_nchw_f32(const void* in, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                              int32_t Cpad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HRP_H_ */
