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

/* ------------------------------------------------------------------------------------------------
 * Robot kinematic table + forward kinematics + projection.
 * Replaces URDFRobot.get_keypoints / get_keypoints_root / get_keypoints_only_fk[_at_specific_root] /
 * get_rotation_at_specific_root / get_TWL (lib/utils/urdf_robot.py:82-199), URDF.link_fk_batch
 * (lib/utils/urdfpytorch/urdf.py:3061-3149, Joint.get_child_poses :2344-2396, _rotation_matrices :2427-2462)
 * and point_projection_from_3d[_tensor] (lib/utils/transforms.py:7-21).
 * The host parses the URDF (host logic, no arithmetic) and hands over a flat table: links in
 * parent-before-child order, pruned to the ancestors of the keypoint links.
 * ------------------------------------------------------------------------------------------------ */
typedef struct hrp_robot hrp_robot;

typedef struct hrp_link_row {
  int32_t parent;      /* index of the parent link, -1 for the base */
  int32_t jtype;       /* 0 fixed, 1 revolute/continuous, 2 prismatic */
  int32_t qcol;        /* column of q driving the joint (mimic joints point at their master), -1 if none */
  int32_t pad;
  double qmul, qoff;   /* cfg = qmul * q[qcol] + qoff */
  double origin[16];   /* 4x4 joint origin, row-major, float64 as parsed from the URDF */
  double axis[3];      /* unit joint axis */
} hrp_link_row;

int hrp_robot_create(const hrp_link_row* rows, int32_t n_links, const int32_t* kp_link, const double* kp_offset,
                     int32_t nkpt, int32_t dof, hrp_robot** out);
void hrp_robot_destroy(hrp_robot* robot);
/* q (B,dof); rot (B,rot_dim), rot_dim 6 (Zhou et al.) or 4 (w,x,y,z); trans (B,3); all device fp32.
 * use_b2c=0 -> "only_fk" variants (rot/trans ignored).  root>0 -> keypoints relative to keypoint `root`
 * (get_keypoints_root).  out_xyz (B,nkpt,3) and/or out_rot (B,rot_dim) (get_rotation_at_specific_root). */
int hrp_fk(hrp_robot* robot, const float* q, const float* rot, int32_t rot_dim, const float* trans, int32_t root,
           int32_t use_b2c, float* out_xyz, float* out_rot, int32_t B, void* stream);
/* K (B,3,3), pts (B,N,3) -> uv (B,N,2), device fp32 */
int hrp_project(const float* K, const float* pts, float* uv, int32_t B, int32_t N, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused head: 3-D heatmap soft-argmax -> uvd -> xyz, root translation, FK and projections in one kernel.
 * Replaces HeatmapIntegralPose.forward (lib/utils/integral.py:97-145), uvd_to_xyz / uvz2xyz_singlepoint /
 * get_intrinsic_matrix_batch (lib/utils/transforms.py:33-73,133-162) and the FK call of
 * RootNetwithRegInt.forward (lib/models/full_net.py:297-305,380-383).
 * heatmap: device bf16 logits, pixel-major (B, 64*64, nkpt*64) with channel = k*64 + d -- the layout the final
 * 1x1 convolution writes (use hrp_nchw_f32_to_nhwc_bf16 to bridge from the reference's NCHW fp32 tensor).
 * ------------------------------------------------------------------------------------------------ */
typedef struct hrp_head_args {
  int32_t B, nkpt, ref_kpt, fix_root;
  float image_size;        /* 256 */
  float depth_factor;      /* bbox_3d_shape[2] * 1e-3 */
  const void* heatmap;
  const float* K;          /* (B,3,3) */
  const float* root_depth; /* (B) metres */
  const hrp_robot* robot;  /* may be NULL: no FK outputs */
  const float* pose;       /* (B,dof) joint angles for the FK, or NULL */
  const float* rot;        /* (B,6) */
  void* workspace;         /* device, hrp_head_workspace_bytes(), zero-initialised once by the caller */
  int64_t workspace_bytes;
  float* uvd;              /* (B,nkpt,3) */
  float* xyz_int;          /* (B,nkpt,3) */
  float* root_uv;          /* (B,2) */
  float* trans;            /* (B,3) */
  float* xyz_fk;           /* (B,nkpt,3) or NULL */
  float* uv_int;           /* (B,nkpt,2) or NULL: projection of xyz_int with K */
  float* uv_fk;            /* (B,nkpt,2) or NULL */
} hrp_head_args;

int hrp_head_workspace_bytes(int32_t B, int32_t nkpt, int64_t* bytes);
int hrp_head(const hrp_head_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HRP_H_ */
