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
 * Handles are bound to the device that was current at creation.  hrp_model entry points take a per-handle lock and
 * order successive users of one plan's buffers with events, so a handle may be driven from several threads / streams;
 * the other handle types are not thread-safe.
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

/* HRP_CONV_UP2: 1x1 conv + nearest upsampling by 2 (the addends of the epilogue live at the OUTPUT resolution): the branch-0
 * term of an HRNet fuse layer, HRnet.py:197-208,254-263 */
enum { HRP_CONV = 0, HRP_DECONV_K4S2P1 = 1, HRP_STEM_S2D = 2, HRP_CONV_UP2 = 3 };
enum { HRP_IMPL_TCGEN05 = 0, HRP_IMPL_SIMT_CHECK = 1 };

typedef struct hrp_conv_desc {
  int32_t kind;               /* HRP_CONV | HRP_DECONV_K4S2P1 | HRP_STEM_S2D | HRP_CONV_UP2 */
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
/* debug (hrp_conv_set_timeline): device buffer of 8 CTAs x 8 tiles x 16 int64 clock64() stamps per pipeline role plus
 * ONE trailing int64 = number of each CTA's leading tiles to skip (8 * 8 * 16 + 1 elements), or NULL to disable */
/* kernel variant of a planned conv: 0 = one tile per CTA, 1 = persistent, 2 = halo-tile (3x3 s1, Cin=Cout in {32,64});
 * set_variant(2) returns HRP_ERR_UNSUPPORTED where the halo kernel is not eligible */
int hrp_conv_set_variant(hrp_conv* conv, int32_t variant);
int hrp_conv_variant(const hrp_conv* conv);
int hrp_conv_set_timeline(hrp_conv* conv, long long* dev_buf);
/* one-line description of the planned launch (kernel variant, tiling, buffering, shared memory, grid) */
int hrp_conv_describe(const hrp_conv* conv, char* buf, int64_t buflen);

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
/* Row f4 (training support): reverse mode of hrp_fk / hrp_project -- the gradients torch autograd computes through
 * URDFRobot.get_keypoints[_root] and point_projection_from_3d_tensor inside the reference's training losses
 * (lib/core/function.py:253-311).  rot is the 6-D representation; grad_xyz (B,nkpt,3) -> grad_q (B,dof), grad_rot (B,6),
 * grad_trans (B,3) (the latter two NULL for the only_fk variants, use_b2c = 0); grad_uv (B,N,2) -> grad_pts (B,N,3). */
int hrp_fk_backward(hrp_robot* robot, const float* q, const float* rot, const float* trans, int32_t root, int32_t use_b2c,
                    const float* grad_xyz, float* grad_q, float* grad_rot, float* grad_trans, int32_t B, void* stream);
int hrp_project_backward(const float* K, const float* pts, const float* grad_uv, float* grad_pts, int32_t B, int32_t N,
                         void* stream);
/* Link transforms.  all_links = 0: URDFRobot.get_TWL (lib/utils/urdf_robot.py:107-111) -- out_T (B,nkpt,4,4), the
 * base-frame transforms of the keypoint links in link_names order, translation scaled by global_scale.
 * all_links = 1: URDF.link_fk_batch (lib/utils/urdfpytorch/urdf.py:3061-3149) over EVERY link of the description --
 * out_T (B,n_links,4,4) in the row order handed to hrp_robot_set_full_tree (any link count; rows parent-before-child). */
int hrp_robot_set_full_tree(hrp_robot* robot, const hrp_link_row* rows, int32_t n_links);
int hrp_link_fk(hrp_robot* robot, const float* q, int32_t B, int32_t all_links, float global_scale, float* out_T,
                void* stream);

/* ------------------------------------------------------------------------------------------------
 * Standalone geometry operators of the head (the fused head kernel evaluates the same expressions on chip).
 *   hrp_inv_intrinsics: get_intrinsic_matrix_batch(f, c, bsz, inv=True) (lib/utils/integral.py:56-73,
 *     lib/utils/transforms.py:145-162): K (B,3,3) -> K^-1 (B,3,3), divisions in fp64, stored fp32.
 *   hrp_uvd_to_xyz: uvd_to_xyz (lib/utils/transforms.py:33-73): uvd (B,N,3), K^-1 (B,3,3), root_trans (B,3) ->
 *     xyz (B,N,3) metres; return_relative subtracts root_trans.
 *   hrp_uvz2xyz_singlepoint: uvz2xyz_singlepoint (lib/utils/transforms.py:133-143): uv (B,2) px, z (B), K (B,3,3) ->
 *     xyz (B,3).
 * ------------------------------------------------------------------------------------------------ */
int hrp_inv_intrinsics(const float* K, float* Kinv, int32_t B, void* stream);
int hrp_uvd_to_xyz(const float* uvd, const float* Kinv, const float* root_trans, float image_size, float depth_factor,
                   int32_t return_relative, int32_t B, int32_t N, float* xyz, void* stream);
int hrp_uvz2xyz_singlepoint(const float* uv, const float* z, const float* K, int32_t B, float* xyz, void* stream);

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
  int32_t heatmap_fp32;    /* 0: `heatmap` is bf16 (the layout the final convolution writes); 1: fp32, same pixel-major layout
                              (a caller's fp32 logits, no rounding: twice the bytes, twice the time) */
} hrp_head_args;

int hrp_head_workspace_bytes(int32_t B, int32_t nkpt, int64_t* bytes);
int hrp_head(const hrp_head_args* args, void* stream);
/* Row f4 (first piece): gradient of the heatmap integral w.r.t. the logits -- what autograd computes through
 * HeatmapIntegralPose.forward (lib/utils/integral.py:97-135) in training (lib/core/function.py:253-311).
 * Must follow hrp_head on the SAME heatmap and workspace (it re-uses the per-chunk softmax statistics the forward
 * left there).  uvd: the forward's output; grad_uvd (B,nkpt,3) fp32; grad_heatmap: layout of `heatmap`, bf16
 * (out_fp32 = 0) or fp32 (out_fp32 = 1); out_fp32 = 3: `heatmap` itself is fp32 (heatmap_fp32 = 1 in the forward) and
 * so is the gradient. */
int hrp_head_backward_heatmap(const void* heatmap, const float* uvd, const float* grad_uvd, const void* workspace,
                              int64_t workspace_bytes, int32_t B, int32_t nkpt, int32_t ref_kpt, int32_t fix_root,
                              int32_t out_fp32, void* grad_heatmap, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input pipeline: full camera frame + bounding box -> the 256x256 crop the network
 * consumes, its intrinsics and the k_value scalar of the depth head.
 * Replaces, per image, resize_image (lib/dataset/roboutils.py:128-157), CropResizeToAspectAugmentation
 * (lib/dataset/augmentations.py:165-233: torch bilinear, align_corners=False, `(x*255).to(uint8)`),
 * get_K_crop_resize (lib/utils/geometries.py:360-402) as chained by DreamDataset._get_rootnet_data /
 * _get_other_data (lib/dataset/dream.py:281-388, no flip / padding), and the k_value formula of
 * scripts/test.py:141-152.  The crop bytes are bit-exact with the reference's CPU path.
 * bbox must lie inside the frame (the reference's numpy slice assignment raises otherwise).
 * ------------------------------------------------------------------------------------------------ */
typedef struct hrp_crop_args {
  int32_t B, frame_h, frame_w;
  int32_t out_size;         /* 256 (multiple of 4) */
  const uint8_t* frames;    /* device uint8 (B, frame_h, frame_w, 3), HWC as decoded */
  const int32_t* bbox;      /* device (B,4): wmin, hmin, wmax, hmax */
  const double* K_in;       /* device fp64 (B,3,3): camera matrix of the full frame (state['camera']['K']) */
  uint8_t* out_u8;          /* device uint8 (B,3,out,out): the dataset's "images" (feed hrp_model_forward_u8) */
  float* K_out;             /* device fp32 (B,3,3): the dataset's "K" */
  const float* k_bbox;      /* device fp32 (B,4) box for k_value (bbox_strict_bounded_original / extended), or NULL */
  float* k_value;           /* device fp32 (B), or NULL */
  int32_t k_use_crop_K;     /* 0: fx, fy of K_in (args.use_origin_bbox); 1: of K_out (root_K) */
} hrp_crop_args;
int hrp_crop_resize(const hrp_crop_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation metrics on the device.
 * hrp_metrics_batch replaces compute_metrics_batch (lib/utils/metrics.py:8-119) given the predicted camera-frame
 * keypoints (hrp_fk provides them, as URDFRobot.get_keypoints[_root] does at :27-33).
 * per_image rows: 0 error3d, 1 error2d, 2 mean_jointerror, 3 error_depth, 4 batch_error_relative,
 * 5 error3d_relative.  hrp_metrics_summary replaces summary_add_pck (:122-162); out[0..10] = ADD/mean, ADD/median,
 * ADD/AUC, ADD_{1,5,10,20,40,60,80,100}_mm; out[11..21] = ADD_2D/mean, ADD_2D/median, PCK/AUC,
 * PCK_{2.5,...,20}_pixel.
 * ------------------------------------------------------------------------------------------------ */
typedef struct hrp_metrics_args {
  int32_t B, nkpt, dof, ref_kpt;
  int32_t drop_last_joint;   /* 1 for Panda: the finger joint is excluded from mean_jointerror (:83-84) */
  float frame_w, frame_h;    /* 640, 480 (:61) */
  const float* pred_kp3d;    /* device fp32 (B,nkpt,3) */
  const float* gt_kp3d;      /* (B,nkpt,3) */
  const float* gt_kp2d;      /* (B,nkpt,2) original-frame pixels */
  const float* K_original;   /* (B,3,3) */
  const float* pred_joint;   /* (B,dof) or NULL */
  const float* gt_joint;     /* (B,dof) or NULL */
  void* workspace;           /* device, hrp_metrics_workspace_bytes() */
  int64_t workspace_bytes;
  float* per_image;          /* (6,B) */
  float* dis3d;              /* (nkpt) batch mean per keypoint */
  float* dis2d;              /* (nkpt) */
  float* l1_jointerror;      /* (dof), or NULL when pred_joint is NULL */
} hrp_metrics_args;
int hrp_metrics_workspace_bytes(int32_t B, int32_t nkpt, int32_t dof, int64_t* bytes);
int hrp_metrics_batch(const hrp_metrics_args* args, void* stream);
/* dis3d / dis2d: device fp32 (n) per-image errors accumulated over the test set; out: device fp64 [22] */
int hrp_metrics_summary(const float* dis3d, const float* dis2d, int64_t n, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batched PnP.
 * Replaces BPnP_m3d.forward (lib/utils/BPnP.py:114-152: per sample cv2.solvePnP EPNP -> ITERATIVE refinement), the
 * ground-truth rotation source of the evaluation loop on real datasets (scripts/test.py:120-125), plus the
 * angle_axis_to_rotation_matrix -> rotmat_to_rot6d conversion applied to its result (lib/utils/geometries.py:164-232,
 * 117-132).  pose6 = (angle-axis, translation); N >= 6 non-coplanar points.  Forward only.
 * ------------------------------------------------------------------------------------------------ */
int hrp_pnp(const float* pts2d /* (B,N,2) px */, const float* pts3d /* (B,N,3) */, const float* K /* (3,3) | (B,3,3) */,
            int32_t K_batched, int32_t B, int32_t N, float* pose6 /* (B,6) */, float* rot6d /* (B,6) or NULL */,
            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-network inference.
 * Replaces RootNetwithRegInt.forward (lib/models/full_net.py:239-397; construction :38-192 and
 * get_rootNetwithRegInt_model :401-435 incl. the backbone.* -> rootnet_backbone.* remap :423-427) and
 * RootNet.forward (lib/models/depth_net.py:92-137) for the shipped configuration family:
 * backbone_name=resnet50, rootnet_backbone_name=hrnet32, rotation_dim=6, fix_root, no add_fc / multi_kp /
 * reg_joint_map / direct_reg_rot / rot_iterative_matmul (configs/{panda,kuka,baxter}/{full,depthnet}.yaml).
 * Weights are ingested by their reference state_dict key.
 * Call order: create -> set_tensor (every key) -> [set_robot] -> finalize -> forward*.
 * ------------------------------------------------------------------------------------------------ */
typedef struct hrp_model hrp_model;
enum { HRP_MODEL_FULL = 0, HRP_MODEL_DEPTHNET = 1 };

typedef struct hrp_model_desc {
  int32_t kind;          /* HRP_MODEL_FULL | HRP_MODEL_DEPTHNET */
  int32_t dof, nkpt;     /* panda 8/7, kuka 7/8, baxter 15/17 (full_net.py:42-51); ignored for the depthnet */
  int32_t ref_kpt;       /* reference_keypoint_id (configs/<robot>/full.yaml) */
  int32_t n_iter;        /* iterations of the pose / rot regressors (4) */
  int32_t fix_root;      /* args.fix_root */
  float image_size;      /* 256 */
  float depth_factor;    /* bbox_3d_shape[2] * 1e-3 */
  int32_t chunk;         /* images per pass through the network (L2-sized sub-batch) */
  int32_t inflight;      /* plan replicas (1..4): consecutive chunks of one large batch alternate between them, and
                            single-chunk forwards enqueued on DIFFERENT caller streams get different replicas */
} hrp_model_desc;

/* device fp32 outputs; any pointer may be NULL.  Shapes as returned by the reference forward (8-tuple). */
typedef struct hrp_outputs {
  float* pose;     /* (B,dof) */
  float* rot;      /* (B,6) */
  float* trans;    /* (B,3) */
  float* root_uv;  /* (B,2) */
  float* depth;    /* (B,1) metres */
  float* uvd;      /* (B,nkpt,3) */
  float* xyz_int;  /* (B,nkpt,3) */
  float* xyz_fk;   /* (B,nkpt,3) */
  float* uv_int;   /* (B,nkpt,2) projection of xyz_int with K (scripts/test.py:179) -- extra */
  float* uv_fk;    /* (B,nkpt,2) projection of xyz_fk with K -- extra */
} hrp_outputs;

int hrp_model_create(const hrp_model_desc* desc, hrp_model** out);
void hrp_model_destroy(hrp_model* model);
/* data: host fp32, copied.  Integer buffers (num_batches_tracked) need not be passed. */
int hrp_model_set_tensor(hrp_model* model, const char* name, const float* data, const int64_t* shape, int32_t ndim);
int hrp_model_set_robot(hrp_model* model, const hrp_robot* robot);   /* robot must outlive the model */
int hrp_model_finalize(hrp_model* model);
/* x_reg, x_root: device fp32 (B,3,256,256) in [0,1]; k_value (B); K (B,3,3); init_pose (B,dof) / init_rot (B,6)
 * or NULL for the registered buffers. */
int hrp_model_forward(hrp_model* model, const float* x_reg, const float* x_root, const float* k_value, const float* K,
                      const float* init_pose, const float* init_rot, int32_t B, const hrp_outputs* out, void* stream);
/* Same, with uint8 images (B,3,256,256) in 0..255: fuses the caller-side `images.float() / 255.` of
 * scripts/test.py:83-86 into the input-packing kernel (4x less host->device traffic). */
int hrp_model_forward_u8(hrp_model* model, const uint8_t* x_reg, const uint8_t* x_root, const float* k_value,
                         const float* K, const float* init_pose, const float* init_rot, int32_t B,
                         const hrp_outputs* out, void* stream);
/* RootNet.forward: x (B,3,256,256), k_value (B) -> depth in millimetres (B,1) */
int hrp_model_depthnet_forward(hrp_model* model, const float* x, const float* k_value, int32_t B, float* depth_mm,
                               void* stream);
/* test / profiling hooks: named intermediate activation of the most recently used plan (bf16 NHWC device
 * pointer; C < 0 means an fp32 (B,|C|) vector), and per-plan statistics */
int hrp_model_activation(hrp_model* model, const char* name, const void** ptr, int32_t* B, int32_t* H, int32_t* W,
                         int32_t* C);
/* asynchronous device-to-device copy on `stream` (used by the shims to snapshot activations) */
int hrp_copy_device(void* dst, const void* src, int64_t bytes, void* stream);
/* per-operation timing of one plan (eager, CUDA events, `iters` back-to-back launches per op): tab-separated text
 * name, kind, lane, in HxW, Cin, Cout, out HxW, taps, n_tile, epilogue, CTAs, stages, us, TFLOP/s, GB/s */
int hrp_model_profile(hrp_model* model, int32_t batch, int32_t iters, char* buf, int64_t buflen);
/* Kernel-variant table: text lines "<conv shape signature> <variant>" (variant 0 one tile per CTA, 1 persistent,
 * 2 halo-tile).  The shim loads the committed table (tuning/b200.txt) before finalize so that every box and every run
 * picks the same kernels (bitwise reproducible results); shapes without an entry use shape heuristics, or -- with
 * HRP_AUTOTUNE=1 in the environment -- are timed at plan build.  get_tuning dumps the current table (set + tuned). */
int hrp_model_set_tuning(hrp_model* model, const char* text);
int hrp_model_get_tuning(hrp_model* model, char* buf, int64_t buflen);
/* Host-only planning pass (no device work; callable before finalize once the tensors are set): the network program for
 * `batch` images is traversed for tensor shapes and liveness and the activation arena is laid out -- alias = 1 re-uses the
 * memory of tensors whose last reader has run (single-lane plans), alias = 0 gives every tensor its own region. */
int hrp_model_plan_memory(hrp_model* model, int32_t batch, int32_t alias, int64_t* arena_bytes, int64_t* tensor_bytes,
                          int32_t* n_tensors, int32_t* n_ops);
int hrp_model_stats(const hrp_model* model, int32_t batch, double* flops, int32_t* kernels, int64_t* activation_bytes);

#ifdef __cplusplus
}
#endif
#endif /* HRP_H_ */
