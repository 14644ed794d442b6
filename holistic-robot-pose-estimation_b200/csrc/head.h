// Head of the HoRoPose forward: heatmap soft-argmax, uvd->xyz, root translation, collapsed pose/rot regressors,
// URDF forward kinematics and pinhole projection (reference: lib/utils/integral.py:97-145,
// lib/utils/transforms.py:11-21,33-73,133-162, lib/models/full_net.py:271-287,305-383,
// lib/utils/urdf_robot.py:82-199, lib/utils/urdfpytorch/urdf.py:2344-2396,2427-2462,3061-3149).
#pragma once
#include "hrp_common.cuh"

namespace hrp {

constexpr int kMaxLinks = 32;
constexpr int kMaxKpt = 32;
constexpr int kMaxDof = 16;

// Flat kinematic table (links in parent-before-child order, pruned to the ancestors of the keypoint links).
struct RobotTable {
  int n_links, nkpt, dof, pad0;
  int parent[kMaxLinks];      // -1 for the base
  int jtype[kMaxLinks];       // 0 fixed, 1 revolute/continuous, 2 prismatic
  int qcol[kMaxLinks];        // -1: joint not driven
  float qmul[kMaxLinks], qoff[kMaxLinks];
  float origin[kMaxLinks][12];  // rows 0..2 of the 4x4 joint origin, fp32 (urdf.py:2383 casts to q's dtype)
  float axis[kMaxLinks][3];     // unit axis (urdf.py:2169,2446)
  float axis_outer[kMaxLinks][9];  // outer(axis,axis) computed in fp64 then cast (urdf.py:2453-2456)
  int kp_link[kMaxKpt];
  float kp_off[kMaxKpt][3];
};

// Collapsed iterative regressor (full_net.py:318-331,365-378 are affine maps: no activation, dropout = identity):
//   delta(p) = Wx * xf + A * p + c ;   p <- p + delta(p)   n_iter times.
struct RegressorTable {
  int dim, n_iter;
  const float* Wx;   // device [dim][2048]
  float A[kMaxDof][kMaxDof];
  float c[kMaxDof];
};

struct HeadParams {
  int B, nkpt, ref_kpt, fix_root;
  float image_size, depth_factor;
  // soft-argmax input: bf16 heatmap logits, pixel-major (B, 64*64, nkpt*64): channel = k*64 + d
  const bf16* heatmap;
  int heatmap_f32;          // the logits are fp32 in the same pixel-major layout (standalone operator only)
  int chunks;               // CTAs per image
  float* partials;          // workspace (B, chunks, nkpt, 5)
  unsigned int* counters;   // workspace (B) zero-initialised; self-resetting
  float* merged;            // launch_head_from_partials only: workspace (B, nkpt, 5) for the per-image merged partials
  const float* K;           // (B,3,3)
  // root depth: either depth_in (B) [m], or gamma = depth_w . feat + depth_b ; depth = gamma*k_value/1000
  const float* depth_in;
  const float* feat;        // (B,2048) fp32 pooled HRNet feature
  const float* depth_w;     // (2048)
  float depth_b;
  const float* k_value;     // (B)
  // regressors: either pose_in/rot_in, or xf (B,2048) with the collapsed tables
  const float* pose_in;     // (B,dof)
  const float* rot_in;      // (B,6)
  const float* xf;          // (B,2048) fp32 pooled ResNet feature
  const float* init_pose;   // (dof) or (B,dof) when init_batched
  const float* init_rot;    // (6)
  int init_batched;
  const RobotTable* robot;  // device
  const RegressorTable* reg_pose;  // device
  const RegressorTable* reg_rot;   // device
  // outputs (fp32)
  float* pose;      // (B,dof)
  float* rot;       // (B,6)
  float* trans;     // (B,3)
  float* root_uv;   // (B,2)
  float* depth;     // (B,1)
  float* uvd;       // (B,nkpt,3)
  float* xyz_int;   // (B,nkpt,3)
  float* xyz_fk;    // (B,nkpt,3)
  float* uv_int;    // (B,nkpt,2) optional projections with K (scripts/test.py:179)
  float* uv_fk;     // (B,nkpt,2) optional
};

int launch_head(const HeadParams& p, cudaStream_t s);
// Soft-argmax folded into the final conv's epilogue (conv_gemm.cu EPI_HEAD): `partials` already holds the per-warp
// partials (B, chunks, nkpt, 5) the conv wrote; this kernel merges them per image (fixed order) and finishes the sample
// exactly as the fused head does (uvd -> xyz, root depth, regressors, FK, projections).  `heatmap` is not read.
int launch_head_from_partials(const HeadParams& p, cudaStream_t s);
size_t head_partials_elems(int B, int nkpt, int chunks);
int head_default_chunks(int B);
int launch_head_backward_heatmap(const bf16* heatmap, const float* partials, const float* uvd, const float* grad_uvd, int B,
                                 int nkpt, int ref_kpt, int fix_root, int chunks, bool out_fp32, void* grad_out,
                                 cudaStream_t s, bool in_fp32 = false);

struct FkParams {
  int B, rot_dim, root, use_b2c;
  const float* q;      // (B,dof)
  const float* rot;    // (B,rot_dim) or null
  const float* trans;  // (B,3) or null
  const RobotTable* robot;
  float* pts;          // (B,nkpt,3) or null
  float* rot_out;      // (B,rot_dim) rotation at `root` (get_rotation_at_specific_root) or null
};
int launch_fk(const FkParams& p, cudaStream_t s);

// Reverse mode of launch_fk: gradients of a scalar loss w.r.t. joint angles, 6-D rotation and
// translation given d loss / d keypoints -- what autograd computes through URDFRobot.get_keypoints[_root] in training
// (lib/core/function.py:253-311).  Analytic: geometric Jacobian of the kinematic tree (revolute: axis x lever arm,
// prismatic: axis; mimic joints scaled), rigid root inverse, Gram-Schmidt-by-cross of the 6-D rotation.
struct FkBwdParams {
  int B, root, use_b2c;
  const float* q;         // (B,dof)
  const float* rot;       // (B,6) or null
  const float* trans;     // (B,3) or null
  const RobotTable* robot;
  const float* grad_pts;  // (B,nkpt,3)
  float* grad_q;          // (B,dof)
  float* grad_rot;        // (B,6) or null
  float* grad_trans;      // (B,3) or null
};
int launch_fk_backward(const FkBwdParams& p, cudaStream_t s);
// reverse mode of launch_project w.r.t. the points: uv = (K p)[:2] / (K p)[2]
int launch_project_backward(const float* K, const float* pts, const float* grad_uv, float* grad_pts, int B, int N,
                            cudaStream_t s);

// One link of an UNPRUNED kinematic tree (URDF.link_fk_batch over all links, urdf.py:3061-3149): any number of links,
// rows in parent-before-child order in global memory.
struct LinkRowDev {
  int parent, jtype, qcol, pad;
  float qmul, qoff;
  float origin[12];
  float axis[3];
  float axis_outer[9];
};
// out_T (B, n_links, 4, 4) row-major homogeneous transforms of every link (base frame)
int launch_link_fk_all(const LinkRowDev* rows, int n_links, const float* q, int dof, int B, float* out_T, cudaStream_t s);
// out_T (B, nkpt, 4, 4): transforms of the keypoint links of the pruned table, translation scaled by `scale`
// (URDFRobot.get_TWL, urdf_robot.py:107-111)
int launch_twl(const RobotTable* robot, int n_links, int nkpt, const float* q, int B, float scale, float* out_T,
               cudaStream_t s);
// standalone geometry operators of the head (transforms.py:33-73,133-162; integral.py:56-73)
int launch_inv_intrinsics(const float* K, float* Kinv, int B, cudaStream_t s);
int launch_uvd_to_xyz(const float* uvd, const float* Kinv, const float* root_trans, float image_size, float depth_factor,
                      int return_relative, int B, int N, float* xyz, cudaStream_t s);
int launch_uvz2xyz(const float* uv, const float* z, const float* K, int B, float* xyz, cudaStream_t s);
int launch_project(const float* K, const float* pts, float* uv, int B, int N, cudaStream_t s);
int launch_depth(const float* feat, const float* w, float b, const float* k, float* out, int B, float out_scale,
                 cudaStream_t s);
int launch_pool_mean(const bf16* in, float* out, int B, int HW, int C, cudaStream_t s);

}  // namespace hrp
