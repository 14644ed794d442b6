// Parameter blocks of the dataset-side / evaluation-side kernels (eval.cu); the C ABI in api.cu forwards to these.
#pragma once
#include "hrp_common.cuh"
namespace hrp {

struct CropParams {
  int B, frame_h, frame_w, out_size;
  const uint8_t* frames;  // (B, frame_h, frame_w, 3) uint8 HWC
  const int* bbox;        // (B, 4) wmin, hmin, wmax, hmax
  const double* K_in;     // (B, 3, 3) fp64 camera matrix of the full frame
  uint8_t* out_u8;        // (B, 3, out, out)
  float* K_out;           // (B, 3, 3)
  const float* k_bbox;    // (B, 4) fp32 box for k_value, or nullptr
  float* k_value;         // (B) or nullptr
  int k_use_crop_K;       // 1: fx, fy of the crop's K; 0: of K_in
};
int launch_crop_resize(const CropParams& p, cudaStream_t s);

struct MetricsParams {
  int B, nkpt, dof, ref_kpt, drop_last_joint;
  float frame_w, frame_h;
  const float *pred_kp3d, *gt_kp3d, *gt_kp2d, *K, *pred_joint, *gt_joint;
  float *kp_err3d, *kp_err2d, *kp_valid, *joint_err;  // scratch, (cols, B)
  float* per_image;                                   // (6, B)
  float *dis3d, *dis2d, *l1_jointerror;
};
int launch_metrics_batch(const MetricsParams& p, cudaStream_t s);

struct SummaryParams {
  const float *dis3d, *dis2d;
  long long n;
  int nthr_add, nthr_pck;
  float table_thr[16];  // 8 ADD thresholds (metres, as fp32) then 8 PCK thresholds (pixels)
  double* out;          // [22]
};
int launch_metrics_summary(const SummaryParams& p, cudaStream_t s);

struct PnpParams {
  int B, N, K_batched, max_iters;
  const float *pts2d, *pts3d, *K;
  float *pose6, *rot6d;
};
int launch_pnp(const PnpParams& p, cudaStream_t s);

}  // namespace hrp
