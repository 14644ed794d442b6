// C ABI of libhrp_b200.so (include/hrp.h): thin extern "C" forwarding layer, no exceptions cross it.
#include "../../include/hrp.h"

#include "conv.h"
#include "eval.h"
#include "head.h"
#include "launch_count.h"
#include "ops.h"

#include <new>
#include <vector>

namespace hrp {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }
std::atomic<int64_t> g_launch_count{0};
thread_local bool g_pdl_launch = false;
}  // namespace hrp

using namespace hrp;

struct hrp_robot {
  RobotTable host;
  RobotTable* dev = nullptr;
  LinkRowDev* full_rows = nullptr;  // optional unpruned tree (hrp_robot_set_full_tree) for hrp_link_fk(all_links = 1)
  int n_full = 0;
};

struct hrp_conv {
  ConvLayerDesc desc;
  ConvPlan plan;
};

static ConvLayerDesc to_internal(const hrp_conv_desc* d) {
  ConvLayerDesc o;
  o.kind = d->kind;
  o.B = d->B;
  o.Hin = d->Hin;
  o.Win = d->Win;
  o.Cin = d->Cin;
  o.Cout = d->Cout;
  o.kh = d->kh;
  o.kw = d->kw;
  o.stride = d->stride;
  o.pad = d->pad;
  o.relu = d->relu;
  o.has_residual = 0;
  o.head_fold = 0;
  o.in_wpitch = o.in_wpad = 0;
  return o;
}

extern "C" {

const char* hrp_last_error(void) { return hrp::last_error(); }
const char* hrp_version(void) { return "hrp_b200 0.1 sm_100a"; }
int64_t hrp_launch_count(void) { return g_launch_count.load(); }

int hrp_conv_packed_weight_elems(const hrp_conv_desc* desc, int64_t* elems) {
  HRP_REQUIRE(desc != nullptr && elems != nullptr, "null argument");
  ConvParams p;
  int rc = conv_geometry(to_internal(desc), &p);
  if (rc != HRP_OK) return rc;
  *elems = (int64_t)conv_packed_weight_elems(p);
  return HRP_OK;
}

int hrp_conv_pack_weights(const hrp_conv_desc* desc, int32_t cin_ref, const float* w, uint16_t* out) {
  HRP_REQUIRE(desc != nullptr && w != nullptr && out != nullptr, "null argument");
  ConvParams p;
  ConvLayerDesc d = to_internal(desc);
  int rc = conv_geometry(d, &p);
  if (rc != HRP_OK) return rc;
  return conv_pack_weights(d, p, cin_ref, w, out);
}

int hrp_conv_create(const hrp_conv_desc* desc, const void* in_dev, const void* w_packed_dev,
                    const hrp_conv_epilogue* epi, hrp_conv** out) {
  HRP_REQUIRE(desc != nullptr && epi != nullptr && out != nullptr, "null argument");
  HRP_REQUIRE(epi->scale != nullptr && epi->bias != nullptr, "scale/bias are required");
  HRP_REQUIRE(epi->out != nullptr || epi->pool_out != nullptr, "an output is required");
  hrp_conv* c = new (std::nothrow) hrp_conv();
  HRP_REQUIRE(c != nullptr, "out of host memory");
  c->desc = to_internal(desc);
  c->desc.has_residual = (epi->pre[0] != nullptr) ? 1 : 0;
  int rc = conv_geometry(c->desc, &c->plan.p);
  if (rc == HRP_OK) {
    ConvParams& p = c->plan.p;
    p.scale = epi->scale;
    p.bias = epi->bias;
    for (int i = 0; i < 3; ++i) {
      p.pre[i] = reinterpret_cast<const bf16*>(epi->pre[i]);
      p.up[i] = reinterpret_cast<const bf16*>(epi->up[i]);
      p.up_shift[i] = epi->up_shift[i];
    }
    p.post = reinterpret_cast<const bf16*>(epi->post);
    p.out = reinterpret_cast<bf16*>(epi->out);
    p.pool_out = epi->pool_out;
    rc = conv_plan_finalize(&c->plan, reinterpret_cast<const bf16*>(in_dev),
                            reinterpret_cast<const bf16*>(w_packed_dev), &c->desc);
  }
  if (rc != HRP_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return HRP_OK;
}

int hrp_conv_out_shape(const hrp_conv* conv, int32_t* Hout, int32_t* Wout) {
  HRP_REQUIRE(conv != nullptr && Hout != nullptr && Wout != nullptr, "null argument");
  *Hout = conv->plan.p.Hout;
  *Wout = conv->plan.p.Wout;
  return HRP_OK;
}

int hrp_conv_run(hrp_conv* conv, int32_t impl, void* stream) {
  HRP_REQUIRE(conv != nullptr, "null conv handle");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (impl == HRP_IMPL_TCGEN05) return conv_plan_launch(conv->plan, s);
  if (impl == HRP_IMPL_SIMT_CHECK) return conv_plan_launch_simt(conv->plan, s);
  set_error("unknown conv impl");
  return HRP_ERR_INVALID;
}

void hrp_conv_destroy(hrp_conv* conv) { delete conv; }

int hrp_conv_set_variant(hrp_conv* conv, int32_t variant) {
  HRP_REQUIRE(conv != nullptr, "null conv handle");
  HRP_REQUIRE(variant >= 0 && variant <= 2, "variant must be 0 (tile), 1 (persistent) or 2 (halo)");
  if (variant == 2 && !conv->plan.halo_ok) {
    set_error("the halo-tile kernel cannot run this layer");
    return HRP_ERR_UNSUPPORTED;
  }
  conv->plan.halo = (variant == 2);
  if (variant < 2) conv->plan.persistent = (variant == 1);
  return HRP_OK;
}

int hrp_conv_variant(const hrp_conv* conv) {
  if (conv == nullptr) return -1;
  return conv->plan.halo ? 2 : (conv->plan.persistent ? 1 : 0);
}

int hrp_conv_describe(const hrp_conv* conv, char* buf, int64_t buflen) {
  HRP_REQUIRE(conv != nullptr && buf != nullptr && buflen > 0, "bad argument");
  const ConvPlan& pl = conv->plan;
  if (pl.halo)
    snprintf(buf, (size_t)buflen, "halo pair=%d T=%d bands=%d ring=%d NR=%d smem=%d grid=%u", pl.hp.pair, pl.hp.T, pl.hp.n_abuf,
             pl.hp.nring, pl.hp.NR, pl.halo_smem, pl.halo_grid);
  else if (pl.persistent)
    snprintf(buf, (size_t)buflen,
             "persistent epi=%d n_tile=%d stages=%d sub=%d nprod=%d dual=%d res_store=%d staged=%d nstag=%d vsh=%d wres=%d smem=%d grid=%u",
             pl.epi, pl.p.n_tile, pl.pcfg.stages, pl.pcfg.sub, pl.pcfg.nprod, pl.pcfg.dual, pl.pcfg.res_store, pl.pcfg.staged,
             pl.pcfg.nstag, pl.pcfg.vsh, pl.pcfg.wres, pl.psmem, pl.pgrid);
  else
    snprintf(buf, (size_t)buflen, "tile epi=%d n_tile=%d stages=%d smem=%d grid=%u", pl.epi, pl.p.n_tile, pl.stages,
             pl.smem_bytes, pl.grid.x * pl.grid.z);
  return HRP_OK;
}

int hrp_conv_set_timeline(hrp_conv* conv, long long* dev_buf) {
  HRP_REQUIRE(conv != nullptr, "null conv handle");
  conv->plan.p.timeline = dev_buf;
  return HRP_OK;
}

int hrp_pack_input_s2d(const float* x_nchw, void* out_s2d, int32_t B, int32_t H, int32_t W, void* stream) {
  HRP_REQUIRE(x_nchw != nullptr && out_s2d != nullptr && B > 0, "bad argument");
  return launch_pack_input_s2d(x_nchw, out_s2d, B, H, W, reinterpret_cast<cudaStream_t>(stream));
}
int hrp_maxpool3x3s2(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  HRP_REQUIRE(in != nullptr && out != nullptr && B > 0, "bad argument");
  return launch_maxpool3x3s2(in, out, B, H, W, C, reinterpret_cast<cudaStream_t>(stream));
}
int hrp_nchw_f32_to_nhwc_bf16(const float* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                              void* stream) {
  HRP_REQUIRE(in != nullptr && out != nullptr && B > 0 && Cpad >= C, "bad argument");
  return launch_nchw_f32_to_nhwc_bf16(in, out, B, C, H, W, Cpad, reinterpret_cast<cudaStream_t>(stream));
}
int hrp_nhwc_bf16_to_nchw_f32(const void* in, float* out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                              void* stream) {
  HRP_REQUIRE(in != nullptr && out != nullptr && B > 0 && Cpad >= C, "bad argument");
  return launch_nhwc_bf16_to_nchw_f32(in, out, B, C, H, W, Cpad, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_robot_create(const hrp_link_row* rows, int32_t n_links, const int32_t* kp_link, const double* kp_offset,
                     int32_t nkpt, int32_t dof, hrp_robot** out) {
  HRP_REQUIRE(rows != nullptr && kp_link != nullptr && kp_offset != nullptr && out != nullptr, "null argument");
  HRP_REQUIRE(n_links > 0 && n_links <= kMaxLinks, "link count out of range (prune the tree to keypoint ancestors)");
  HRP_REQUIRE(nkpt > 0 && nkpt <= kMaxKpt && dof > 0 && dof <= kMaxDof, "keypoint / dof count out of range");
  hrp_robot* r = new (std::nothrow) hrp_robot();
  HRP_REQUIRE(r != nullptr, "out of host memory");
  RobotTable& t = r->host;
  memset(&t, 0, sizeof(t));
  t.n_links = n_links;
  t.nkpt = nkpt;
  t.dof = dof;
  for (int i = 0; i < n_links; ++i) {
    const hrp_link_row& row = rows[i];
    if (!(row.parent < i && row.parent >= -1) || row.jtype < 0 || row.jtype > 2 || row.qcol >= dof ||
        (row.jtype != 0 && row.qcol < 0)) {
      delete r;
      set_error("invalid link row " + std::to_string(i));
      return HRP_ERR_INVALID;
    }
    t.parent[i] = row.parent;
    t.jtype[i] = row.jtype;
    t.qcol[i] = row.qcol;
    t.qmul[i] = (float)row.qmul;
    t.qoff[i] = (float)row.qoff;
    for (int e = 0; e < 12; ++e) t.origin[i][e] = (float)row.origin[e];
    for (int a = 0; a < 3; ++a) t.axis[i][a] = (float)row.axis[a];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) t.axis_outer[i][a * 3 + b] = (float)(row.axis[a] * row.axis[b]);
  }
  for (int k = 0; k < nkpt; ++k) {
    if (kp_link[k] < 0 || kp_link[k] >= n_links) {
      delete r;
      set_error("keypoint link index out of range");
      return HRP_ERR_INVALID;
    }
    t.kp_link[k] = kp_link[k];
    for (int a = 0; a < 3; ++a) t.kp_off[k][a] = (float)kp_offset[k * 3 + a];
  }
  cudaError_t e = cudaMalloc(&r->dev, sizeof(RobotTable));
  if (e == cudaSuccess) e = cudaMemcpy(r->dev, &t, sizeof(RobotTable), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (r->dev) cudaFree(r->dev);
    delete r;
    set_error(std::string("robot table upload failed: ") + cudaGetErrorString(e));
    return HRP_ERR_CUDA;
  }
  *out = r;
  return HRP_OK;
}

// internal accessors for model.cu (not part of include/hrp.h)
const void* hrp_robot_device_table(const hrp_robot* robot) { return robot ? robot->dev : nullptr; }
int hrp_robot_dims(const hrp_robot* robot, int32_t* nkpt, int32_t* dof) {
  *nkpt = robot->host.nkpt;
  *dof = robot->host.dof;
  return HRP_OK;
}

void hrp_robot_destroy(hrp_robot* robot) {
  if (robot == nullptr) return;
  if (robot->dev) cudaFree(robot->dev);
  if (robot->full_rows) cudaFree(robot->full_rows);
  delete robot;
}

int hrp_robot_set_full_tree(hrp_robot* robot, const hrp_link_row* rows, int32_t n_links) {
  HRP_REQUIRE(robot != nullptr && rows != nullptr && n_links > 0 && n_links <= 4096, "bad argument");
  std::vector<LinkRowDev> h((size_t)n_links);
  for (int i = 0; i < n_links; ++i) {
    const hrp_link_row& row = rows[i];
    if (!(row.parent < i && row.parent >= -1) || row.jtype < 0 || row.jtype > 2 || row.qcol >= robot->host.dof ||
        (row.jtype != 0 && row.qcol < 0)) {
      set_error("invalid link row " + std::to_string(i));
      return HRP_ERR_INVALID;
    }
    LinkRowDev& d = h[i];
    memset(&d, 0, sizeof(d));
    d.parent = row.parent;
    d.jtype = row.jtype;
    d.qcol = row.qcol;
    d.qmul = (float)row.qmul;
    d.qoff = (float)row.qoff;
    for (int e = 0; e < 12; ++e) d.origin[e] = (float)row.origin[e];
    for (int a = 0; a < 3; ++a) d.axis[a] = (float)row.axis[a];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) d.axis_outer[a * 3 + b] = (float)(row.axis[a] * row.axis[b]);
  }
  if (robot->full_rows) cudaFree(robot->full_rows);
  robot->full_rows = nullptr;
  robot->n_full = 0;
  HRP_CUDA_CHECK(cudaMalloc(&robot->full_rows, h.size() * sizeof(LinkRowDev)));
  HRP_CUDA_CHECK(cudaMemcpy(robot->full_rows, h.data(), h.size() * sizeof(LinkRowDev), cudaMemcpyHostToDevice));
  robot->n_full = n_links;
  return HRP_OK;
}

int hrp_link_fk(hrp_robot* robot, const float* q, int32_t B, int32_t all_links, float global_scale, float* out_T,
                void* stream) {
  HRP_REQUIRE(robot != nullptr && q != nullptr && out_T != nullptr && B > 0, "bad argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (all_links) {
    if (robot->full_rows == nullptr) {
      set_error("hrp_link_fk(all_links = 1) needs hrp_robot_set_full_tree first");
      return HRP_ERR_STATE;
    }
    return launch_link_fk_all(robot->full_rows, robot->n_full, q, robot->host.dof, B, out_T, s);
  }
  return launch_twl(robot->dev, robot->host.n_links, robot->host.nkpt, q, B, global_scale, out_T, s);
}

int hrp_inv_intrinsics(const float* K, float* Kinv, int32_t B, void* stream) {
  return launch_inv_intrinsics(K, Kinv, B, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_uvd_to_xyz(const float* uvd, const float* Kinv, const float* root_trans, float image_size, float depth_factor,
                   int32_t return_relative, int32_t B, int32_t N, float* xyz, void* stream) {
  return launch_uvd_to_xyz(uvd, Kinv, root_trans, image_size, depth_factor, return_relative, B, N, xyz,
                           reinterpret_cast<cudaStream_t>(stream));
}

int hrp_uvz2xyz_singlepoint(const float* uv, const float* z, const float* K, int32_t B, float* xyz, void* stream) {
  return launch_uvz2xyz(uv, z, K, B, xyz, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_fk(hrp_robot* robot, const float* q, const float* rot, int32_t rot_dim, const float* trans, int32_t root,
           int32_t use_b2c, float* out_xyz, float* out_rot, int32_t B, void* stream) {
  HRP_REQUIRE(robot != nullptr && q != nullptr && B > 0, "bad argument");
  HRP_REQUIRE(out_xyz != nullptr || out_rot != nullptr, "an output is required");
  HRP_REQUIRE(root >= 0 && root < robot->host.nkpt, "root keypoint out of range");
  HRP_REQUIRE(out_rot == nullptr || use_b2c, "rotation output needs the base-to-camera transform");
  if (use_b2c && rot_dim == 9) {
    set_error("9-D (SVD) rotation input is outside the hot path (no shipped config uses it)");
    return HRP_ERR_UNSUPPORTED;
  }
  FkParams p;
  p.B = B;
  p.rot_dim = rot_dim;
  p.root = root;
  p.use_b2c = use_b2c;
  p.q = q;
  p.rot = rot;
  p.trans = trans;
  p.robot = robot->dev;
  p.pts = out_xyz;
  p.rot_out = out_rot;
  return launch_fk(p, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_fk_backward(hrp_robot* robot, const float* q, const float* rot, const float* trans, int32_t root, int32_t use_b2c,
                    const float* grad_xyz, float* grad_q, float* grad_rot, float* grad_trans, int32_t B, void* stream) {
  HRP_REQUIRE(robot != nullptr && q != nullptr && grad_xyz != nullptr && grad_q != nullptr && B > 0, "bad argument");
  HRP_REQUIRE(root >= 0 && root < robot->host.nkpt, "root keypoint out of range");
  FkBwdParams p;
  p.B = B;
  p.root = root;
  p.use_b2c = use_b2c;
  p.q = q;
  p.rot = rot;
  p.trans = trans;
  p.robot = robot->dev;
  p.grad_pts = grad_xyz;
  p.grad_q = grad_q;
  p.grad_rot = grad_rot;
  p.grad_trans = grad_trans;
  return launch_fk_backward(p, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_project_backward(const float* K, const float* pts, const float* grad_uv, float* grad_pts, int32_t B, int32_t N,
                         void* stream) {
  return launch_project_backward(K, pts, grad_uv, grad_pts, B, N, reinterpret_cast<cudaStream_t>(stream));
}

// ---- input pipeline / metrics (eval.cu) ----------------------------------------------------------------
int hrp_crop_resize(const hrp_crop_args* a, void* stream) {
  HRP_REQUIRE(a != nullptr, "null argument");
  HRP_REQUIRE(a->B > 0 && a->frame_h > 0 && a->frame_w > 0, "bad frame shape");
  HRP_REQUIRE(a->out_size > 0 && a->out_size % 4 == 0, "out_size must be a positive multiple of 4");
  HRP_REQUIRE(a->frames != nullptr && a->bbox != nullptr && a->K_in != nullptr, "frames / bbox / K_in are required");
  HRP_REQUIRE(a->out_u8 != nullptr && a->K_out != nullptr, "out_u8 / K_out are required");
  HRP_REQUIRE((a->k_value == nullptr) || (a->k_bbox != nullptr), "k_value needs k_bbox");
  CropParams p;
  p.B = a->B;
  p.frame_h = a->frame_h;
  p.frame_w = a->frame_w;
  p.out_size = a->out_size;
  p.frames = a->frames;
  p.bbox = a->bbox;
  p.K_in = a->K_in;
  p.out_u8 = a->out_u8;
  p.K_out = a->K_out;
  p.k_bbox = a->k_bbox;
  p.k_value = a->k_value;
  p.k_use_crop_K = a->k_use_crop_K;
  return launch_crop_resize(p, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_metrics_workspace_bytes(int32_t B, int32_t nkpt, int32_t dof, int64_t* bytes) {
  HRP_REQUIRE(bytes != nullptr && B > 0 && nkpt > 0 && dof >= 0, "bad argument");
  *bytes = (int64_t)(3 * nkpt + dof) * B * (int64_t)sizeof(float);
  return HRP_OK;
}

int hrp_metrics_batch(const hrp_metrics_args* a, void* stream) {
  HRP_REQUIRE(a != nullptr, "null argument");
  HRP_REQUIRE(a->B > 0 && a->nkpt > 0 && a->nkpt <= 32 && a->dof >= 0 && a->dof <= 32, "bad sizes (nkpt, dof <= 32)");
  HRP_REQUIRE(a->ref_kpt >= 0 && a->ref_kpt < a->nkpt, "reference keypoint out of range");
  HRP_REQUIRE(a->pred_kp3d && a->gt_kp3d && a->gt_kp2d && a->K_original, "keypoints / K are required");
  HRP_REQUIRE(a->per_image && a->dis3d && a->dis2d, "outputs are required");
  HRP_REQUIRE(a->pred_joint == nullptr || (a->gt_joint != nullptr && a->l1_jointerror != nullptr && a->dof > 0),
              "pred_joint needs gt_joint, l1_jointerror and dof");
  int64_t need = 0;
  hrp_metrics_workspace_bytes(a->B, a->nkpt, a->dof, &need);
  HRP_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= need, "workspace too small");
  MetricsParams p;
  p.B = a->B;
  p.nkpt = a->nkpt;
  p.dof = a->dof;
  p.ref_kpt = a->ref_kpt;
  p.drop_last_joint = a->drop_last_joint;
  p.frame_w = a->frame_w;
  p.frame_h = a->frame_h;
  p.pred_kp3d = a->pred_kp3d;
  p.gt_kp3d = a->gt_kp3d;
  p.gt_kp2d = a->gt_kp2d;
  p.K = a->K_original;
  p.pred_joint = a->pred_joint;
  p.gt_joint = a->gt_joint;
  float* ws = static_cast<float*>(a->workspace);
  p.kp_err3d = ws;
  p.kp_err2d = ws + (size_t)a->nkpt * a->B;
  p.kp_valid = ws + (size_t)2 * a->nkpt * a->B;
  p.joint_err = ws + (size_t)3 * a->nkpt * a->B;
  p.per_image = a->per_image;
  p.dis3d = a->dis3d;
  p.dis2d = a->dis2d;
  p.l1_jointerror = a->l1_jointerror;
  return launch_metrics_batch(p, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_metrics_summary(const float* dis3d, const float* dis2d, int64_t n, double* out, void* stream) {
  HRP_REQUIRE(dis3d != nullptr && dis2d != nullptr && out != nullptr && n > 0, "bad argument");
  SummaryParams p;
  p.dis3d = dis3d;
  p.dis2d = dis2d;
  p.n = n;
  // len(np.arange(0, limit, delta)) = ceil(limit / delta) in fp64 (metrics.py:131-134,144-147)
  p.nthr_add = (int)ceil(0.1 / 0.00001);
  p.nthr_pck = (int)ceil(20.0 / 0.01);
  static const double add_mm[8] = {1, 5, 10, 20, 40, 60, 80, 100};
  static const double pck_px[8] = {2.5, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0};
  for (int i = 0; i < 8; ++i) {
    p.table_thr[i] = (float)(add_mm[i] * 1e-3);  // `dis3d <= th_mm * 1e-3`: weak python scalar -> fp32 compare
    p.table_thr[8 + i] = (float)pck_px[i];
  }
  p.out = out;
  return launch_metrics_summary(p, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_pnp(const float* pts2d, const float* pts3d, const float* K, int32_t K_batched, int32_t B, int32_t N, float* pose6,
            float* rot6d, void* stream) {
  HRP_REQUIRE(pts2d != nullptr && pts3d != nullptr && K != nullptr && pose6 != nullptr, "null argument");
  HRP_REQUIRE(B > 0, "empty batch");
  HRP_REQUIRE(N >= 6 && N <= 64, "the linear initialisation needs 6..64 points");
  PnpParams p;
  p.B = B;
  p.N = N;
  p.K_batched = K_batched;
  p.max_iters = 30;
  p.pts2d = pts2d;
  p.pts3d = pts3d;
  p.K = K;
  p.pose6 = pose6;
  p.rot6d = rot6d;
  return launch_pnp(p, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_project(const float* K, const float* pts, float* uv, int32_t B, int32_t N, void* stream) {
  return launch_project(K, pts, uv, B, N, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_head_workspace_bytes(int32_t B, int32_t nkpt, int64_t* bytes) {
  HRP_REQUIRE(B > 0 && nkpt > 0 && nkpt <= kMaxKpt && bytes != nullptr, "bad argument");
  const int chunks = head_default_chunks(B);
  *bytes = (int64_t)(((size_t)B * sizeof(unsigned int) + 255) / 256 * 256 +
                     head_partials_elems(B, nkpt, chunks) * sizeof(float));
  return HRP_OK;
}

int hrp_head(const hrp_head_args* a, void* stream) {
  HRP_REQUIRE(a != nullptr, "null argument");
  HRP_REQUIRE(a->workspace != nullptr, "workspace is required");
  int64_t need = 0;
  int rc = hrp_head_workspace_bytes(a->B, a->nkpt, &need);
  if (rc != HRP_OK) return rc;
  HRP_REQUIRE(a->workspace_bytes >= need, "workspace too small");
  HRP_REQUIRE(a->root_depth != nullptr, "root depth is required");
  HRP_REQUIRE(a->robot == nullptr || a->robot->host.nkpt == a->nkpt, "robot keypoint count mismatch");
  HeadParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B;
  p.nkpt = a->nkpt;
  p.ref_kpt = a->ref_kpt;
  p.fix_root = a->fix_root;
  p.image_size = a->image_size;
  p.depth_factor = a->depth_factor;
  p.heatmap = reinterpret_cast<const bf16*>(a->heatmap);
  p.heatmap_f32 = a->heatmap_fp32 != 0 ? 1 : 0;
  p.chunks = head_default_chunks(a->B);
  p.counters = reinterpret_cast<unsigned int*>(a->workspace);
  p.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(a->workspace) +
                                        ((size_t)a->B * sizeof(unsigned int) + 255) / 256 * 256);
  p.K = a->K;
  p.depth_in = a->root_depth;
  p.pose_in = a->pose;
  p.rot_in = a->rot;
  p.robot = (a->robot != nullptr && a->pose != nullptr && a->rot != nullptr) ? a->robot->dev : nullptr;
  p.trans = a->trans;
  p.root_uv = a->root_uv;
  p.uvd = a->uvd;
  p.xyz_int = a->xyz_int;
  p.xyz_fk = a->xyz_fk;
  p.uv_int = a->uv_int;
  p.uv_fk = a->uv_fk;
  return launch_head(p, reinterpret_cast<cudaStream_t>(stream));
}

int hrp_head_backward_heatmap(const void* heatmap, const float* uvd, const float* grad_uvd, const void* workspace,
                              int64_t workspace_bytes, int32_t B, int32_t nkpt, int32_t ref_kpt, int32_t fix_root,
                              int32_t out_fp32, void* grad_heatmap, void* stream) {
  HRP_REQUIRE(heatmap != nullptr && uvd != nullptr && grad_uvd != nullptr && grad_heatmap != nullptr, "null argument");
  HRP_REQUIRE(workspace != nullptr, "the workspace of the preceding hrp_head call is required");
  int64_t need = 0;
  int rc = hrp_head_workspace_bytes(B, nkpt, &need);
  if (rc != HRP_OK) return rc;
  HRP_REQUIRE(workspace_bytes >= need, "workspace too small");
  const float* partials = reinterpret_cast<const float*>(reinterpret_cast<const char*>(workspace) +
                                                         ((size_t)B * sizeof(unsigned int) + 255) / 256 * 256);
  return launch_head_backward_heatmap(reinterpret_cast<const bf16*>(heatmap), partials, uvd, grad_uvd, B, nkpt, ref_kpt,
                                      fix_root, head_default_chunks(B), (out_fp32 & 1) != 0, grad_heatmap,
                                      reinterpret_cast<cudaStream_t>(stream), (out_fp32 & 2) != 0);
}

int hrp_copy_device(void* dst, const void* src, int64_t bytes, void* stream) {
  HRP_REQUIRE(dst != nullptr && src != nullptr && bytes >= 0, "bad argument");
  HRP_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return HRP_OK;
}

}  // extern "C"
