// Fused head kernel (memory-bound): one streaming pass over the bf16 heatmap logits computes the 3-D
// soft-argmax of every keypoint; the last CTA of each image then finishes the sample on chip: uvd -> xyz,
// root depth (2048-dot), root translation, collapsed pose/rot regressors, URDF forward kinematics and
// projections.  Also the standalone FK / projection / pooling kernels behind URDFRobot and
// point_projection_from_3d (head.h lists the reference lines).
#include "head.h"
#include "launch_count.h"

#include <algorithm>

namespace hrp {

// ------------------------------------------------------------------------------------------------------
// small rigid-transform helpers; a 3x4 matrix lives at T[e*stride], e = row*4+col
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rigid_mul(const float* a, int sa, const float* b, int sb, float* c, int sc) {
  // c = a * b for [R t; 0 0 0 1] matrices (the bottom row contributes exactly as in a full 4x4 product)
  float r[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a0 = a[(i * 4 + 0) * sa], a1 = a[(i * 4 + 1) * sa], a2 = a[(i * 4 + 2) * sa], a3 = a[(i * 4 + 3) * sa];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = a0 * b[(0 * 4 + j) * sb];
      v = fmaf(a1, b[(1 * 4 + j) * sb], v);
      v = fmaf(a2, b[(2 * 4 + j) * sb], v);
      if (j == 3) v += a3;
      r[i * 4 + j] = v;
    }
  }
#pragma unroll
  for (int e = 0; e < 12; ++e) c[e * sc] = r[e];
}

__device__ __forceinline__ void rigid_inverse(const float* a, int sa, float* c) {
  // [R t]^-1 = [R^T, -R^T t]
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) c[i * 4 + j] = a[(j * 4 + i) * sa];
    c[i * 4 + 3] = -(a[(0 * 4 + i) * sa] * a[3 * sa] + a[(1 * 4 + i) * sa] * a[7 * sa] + a[(2 * 4 + i) * sa] * a[11 * sa]);
  }
}

// rot6d -> rotation matrix whose ROWS are x, y, z (geometries.py:100-115)
__device__ __forceinline__ void rot6d_to_rows(const float* r, float* R) {
  const float n1 = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const float x0 = r[0] / n1, x1 = r[1] / n1, x2 = r[2] / n1;
  float z0 = x1 * r[5] - x2 * r[4], z1 = x2 * r[3] - x0 * r[5], z2 = x0 * r[4] - x1 * r[3];
  const float n2 = sqrtf(z0 * z0 + z1 * z1 + z2 * z2);
  z0 /= n2;
  z1 /= n2;
  z2 /= n2;
  const float y0 = z1 * x2 - z2 * x1, y1 = z2 * x0 - z0 * x2, y2 = z0 * x1 - z1 * x0;
  R[0] = x0; R[1] = x1; R[2] = x2;
  R[3] = y0; R[4] = y1; R[5] = y2;
  R[6] = z0; R[7] = z1; R[8] = z2;
}

// quaternion (w,x,y,z) -> rotation matrix (geometries.py:21-41)
__device__ __forceinline__ void quat_to_rows(const float* qv, float* R) {
  const float nrm = sqrtf(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2] + qv[3] * qv[3]) + 1e-9f;
  const float w = qv[0] / nrm, x = qv[1] / nrm, y = qv[2] / nrm, z = qv[3] / nrm;
  const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;     R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2;   R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;     R[8] = w2 - x2 - y2 + z2;
}

// rotation matrix -> quaternion (geometries.py:63-82)
__device__ __forceinline__ void rows_to_quat(const float* R, float* qo) {
  float w = sqrtf(fmaxf(1.0f + R[0] + R[4] + R[8], 0.f)) / 2.0f;
  w = fmaxf(w, 1e-8f);
  const float w4 = 4.0f * w;
  float x = (R[7] - R[5]) / w4, y = (R[2] - R[6]) / w4, z = (R[3] - R[1]) / w4;
  float mag = fmaxf(sqrtf(w * w + x * x + y * y + z * z), 1e-8f);
  qo[0] = w / mag; qo[1] = x / mag; qo[2] = y / mag; qo[3] = z / mag;
}

// Tree forward kinematics: T_link = T_parent * (O_j * M_j(q))  (urdf.py:3116-3140)
__device__ void fk_tree(const RobotTable* __restrict__ rb, const float* __restrict__ q, float* T, int stride) {
  const int nl = rb->n_links;
  for (int i = 0; i < nl; ++i) {
    float* Ti = T + (size_t)i * 12 * stride;
    const int par = rb->parent[i];
    if (par < 0) {
#pragma unroll
      for (int e = 0; e < 12; ++e) Ti[e * stride] = (e == 0 || e == 5 || e == 10) ? 1.f : 0.f;
      continue;
    }
    float child[12];
    const float* O = rb->origin[i];
    const int jt = rb->jtype[i];
    if (jt == 0) {
#pragma unroll
      for (int e = 0; e < 12; ++e) child[e] = O[e];
    } else {
      const float cfg = rb->qmul[i] * q[rb->qcol[i]] + rb->qoff[i];
      if (jt == 1) {  // Rodrigues, urdf.py:2447-2462
        float s, c;
        sincosf(cfg, &s, &c);
        const float* ax = rb->axis[i];
        const float* oo = rb->axis_outer[i];
        const float omc = 1.0f - c;
        float M[9];
        M[0] = (c + oo[0] * omc);
        M[1] = (oo[1] * omc) + (-ax[2]) * s;
        M[2] = (oo[2] * omc) + ax[1] * s;
        M[3] = (oo[3] * omc) + ax[2] * s;
        M[4] = (c + oo[4] * omc);
        M[5] = (oo[5] * omc) + (-ax[0]) * s;
        M[6] = (oo[6] * omc) + (-ax[1]) * s;
        M[7] = (oo[7] * omc) + ax[0] * s;
        M[8] = (c + oo[8] * omc);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            child[r * 4 + cc] = fmaf(O[r * 4 + 2], M[6 + cc], fmaf(O[r * 4 + 1], M[3 + cc], O[r * 4 + 0] * M[cc]));
          child[r * 4 + 3] = O[r * 4 + 3];
        }
      } else {  // prismatic, urdf.py:2388-2390
        const float* ax = rb->axis[i];
        const float t0 = ax[0] * cfg, t1 = ax[1] * cfg, t2 = ax[2] * cfg;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          child[r * 4 + 0] = O[r * 4 + 0];
          child[r * 4 + 1] = O[r * 4 + 1];
          child[r * 4 + 2] = O[r * 4 + 2];
          child[r * 4 + 3] = fmaf(O[r * 4 + 2], t2, fmaf(O[r * 4 + 1], t1, O[r * 4 + 0] * t0)) + O[r * 4 + 3];
        }
      }
    }
    rigid_mul(T + (size_t)par * 12 * stride, stride, child, 1, Ti, stride);
  }
}

// keypoints from link transforms: pts = (b2c * [Troot^-1 *] T_link) applied to the link offset
// (urdf_robot.py:100-104,195-198).  b2c == nullptr -> only_fk variants.
__device__ void fk_keypoints(const RobotTable* __restrict__ rb, const float* T, int stride, const float* b2c, int root,
                             float* __restrict__ pts) {
  float Tinv[12];
  if (root > 0) rigid_inverse(T + (size_t)rb->kp_link[root] * 12 * stride, stride, Tinv);
  for (int k = 0; k < rb->nkpt; ++k) {
    float X[12], Y[12];
    const float* Tl = T + (size_t)rb->kp_link[k] * 12 * stride;
    if (root > 0) {
      rigid_mul(Tinv, 1, Tl, stride, X, 1);
    } else {
#pragma unroll
      for (int e = 0; e < 12; ++e) X[e] = Tl[e * stride];
    }
    if (b2c != nullptr) {
      rigid_mul(b2c, 1, X, 1, Y, 1);
    } else {
#pragma unroll
      for (int e = 0; e < 12; ++e) Y[e] = X[e];
    }
    const float o0 = rb->kp_off[k][0], o1 = rb->kp_off[k][1], o2 = rb->kp_off[k][2];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      pts[k * 3 + r] = fmaf(Y[r * 4 + 2], o2, fmaf(Y[r * 4 + 1], o1, Y[r * 4 + 0] * o0)) + Y[r * 4 + 3];
  }
}

// ------------------------------------------------------------------------------------------------------
// standalone FK kernel: one thread per sample, link transforms in shared memory ([element][thread] layout)
// ------------------------------------------------------------------------------------------------------
constexpr int kFkThreads = 32;

__global__ void __launch_bounds__(kFkThreads) fk_kernel(const FkParams p) {
  extern __shared__ float fk_smem[];
  const int b = blockIdx.x * kFkThreads + threadIdx.x;
  if (b >= p.B) return;
  const RobotTable* rb = p.robot;
  float* T = fk_smem + threadIdx.x;
  float q[kMaxDof];
  for (int i = 0; i < rb->dof; ++i) q[i] = p.q[(size_t)b * rb->dof + i];
  fk_tree(rb, q, T, kFkThreads);
  float b2c[12];
  if (p.use_b2c) {
    float R[9];
    if (p.rot_dim == 6) rot6d_to_rows(p.rot + (size_t)b * 6, R);
    else quat_to_rows(p.rot + (size_t)b * 4, R);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      b2c[r * 4 + 0] = R[r * 3 + 0];
      b2c[r * 4 + 1] = R[r * 3 + 1];
      b2c[r * 4 + 2] = R[r * 3 + 2];
      b2c[r * 4 + 3] = p.trans[(size_t)b * 3 + r];
    }
  }
  if (p.pts != nullptr) {
    float pts[kMaxKpt * 3];
    fk_keypoints(rb, T, kFkThreads, p.use_b2c ? b2c : nullptr, p.root, pts);
    for (int i = 0; i < rb->nkpt * 3; ++i) p.pts[(size_t)b * rb->nkpt * 3 + i] = pts[i];
  }
  if (p.rot_out != nullptr) {  // get_rotation_at_specific_root, urdf_robot.py:113-138: rotation of b2c * T_root
    float Y[12];
    rigid_mul(b2c, 1, T + (size_t)rb->kp_link[p.root] * 12 * kFkThreads, kFkThreads, Y, 1);
    if (p.rot_dim == 6) {
      for (int i = 0; i < 6; ++i) p.rot_out[(size_t)b * 6 + i] = Y[(i / 3) * 4 + (i % 3)];
    } else {
      const float R[9] = {Y[0], Y[1], Y[2], Y[4], Y[5], Y[6], Y[8], Y[9], Y[10]};
      rows_to_quat(R, p.rot_out + (size_t)b * 4);
    }
  }
}

int launch_fk(const FkParams& p, cudaStream_t s) {
  HRP_REQUIRE(p.B > 0 && p.q != nullptr && p.robot != nullptr, "bad FK arguments");
  HRP_REQUIRE(!p.use_b2c || (p.rot != nullptr && p.trans != nullptr), "rotation / translation required");
  HRP_REQUIRE(!p.use_b2c || p.rot_dim == 6 || p.rot_dim == 4, "rotation must be 6-D or a quaternion");
  const int smem = kMaxLinks * 12 * kFkThreads * (int)sizeof(float);
  // (set on every launch: the attribute is per device, and a process may drive several GPUs)
  HRP_CUDA_CHECK(cudaFuncSetAttribute(fk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  fk_kernel<<<(p.B + kFkThreads - 1) / kFkThreads, kFkThreads, smem, s>>>(p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// FK backward (row f4): one thread per sample, link transforms in shared memory like fk_kernel
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// d loss / d q from a force `w` (gradient w.r.t. a base-frame point `x` attached to link `l`): every moving ancestor joint
// contributes  revolute: a . ((x - o) x w),  prismatic: a . w  (a, o = joint axis / origin in the base frame)
__device__ void fk_bwd_point(const RobotTable* __restrict__ rb, const float* T, int stride, int l, const float* x,
                             const float* w, float sign, float* gq) {
  for (int i = l; i >= 0 && rb->parent[i] >= 0; i = rb->parent[i]) {
    const int jt = rb->jtype[i];
    if (jt == 0) continue;
    const float* Ti = T + (size_t)i * 12 * stride;
    const float* ax = rb->axis[i];
    float a[3], d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      a[r] = Ti[(r * 4 + 0) * stride] * ax[0] + Ti[(r * 4 + 1) * stride] * ax[1] + Ti[(r * 4 + 2) * stride] * ax[2];
    float g;
    if (jt == 1) {
#pragma unroll
      for (int r = 0; r < 3; ++r) d[r] = x[r] - Ti[(r * 4 + 3) * stride];
      float c[3];
      cross3(d, w, c);
      g = a[0] * c[0] + a[1] * c[1] + a[2] * c[2];
    } else {
      g = a[0] * w[0] + a[1] * w[1] + a[2] * w[2];
    }
    gq[rb->qcol[i]] += sign * rb->qmul[i] * g;
  }
}

// d loss / d q from a torque `tau` on the rotation of link `l` (d R_l = [a]x R_l dq for every revolute ancestor)
__device__ void fk_bwd_rotation(const RobotTable* __restrict__ rb, const float* T, int stride, int l, const float* tau,
                                float sign, float* gq) {
  for (int i = l; i >= 0 && rb->parent[i] >= 0; i = rb->parent[i]) {
    if (rb->jtype[i] != 1) continue;
    const float* Ti = T + (size_t)i * 12 * stride;
    const float* ax = rb->axis[i];
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
      g += (Ti[(r * 4 + 0) * stride] * ax[0] + Ti[(r * 4 + 1) * stride] * ax[1] + Ti[(r * 4 + 2) * stride] * ax[2]) * tau[r];
    gq[rb->qcol[i]] += sign * rb->qmul[i] * g;
  }
}

__global__ void __launch_bounds__(kFkThreads) fk_bwd_kernel(const FkBwdParams p) {
  extern __shared__ float fk_smem[];
  const int b = blockIdx.x * kFkThreads + threadIdx.x;
  if (b >= p.B) return;
  const RobotTable* rb = p.robot;
  float* T = fk_smem + threadIdx.x;
  const int S = kFkThreads;
  float q[kMaxDof], gq[kMaxDof];
  for (int i = 0; i < rb->dof; ++i) {
    q[i] = p.q[(size_t)b * rb->dof + i];
    gq[i] = 0.f;
  }
  fk_tree(rb, q, T, S);
  float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  if (p.use_b2c) rot6d_to_rows(p.rot + (size_t)b * 6, R);
  // root frame: p_k = Rr^T (x_k - o_r)
  const int lr = (p.root > 0) ? rb->kp_link[p.root] : -1;
  float Rr[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, orr[3] = {0.f, 0.f, 0.f};
  if (lr >= 0) {
    const float* Tr = T + (size_t)lr * 12 * S;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) Rr[r * 3 + c] = Tr[(r * 4 + c) * S];
      orr[r] = Tr[(r * 4 + 3) * S];
    }
  }
  float gt[3] = {0.f, 0.f, 0.f}, gR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float g_or[3] = {0.f, 0.f, 0.f}, tau_r[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < rb->nkpt; ++k) {
    const int l = rb->kp_link[k];
    const float* Tl = T + (size_t)l * 12 * S;
    const float o0 = rb->kp_off[k][0], o1 = rb->kp_off[k][1], o2 = rb->kp_off[k][2];
    float x[3], v[3], pk[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      x[r] = fmaf(Tl[(r * 4 + 2) * S], o2, fmaf(Tl[(r * 4 + 1) * S], o1, Tl[(r * 4 + 0) * S] * o0)) + Tl[(r * 4 + 3) * S];
#pragma unroll
    for (int r = 0; r < 3; ++r) v[r] = x[r] - orr[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) pk[c] = Rr[0 * 3 + c] * v[0] + Rr[1 * 3 + c] * v[1] + Rr[2 * 3 + c] * v[2];  // Rr^T v
    const float* g = p.grad_pts + ((size_t)b * rb->nkpt + k) * 3;
    float gp[3];
    if (p.use_b2c) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        gt[r] += g[r];
#pragma unroll
        for (int c = 0; c < 3; ++c) gR[r * 3 + c] = fmaf(g[r], pk[c], gR[r * 3 + c]);   // out_r = R_row_r . p_k + t_r
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) gp[c] = R[0 * 3 + c] * g[0] + R[1 * 3 + c] * g[1] + R[2 * 3 + c] * g[2];  // R^T g
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) gp[c] = g[c];
    }
    float w[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) w[r] = Rr[r * 3 + 0] * gp[0] + Rr[r * 3 + 1] * gp[1] + Rr[r * 3 + 2] * gp[2];   // Rr gp
    fk_bwd_point(rb, T, S, l, x, w, 1.0f, gq);
    if (lr >= 0) {
      float c[3];
      cross3(v, w, c);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        g_or[r] += w[r];
        tau_r[r] += c[r];
      }
    }
  }
  if (lr >= 0) {
    fk_bwd_point(rb, T, S, lr, orr, g_or, -1.0f, gq);      // the root origin moves with its own ancestors
    fk_bwd_rotation(rb, T, S, lr, tau_r, -1.0f, gq);       // ... and so does the root orientation
  }
  for (int i = 0; i < rb->dof; ++i) p.grad_q[(size_t)b * rb->dof + i] = gq[i];
  if (p.use_b2c) {
    if (p.grad_trans != nullptr)
      for (int r = 0; r < 3; ++r) p.grad_trans[(size_t)b * 3 + r] = gt[r];
    if (p.grad_rot != nullptr) {
      // rows x = a/|a|, z = (x x b)/|x x b|, y = z x x  (geometries.py:100-115); cross adjoints: c = u x v ->
      // gu += v x gc, gv += gc x u
      const float* rr = p.rot + (size_t)b * 6;
      const float a[3] = {rr[0], rr[1], rr[2]}, bb[3] = {rr[3], rr[4], rr[5]};
      const float na = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      const float x[3] = {R[0], R[1], R[2]}, z[3] = {R[6], R[7], R[8]};
      float wv[3];
      cross3(x, bb, wv);
      const float nw = sqrtf(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
      float gx[3] = {gR[0], gR[1], gR[2]}, gy[3] = {gR[3], gR[4], gR[5]}, gz[3] = {gR[6], gR[7], gR[8]};
      float t1[3];
      cross3(x, gy, t1);   // y = z x x : gz += x x gy
      for (int r = 0; r < 3; ++r) gz[r] += t1[r];
      cross3(gy, z, t1);   //            gx += gy x z
      for (int r = 0; r < 3; ++r) gx[r] += t1[r];
      const float zg = z[0] * gz[0] + z[1] * gz[1] + z[2] * gz[2];
      float gw[3];
      for (int r = 0; r < 3; ++r) gw[r] = (gz[r] - z[r] * zg) / nw;   // z = w / |w|
      cross3(bb, gw, t1);  // w = x x b : gx += b x gw
      for (int r = 0; r < 3; ++r) gx[r] += t1[r];
      float gb[3];
      cross3(gw, x, gb);   //             gb = gw x x
      const float xg = x[0] * gx[0] + x[1] * gx[1] + x[2] * gx[2];
      for (int r = 0; r < 3; ++r) {
        p.grad_rot[(size_t)b * 6 + r] = (gx[r] - x[r] * xg) / na;     // x = a / |a|
        p.grad_rot[(size_t)b * 6 + 3 + r] = gb[r];
      }
    }
  }
}

int launch_fk_backward(const FkBwdParams& p, cudaStream_t s) {
  HRP_REQUIRE(p.B > 0 && p.q != nullptr && p.robot != nullptr && p.grad_pts != nullptr && p.grad_q != nullptr,
              "bad FK-backward arguments");
  HRP_REQUIRE(!p.use_b2c || (p.rot != nullptr && p.trans != nullptr), "rotation / translation required");
  const int smem = kMaxLinks * 12 * kFkThreads * (int)sizeof(float);
  HRP_CUDA_CHECK(cudaFuncSetAttribute(fk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  fk_bwd_kernel<<<(p.B + kFkThreads - 1) / kFkThreads, kFkThreads, smem, s>>>(p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// link transforms: URDFRobot.get_TWL (urdf_robot.py:107-111) and URDF.link_fk_batch over ALL links
// (urdf.py:3061-3149) -- one thread per sample
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFkThreads) twl_kernel(const RobotTable* __restrict__ rb, const float* __restrict__ q, int B,
                                                         float scale, float* __restrict__ out) {
  extern __shared__ float fk_smem[];
  const int b = blockIdx.x * kFkThreads + threadIdx.x;
  if (b >= B) return;
  float* T = fk_smem + threadIdx.x;
  float qv[kMaxDof];
  for (int i = 0; i < rb->dof; ++i) qv[i] = q[(size_t)b * rb->dof + i];
  fk_tree(rb, qv, T, kFkThreads);
  for (int k = 0; k < rb->nkpt; ++k) {
    const float* Tl = T + (size_t)rb->kp_link[k] * 12 * kFkThreads;
    float* o = out + ((size_t)b * rb->nkpt + k) * 16;
#pragma unroll
    for (int e = 0; e < 12; ++e) o[e] = ((e & 3) == 3) ? Tl[e * kFkThreads] * scale : Tl[e * kFkThreads];
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
  }
}

int launch_twl(const RobotTable* robot, int n_links, int nkpt, const float* q, int B, float scale, float* out_T,
               cudaStream_t s) {
  HRP_REQUIRE(robot != nullptr && q != nullptr && out_T != nullptr && B > 0, "bad get_TWL arguments");
  (void)n_links; (void)nkpt;
  const int smem = kMaxLinks * 12 * kFkThreads * (int)sizeof(float);
  HRP_CUDA_CHECK(cudaFuncSetAttribute(twl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  twl_kernel<<<(B + kFkThreads - 1) / kFkThreads, kFkThreads, smem, s>>>(robot, q, B, scale, out_T);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

__global__ void link_fk_all_kernel(const LinkRowDev* __restrict__ rows, int n_links, const float* __restrict__ q, int dof,
                                   int B, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* Tb = out + (size_t)b * n_links * 16;
  for (int i = 0; i < n_links; ++i) {
    const LinkRowDev& r = rows[i];
    float* Ti = Tb + (size_t)i * 16;
    Ti[12] = 0.f; Ti[13] = 0.f; Ti[14] = 0.f; Ti[15] = 1.f;
    if (r.parent < 0) {
#pragma unroll
      for (int e = 0; e < 12; ++e) Ti[e] = (e == 0 || e == 5 || e == 10) ? 1.f : 0.f;
      continue;
    }
    float child[12];
    const float* O = r.origin;
    if (r.jtype == 0) {
#pragma unroll
      for (int e = 0; e < 12; ++e) child[e] = O[e];
    } else {
      const float cfg = r.qmul * q[(size_t)b * dof + r.qcol] + r.qoff;
      if (r.jtype == 1) {  // same Rodrigues form as fk_tree (urdf.py:2447-2462)
        float sn, c;
        sincosf(cfg, &sn, &c);
        const float* ax = r.axis;
        const float* oo = r.axis_outer;
        const float omc = 1.0f - c;
        float M[9];
        M[0] = (c + oo[0] * omc);
        M[1] = (oo[1] * omc) + (-ax[2]) * sn;
        M[2] = (oo[2] * omc) + ax[1] * sn;
        M[3] = (oo[3] * omc) + ax[2] * sn;
        M[4] = (c + oo[4] * omc);
        M[5] = (oo[5] * omc) + (-ax[0]) * sn;
        M[6] = (oo[6] * omc) + (-ax[1]) * sn;
        M[7] = (oo[7] * omc) + ax[0] * sn;
        M[8] = (c + oo[8] * omc);
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            child[rr * 4 + cc] = fmaf(O[rr * 4 + 2], M[6 + cc], fmaf(O[rr * 4 + 1], M[3 + cc], O[rr * 4 + 0] * M[cc]));
          child[rr * 4 + 3] = O[rr * 4 + 3];
        }
      } else {
        const float t0 = r.axis[0] * cfg, t1 = r.axis[1] * cfg, t2 = r.axis[2] * cfg;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          child[rr * 4 + 0] = O[rr * 4 + 0];
          child[rr * 4 + 1] = O[rr * 4 + 1];
          child[rr * 4 + 2] = O[rr * 4 + 2];
          child[rr * 4 + 3] = fmaf(O[rr * 4 + 2], t2, fmaf(O[rr * 4 + 1], t1, O[rr * 4 + 0] * t0)) + O[rr * 4 + 3];
        }
      }
    }
    rigid_mul(Tb + (size_t)r.parent * 16, 1, child, 1, Ti, 1);  // (this thread wrote the parent's transform itself)
  }
}

int launch_link_fk_all(const LinkRowDev* rows, int n_links, const float* q, int dof, int B, float* out_T, cudaStream_t s) {
  HRP_REQUIRE(rows != nullptr && q != nullptr && out_T != nullptr && B > 0 && n_links > 0 && dof > 0, "bad link-FK arguments");
  link_fk_all_kernel<<<(B + 63) / 64, 64, 0, s>>>(rows, n_links, q, dof, B, out_T);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// standalone geometry operators (the fused head computes the same expressions in head_finalize)
// ------------------------------------------------------------------------------------------------------
// get_intrinsic_matrix_batch(inv=True) (integral.py:56-73, transforms.py:145-162): divisions in fp64, stored fp32
__global__ void inv_intrinsics_kernel(const float* __restrict__ K, float* __restrict__ Ki, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* k = K + (size_t)b * 9;
  float* o = Ki + (size_t)b * 9;
  const double fx = (double)k[0], fy = (double)k[4], cx = (double)k[2], cy = (double)k[5];
  o[0] = (float)(1.0 / fx); o[1] = 0.f; o[2] = (float)(-cx / fx);
  o[3] = 0.f; o[4] = (float)(1.0 / fy); o[5] = (float)(-cy / fy);
  o[6] = 0.f; o[7] = 0.f; o[8] = 1.f;
}

int launch_inv_intrinsics(const float* K, float* Kinv, int B, cudaStream_t s) {
  HRP_REQUIRE(K != nullptr && Kinv != nullptr && B > 0, "bad intrinsics arguments");
  inv_intrinsics_kernel<<<(B + 127) / 128, 128, 0, s>>>(K, Kinv, B);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// uvd_to_xyz (transforms.py:33-73) with a GENERAL 3x3 inverse intrinsic matrix, as the reference's batched matmul
__global__ void uvd_to_xyz_kernel(const float* __restrict__ uvd, const float* __restrict__ Kinv,
                                  const float* __restrict__ root_trans, float image_size, float depth_factor, int relative,
                                  int B, int N, float* __restrict__ xyz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N;
  const float* k = Kinv + (size_t)b * 9;
  const float u = (uvd[(size_t)i * 3] + 0.5f) * image_size, v = (uvd[(size_t)i * 3 + 1] + 0.5f) * image_size;
  const float dz = uvd[(size_t)i * 3 + 2] * depth_factor;
  const float az = dz + root_trans[(size_t)b * 3 + 2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float h = fmaf(k[r * 3 + 2], 1.0f, fmaf(k[r * 3 + 1], v, k[r * 3] * u));
    h *= az;
    if (relative) h -= root_trans[(size_t)b * 3 + r];
    xyz[(size_t)i * 3 + r] = h;
  }
}

int launch_uvd_to_xyz(const float* uvd, const float* Kinv, const float* root_trans, float image_size, float depth_factor,
                      int return_relative, int B, int N, float* xyz, cudaStream_t s) {
  HRP_REQUIRE(uvd != nullptr && Kinv != nullptr && root_trans != nullptr && xyz != nullptr && B > 0 && N > 0,
              "bad uvd_to_xyz arguments");
  uvd_to_xyz_kernel<<<(B * N + 127) / 128, 128, 0, s>>>(uvd, Kinv, root_trans, image_size, depth_factor, return_relative, B, N,
                                                      xyz);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// uvz2xyz_singlepoint (transforms.py:133-143): xyz = Kinv(K) * [u z, v z, z]
__global__ void uvz2xyz_kernel(const float* __restrict__ uv, const float* __restrict__ z, const float* __restrict__ K, int B,
                               float* __restrict__ xyz) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* k = K + (size_t)b * 9;
  const double fx = (double)k[0], fy = (double)k[4], cx = (double)k[2], cy = (double)k[5];
  const float ifx = (float)(1.0 / fx), ify = (float)(1.0 / fy), icx = (float)(-cx / fx), icy = (float)(-cy / fy);
  const float zz = z[b];
  const float a = uv[(size_t)b * 2] * zz, c = uv[(size_t)b * 2 + 1] * zz;
  xyz[(size_t)b * 3] = fmaf(icx, zz, ifx * a);
  xyz[(size_t)b * 3 + 1] = fmaf(icy, zz, ify * c);
  xyz[(size_t)b * 3 + 2] = zz;
}

int launch_uvz2xyz(const float* uv, const float* z, const float* K, int B, float* xyz, cudaStream_t s) {
  HRP_REQUIRE(uv != nullptr && z != nullptr && K != nullptr && xyz != nullptr && B > 0, "bad uvz2xyz arguments");
  uvz2xyz_kernel<<<(B + 127) / 128, 128, 0, s>>>(uv, z, K, B, xyz);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// pinhole projection: uv = (K p)[:2] / (K p)[2]   (transforms.py:7-21)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void project_point(const float* __restrict__ K, float x, float y, float z, float* uv) {
  const float v0 = fmaf(K[2], z, fmaf(K[1], y, K[0] * x));
  const float v1 = fmaf(K[5], z, fmaf(K[4], y, K[3] * x));
  const float v2 = fmaf(K[8], z, fmaf(K[7], y, K[6] * x));
  uv[0] = v0 / v2;
  uv[1] = v1 / v2;
}

__global__ void project_kernel(const float* __restrict__ K, const float* __restrict__ pts, float* __restrict__ uv, int B,
                               int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N;
  float k[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) k[e] = __ldg(K + (size_t)b * 9 + e);
  float o[2];
  project_point(k, pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2], o);
  uv[(size_t)i * 2] = o[0];
  uv[(size_t)i * 2 + 1] = o[1];
}

int launch_project(const float* K, const float* pts, float* uv, int B, int N, cudaStream_t s) {
  HRP_REQUIRE(K != nullptr && pts != nullptr && uv != nullptr && B > 0 && N > 0, "bad projection arguments");
  project_kernel<<<(B * N + 127) / 128, 128, 0, s>>>(K, pts, uv, B, N);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

__global__ void project_bwd_kernel(const float* __restrict__ K, const float* __restrict__ pts, const float* __restrict__ guv,
                                   float* __restrict__ gpts, int B, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N;
  float k[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) k[e] = __ldg(K + (size_t)b * 9 + e);
  const float x = pts[(size_t)i * 3], y = pts[(size_t)i * 3 + 1], z = pts[(size_t)i * 3 + 2];
  const float v0 = fmaf(k[2], z, fmaf(k[1], y, k[0] * x));
  const float v1 = fmaf(k[5], z, fmaf(k[4], y, k[3] * x));
  const float v2 = fmaf(k[8], z, fmaf(k[7], y, k[6] * x));
  const float u0 = v0 / v2, u1 = v1 / v2;
  const float g0 = guv[(size_t)i * 2] / v2, g1 = guv[(size_t)i * 2 + 1] / v2;
  // d u0 / d p_j = (K0j - u0 K2j) / v2 ,  d u1 / d p_j = (K1j - u1 K2j) / v2
#pragma unroll
  for (int j = 0; j < 3; ++j) gpts[(size_t)i * 3 + j] = g0 * (k[j] - u0 * k[6 + j]) + g1 * (k[3 + j] - u1 * k[6 + j]);
}

int launch_project_backward(const float* K, const float* pts, const float* grad_uv, float* grad_pts, int B, int N,
                            cudaStream_t s) {
  HRP_REQUIRE(K != nullptr && pts != nullptr && grad_uv != nullptr && grad_pts != nullptr && B > 0 && N > 0,
              "bad projection-backward arguments");
  project_bwd_kernel<<<(B * N + 127) / 128, 128, 0, s>>>(K, pts, grad_uv, grad_pts, B, N);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// global average pool of a bf16 NHWC tensor -> fp32 (B,C)   (full_net.py:79,294)
// ------------------------------------------------------------------------------------------------------
__global__ void pool_mean_kernel(const bf16* __restrict__ in, float* __restrict__ out, int HW, int C) {
  const int b = blockIdx.y;
  const int c2 = blockIdx.x * blockDim.x + threadIdx.x;  // channel pair
  if (c2 * 2 >= C) return;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (size_t)b * HW * C) + c2;
  float s0 = 0.f, s1 = 0.f;
  for (int pix = 0; pix < HW; ++pix) {
    const uint32_t v = __ldg(src + (size_t)pix * (C / 2));
    s0 += bf16lo_to_f32(v);
    s1 += bf16hi_to_f32(v);
  }
  const float inv = 1.0f / (float)HW;
  out[(size_t)b * C + 2 * c2] = s0 * inv;
  out[(size_t)b * C + 2 * c2 + 1] = s1 * inv;
}

int launch_pool_mean(const bf16* in, float* out, int B, int HW, int C, cudaStream_t s) {
  HRP_REQUIRE(C % 2 == 0, "pool: even channel count required");
  dim3 grid((C / 2 + 127) / 128, B);
  pool_mean_kernel<<<grid, 128, 0, s>>>(in, out, HW, C);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// standalone root-depth head (RootNet.forward, depth_net.py:121-125): out = (w . feat + b) * k_value * out_scale
// ------------------------------------------------------------------------------------------------------
__global__ void depth_kernel(const float* __restrict__ feat, const float* __restrict__ w, float b, const float* __restrict__ k,
                             float* __restrict__ out, float out_scale) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) acc = fmaf(feat[(size_t)n * 2048 + i], __ldg(w + i), acc);
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  acc += __shfl_xor_sync(0xffffffffu, acc, 8);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    out[n] = ((t + b) * k[n]) * out_scale;
  }
}

int launch_depth(const float* feat, const float* w, float b, const float* k, float* out, int B, float out_scale,
                 cudaStream_t s) {
  depth_kernel<<<B, 256, 0, s>>>(feat, w, b, k, out, out_scale);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// fused head
// ------------------------------------------------------------------------------------------------------
constexpr int kHeadMaxThreads = 320;
constexpr int kHeadMaxSlots = 8;
constexpr float kLog2e = 1.4426950408889634f;

struct HeadSmem {
  float part[kHeadMaxSlots][kMaxKpt][5];
  float uvd[kMaxKpt][3];
  float reg_a[kMaxDof + 6];
  float red[32];
  float T[kMaxLinks * 12];
  float pts[kMaxKpt * 3];
  float depth;
  int is_last;
  // copies of the robot table and of the regressors' small matrices for the single-thread tail of head_finalize: read
  // from global memory they were ~100 dependent first-touch loads on ONE thread (~100 us per sample, on the critical
  // path of every forward)
  RobotTable rb;
  float regA[2][kMaxDof][kMaxDof];
};

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// 2^x for x <= 0 on the MUFU pipe alone: exp2f() brackets the same instruction with a range test and two predicated
// multiplies to keep denormal results (4 issue slots per logit instead of 1 -- the streaming loop is issue-bound, not
// HBM-bound, with them).  Results below 2^-126 flush to zero: < 1.2e-38 of a sum that is >= 1.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__device__ __noinline__ void head_finalize(const HeadParams& p, HeadSmem& sm, int b) {
  const int tid = threadIdx.x;
  const int nk = p.nkpt;
  const RobotTable* rb = p.robot;
  if (rb != nullptr) {  // stage the tables of the serial tail (all threads, coalesced)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(rb);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.rb);
    for (int i = tid; i < (int)(sizeof(RobotTable) / 4); i += blockDim.x) dst[i] = __ldg(src + i);
    if (p.xf != nullptr) {
      for (int i = tid; i < kMaxDof * kMaxDof; i += blockDim.x) {
        (&sm.regA[0][0][0])[i] = __ldg(&p.reg_pose->A[0][0] + i);
        (&sm.regA[1][0][0])[i] = __ldg(&p.reg_rot->A[0][0] + i);
      }
    }
  }
  // (a) merge the per-chunk partial sums of every keypoint -> uvd (integral.py:114-135)
  if (tid < nk) {
    float M = -INFINITY;
    for (int c = 0; c < p.chunks; ++c) M = fmaxf(M, __ldcg(p.partials + (((size_t)b * p.chunks + c) * nk + tid) * 5));
    float S = 0.f, Sx = 0.f, Sy = 0.f, Sz = 0.f;
    for (int c = 0; c < p.chunks; ++c) {
      const float* pp = p.partials + (((size_t)b * p.chunks + c) * nk + tid) * 5;
      const float f = exp2f(__ldcg(pp) - M);
      S = fmaf(__ldcg(pp + 1), f, S);
      Sx = fmaf(__ldcg(pp + 2), f, Sx);
      Sy = fmaf(__ldcg(pp + 3), f, Sy);
      Sz = fmaf(__ldcg(pp + 4), f, Sz);
    }
    float u = (Sx / S) / 64.0f - 0.5f, v = (Sy / S) / 64.0f - 0.5f, d = (Sz / S) / 64.0f - 0.5f;
    if (p.fix_root && tid == p.ref_kpt) d = 0.0f;
    sm.uvd[tid][0] = u;
    sm.uvd[tid][1] = v;
    sm.uvd[tid][2] = d;
  }
  // (b) root depth (full_net.py:271-287)
  if (p.depth_in != nullptr) {
    if (tid == 0) sm.depth = p.depth_in[b];
  } else {
    float acc = 0.f;
    for (int i = tid; i < 2048; i += blockDim.x) acc = fmaf(__ldcg(p.feat + (size_t)b * 2048 + i), __ldg(p.depth_w + i), acc);
    acc = block_sum(acc, sm.red);
    if (tid == 0) sm.depth = ((acc + p.depth_b) * p.k_value[b]) / 1000.0f;
  }
  // (c) collapsed regressors: a = Wx * xf + c  (rows = dof, then 6)
  const int dof = (rb != nullptr) ? rb->dof : 0;
  if (p.xf != nullptr && rb != nullptr) {
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    const float4* x4 = reinterpret_cast<const float4*>(p.xf + (size_t)b * 2048);
    for (int r = warp; r < dof + 6; r += nw) {
      const RegressorTable* rt = (r < dof) ? p.reg_pose : p.reg_rot;
      const int rr = (r < dof) ? r : r - dof;
      const float4* w4 = reinterpret_cast<const float4*>(rt->Wx + (size_t)rr * 2048);
      float acc = 0.f;
      for (int i = lane; i < 512; i += 32) {
        const float4 a = __ldg(w4 + i);
        const float4 x = __ldcg(x4 + i);
        acc = fmaf(a.x, x.x, acc);
        acc = fmaf(a.y, x.y, acc);
        acc = fmaf(a.z, x.z, acc);
        acc = fmaf(a.w, x.w, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      acc += __shfl_xor_sync(0xffffffffu, acc, 8);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (lane == 0) sm.reg_a[r] = acc + rt->c[rr];
    }
  }
  __syncthreads();
  if (tid != 0) return;
  // (d) everything per-sample and tiny: one thread
  const float* K = p.K + (size_t)b * 9;
  const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const float ifx = (float)(1.0 / (double)fx), ify = (float)(1.0 / (double)fy);      // transforms.py:147-162
  const float icx = (float)(-(double)cx / (double)fx), icy = (float)(-(double)cy / (double)fy);
  const float z_root = sm.depth;
  if (p.depth != nullptr) p.depth[b] = z_root;
  for (int k = 0; k < nk; ++k) {  // uvd_to_xyz, transforms.py:33-73
    const float u = sm.uvd[k][0], v = sm.uvd[k][1], d = sm.uvd[k][2];
    const float upx = (u + 0.5f) * p.image_size, vpx = (v + 0.5f) * p.image_size;
    const float az = d * p.depth_factor + z_root;
    const float x = (ifx * upx + icx) * az, y = (ify * vpx + icy) * az, zz = az;
    if (p.uvd != nullptr) {
      p.uvd[((size_t)b * nk + k) * 3 + 0] = u;
      p.uvd[((size_t)b * nk + k) * 3 + 1] = v;
      p.uvd[((size_t)b * nk + k) * 3 + 2] = d;
    }
    if (p.xyz_int != nullptr) {
      p.xyz_int[((size_t)b * nk + k) * 3 + 0] = x;
      p.xyz_int[((size_t)b * nk + k) * 3 + 1] = y;
      p.xyz_int[((size_t)b * nk + k) * 3 + 2] = zz;
    }
    if (p.uv_int != nullptr) project_point(K, x, y, zz, p.uv_int + ((size_t)b * nk + k) * 2);
  }
  // root uv and translation (full_net.py:298,305; transforms.py:133-143)
  const float ru = (sm.uvd[p.ref_kpt][0] + 0.5f) * p.image_size, rv = (sm.uvd[p.ref_kpt][1] + 0.5f) * p.image_size;
  float tr[3];
  tr[0] = ifx * (ru * z_root) + icx * z_root;
  tr[1] = ify * (rv * z_root) + icy * z_root;
  tr[2] = z_root;
  if (p.root_uv != nullptr) {
    p.root_uv[(size_t)b * 2] = ru;
    p.root_uv[(size_t)b * 2 + 1] = rv;
  }
  if (p.trans != nullptr) {
    p.trans[(size_t)b * 3] = tr[0];
    p.trans[(size_t)b * 3 + 1] = tr[1];
    p.trans[(size_t)b * 3 + 2] = tr[2];
  }
  if (rb == nullptr) return;
  // pose / rot (full_net.py:308-378)
  float pose[kMaxDof], rot[6];
  if (p.xf != nullptr) {
    for (int i = 0; i < dof; ++i) pose[i] = p.init_pose[(p.init_batched ? (size_t)b * dof : 0) + i];
    for (int i = 0; i < 6; ++i) rot[i] = p.init_rot[(p.init_batched ? (size_t)b * 6 : 0) + i];
    for (int it = 0; it < p.reg_pose->n_iter; ++it) {
      float nx[kMaxDof];
      for (int i = 0; i < dof; ++i) {
        float dlt = sm.reg_a[i];
        for (int j = 0; j < dof; ++j) dlt = fmaf(sm.regA[0][i][j], pose[j], dlt);
        nx[i] = pose[i] + dlt;
      }
      for (int i = 0; i < dof; ++i) pose[i] = nx[i];
    }
    for (int it = 0; it < p.reg_rot->n_iter; ++it) {
      float nx[6];
      for (int i = 0; i < 6; ++i) {
        float dlt = sm.reg_a[dof + i];
        for (int j = 0; j < 6; ++j) dlt = fmaf(sm.regA[1][i][j], rot[j], dlt);
        nx[i] = rot[i] + dlt;
      }
      for (int i = 0; i < 6; ++i) rot[i] = nx[i];
    }
  } else if (p.pose_in != nullptr) {
    for (int i = 0; i < dof; ++i) pose[i] = p.pose_in[(size_t)b * dof + i];
    for (int i = 0; i < 6; ++i) rot[i] = p.rot_in[(size_t)b * 6 + i];
  } else {
    return;
  }
  if (p.pose != nullptr)
    for (int i = 0; i < dof; ++i) p.pose[(size_t)b * dof + i] = pose[i];
  if (p.rot != nullptr)
    for (int i = 0; i < 6; ++i) p.rot[(size_t)b * 6 + i] = rot[i];
  // FK (full_net.py:380-383)
  fk_tree(&sm.rb, pose, sm.T, 1);
  float R[9], b2c[12];
  rot6d_to_rows(rot, R);
  for (int r = 0; r < 3; ++r) {
    b2c[r * 4 + 0] = R[r * 3 + 0];
    b2c[r * 4 + 1] = R[r * 3 + 1];
    b2c[r * 4 + 2] = R[r * 3 + 2];
    b2c[r * 4 + 3] = tr[r];
  }
  fk_keypoints(&sm.rb, sm.T, 1, b2c, p.ref_kpt, sm.pts);
  for (int k = 0; k < nk; ++k) {
    if (p.xyz_fk != nullptr)
      for (int r = 0; r < 3; ++r) p.xyz_fk[((size_t)b * nk + k) * 3 + r] = sm.pts[k * 3 + r];
    if (p.uv_fk != nullptr)
      project_point(K, sm.pts[k * 3], sm.pts[k * 3 + 1], sm.pts[k * 3 + 2], p.uv_fk + ((size_t)b * nk + k) * 2);
  }
}

// online-softmax state merge: (m, S, Sx, Sy, Sz) <- combine with (m2, ...), all in the log2 domain
__device__ __forceinline__ void softmax_merge(float& m, float& S, float& Sx, float& Sy, float& Sz, float m2, float S2,
                                              float Sx2, float Sy2, float Sz2) {
  const float M = fmaxf(m, m2);
  const float f1 = (m == -INFINITY) ? 0.f : exp2f(m - M);
  const float f2 = (m2 == -INFINITY) ? 0.f : exp2f(m2 - M);
  S = S * f1 + S2 * f2;
  Sx = Sx * f1 + Sx2 * f2;
  Sy = Sy * f1 + Sy2 * f2;
  Sz = Sz * f1 + Sz2 * f2;
  m = M;
}

// Everything after the streaming loop of a head CTA: merge the 8 depth-vector lanes of each keypoint and the pixel slots,
// publish the per-chunk partials, and let the last CTA of the image finish the sample.
__device__ __forceinline__ void head_chunk_finish(const HeadParams& p, HeadSmem& sm, int b, int chunk, int nk, int slots,
                                                  int slot, int k, int vec, bool kvalid, float m, float S, float Sx, float Sy,
                                                  float Sz) {
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off), S2 = __shfl_xor_sync(0xffffffffu, S, off);
    const float Sx2 = __shfl_xor_sync(0xffffffffu, Sx, off), Sy2 = __shfl_xor_sync(0xffffffffu, Sy, off);
    const float Sz2 = __shfl_xor_sync(0xffffffffu, Sz, off);
    softmax_merge(m, S, Sx, Sy, Sz, m2, S2, Sx2, Sy2, Sz2);
  }
  if (vec == 0 && kvalid) {
    float* d = sm.part[slot][k];
    d[0] = m; d[1] = S; d[2] = Sx; d[3] = Sy; d[4] = Sz;
  }
  __syncthreads();
  if ((int)threadIdx.x < nk) {
    const int kk = threadIdx.x;
    float M = -INFINITY, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
    for (int s = 0; s < slots; ++s) {
      const float* pp = sm.part[s][kk];
      softmax_merge(M, a1, a2, a3, a4, pp[0], pp[1], pp[2], pp[3], pp[4]);
    }
    float* dst = p.partials + (((size_t)b * p.chunks + chunk) * nk + kk) * 5;
    __stcg(dst, M);
    __stcg(dst + 1, a1);
    __stcg(dst + 2, a2);
    __stcg(dst + 3, a3);
    __stcg(dst + 4, a4);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(p.counters + b, 1u);
    sm.is_last = (ticket == (unsigned int)p.chunks - 1u);
    if (sm.is_last) p.counters[b] = 0u;  // self-reset for the next launch
  }
  __syncthreads();
  if (!sm.is_last) return;
  __threadfence();
  head_finalize(p, sm, b);
}

// Streaming layout: thread t of a CTA owns the 16-byte vector (t mod vpp) of a pixel row -- a fixed (keypoint, 8-bin
// depth vector) -> 5 fp32 accumulators -- and walks the pixels slot, slot + slots, ... (slot = t / vpp); a warp reads
// 512 contiguous bytes per load instruction.
__global__ void __launch_bounds__(kHeadMaxThreads, 3) head_kernel(const __grid_constant__ HeadParams p) {
  __shared__ HeadSmem sm;
  const int nk = p.nkpt;
  const int vpp = nk * 8;                 // 16-byte vectors per pixel
  // thread -> (pixel slot, vector of the pixel row): consecutive threads take consecutive 16-byte vectors, so every lane
  // of every warp is busy for any keypoint count (a warp-per-4-keypoints mapping left 12 % / 15 % of the lanes idle for
  // 7 / 17 keypoints); vpp is a multiple of 8, so the 8 depth vectors of a keypoint stay in one aligned 8-lane group
  const int slots = (int)blockDim.x / vpp;
  const int slot = (int)threadIdx.x / vpp;
  const int kv = (int)threadIdx.x - slot * vpp;
  const int k = kv >> 3;
  const bool kvalid = (slot < slots);
  const int vec = kv & 7;
  const int b = blockIdx.x / p.chunks, chunk = blockIdx.x - b * p.chunks;
  const int ppc = 4096 / p.chunks;        // pixels per chunk
  const float dbase = (float)(vec * 8);
  const uint4* base =
      reinterpret_cast<const uint4*>(p.heatmap + ((size_t)b * 4096 + (size_t)chunk * ppc) * (size_t)(nk * 64)) + k * 8 + vec;

  float m = -INFINITY, S = 0.f, Sx = 0.f, Sy = 0.f, Sz = 0.f;
  constexpr int UNROLL = 4;
  // one 16-byte vector (8 depth bins of keypoint k at pixel gp) into the running sums; m already bounds its logits
  auto accumulate = [&](const uint4& raw, int gp) {
    const float fw = (float)(gp & 63), fh = (float)(gp >> 6);
    const uint32_t xs[4] = {raw.x, raw.y, raw.z, raw.w};
    float s8 = 0.f, sz8 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float e0 = ex2_ftz(fmaf(bf16lo_to_f32(xs[i]), kLog2e, -m));
      const float e1 = ex2_ftz(fmaf(bf16hi_to_f32(xs[i]), kLog2e, -m));
      s8 += e0 + e1;
      sz8 = fmaf(e0, dbase + (float)(2 * i), sz8);
      sz8 = fmaf(e1, dbase + (float)(2 * i + 1), sz8);
    }
    S += s8;
    Sx = fmaf(s8, fw, Sx);
    Sy = fmaf(s8, fh, Sy);
    Sz += sz8;
  };
  // online-softmax rescale to a new bound (rare after the first few pixels)
  auto raise_max = [&](const __nv_bfloat162 mx) {
    const float vmax = fmaxf(__low2float(mx), __high2float(mx)) * kLog2e;
    if (vmax > m) {
      const float f = exp2f(m - vmax);
      S *= f; Sx *= f; Sy *= f; Sz *= f;
      m = vmax;
    }
  };
  auto vec_max = [](const uint4& raw) {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
    return __hmax2(__hmax2(h2[0], h2[1]), __hmax2(h2[2], h2[3]));
  };
  if (kvalid) {
    // main loop: UNROLL loads in flight per thread, no predicates (whole groups only), pointer-increment addressing,
    // ONE running-max test per group of UNROLL vectors (the loop is issue-bound: every instruction per logit counts)
    const int step = slots * UNROLL;
    const int groups_n = ppc / step;
    const size_t ustride = (size_t)slots * vpp;
    const uint4* ptr = base + (size_t)slot * vpp;
    int gp = chunk * ppc + slot;
    for (int it = 0; it < groups_n; ++it) {
      uint4 raw[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) raw[u] = ld_stream(ptr + u * ustride);
      ptr += (size_t)step * vpp;
      __nv_bfloat162 mx = vec_max(raw[0]);
#pragma unroll
      for (int u = 1; u < UNROLL; ++u) mx = __hmax2(mx, vec_max(raw[u]));
      raise_max(mx);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) accumulate(raw[u], gp + u * slots);
      gp += step;
    }
    // ragged tail (slots * UNROLL does not divide the chunk, e.g. 3 slots): one vector at a time
    for (int pix = groups_n * step + slot; pix < ppc; pix += slots) {
      const uint4 raw = ld_stream(base + (size_t)pix * vpp);
      raise_max(vec_max(raw));
      accumulate(raw, chunk * ppc + pix);
    }
  }
  head_chunk_finish(p, sm, b, chunk, nk, slots, slot, k, vec, kvalid, m, S, Sx, Sy, Sz);
}

// The same kernel for fp32 logits (HeadParams::heatmap_f32: the standalone operator on a caller's fp32 heatmap, no bf16
// rounding of the logits): a thread's 8 depth bins are two 16-byte vectors, otherwise the identical mapping, order of
// operations and tail.  Twice the bytes per logit, so it is HBM-bound at twice the time; the network path never uses it.
__global__ void __launch_bounds__(kHeadMaxThreads, 3) head_kernel_f32(const __grid_constant__ HeadParams p) {
  __shared__ HeadSmem sm;
  const int nk = p.nkpt;
  const int vpp = nk * 8;
  const int slots = (int)blockDim.x / vpp;
  const int slot = (int)threadIdx.x / vpp;
  const int kv = (int)threadIdx.x - slot * vpp;
  const int k = kv >> 3;
  const bool kvalid = (slot < slots);
  const int vec = kv & 7;
  const int b = blockIdx.x / p.chunks, chunk = blockIdx.x - b * p.chunks;
  const int ppc = 4096 / p.chunks;
  const float dbase = (float)(vec * 8);
  const uint4* base = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.heatmap) +
                                                     ((size_t)b * 4096 + (size_t)chunk * ppc) * (size_t)(nk * 64)) +
                      (size_t)(k * 8 + vec) * 2;
  const size_t pstride = (size_t)vpp * 2;   // 16-byte vectors per pixel
  float m = -INFINITY, S = 0.f, Sx = 0.f, Sy = 0.f, Sz = 0.f;
  if (kvalid) {
    constexpr int UNROLL = 2;
    int pix = slot;
    for (; pix + (UNROLL - 1) * slots < ppc; pix += UNROLL * slots) {
      uint4 raw[UNROLL][2];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        raw[u][0] = ld_stream(base + (size_t)(pix + u * slots) * pstride);
        raw[u][1] = ld_stream(base + (size_t)(pix + u * slots) * pstride + 1);
      }
      float vmax = -INFINITY;
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const uint32_t xs[8] = {raw[u][0].x, raw[u][0].y, raw[u][0].z, raw[u][0].w, raw[u][1].x, raw[u][1].y, raw[u][1].z, raw[u][1].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) vmax = fmaxf(vmax, __uint_as_float(xs[i]));
      }
      vmax *= kLog2e;
      if (vmax > m) {
        const float f = exp2f(m - vmax);
        S *= f; Sx *= f; Sy *= f; Sz *= f;
        m = vmax;
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int gp = chunk * ppc + pix + u * slots;
        const float fw = (float)(gp & 63), fh = (float)(gp >> 6);
        const uint32_t xs[8] = {raw[u][0].x, raw[u][0].y, raw[u][0].z, raw[u][0].w, raw[u][1].x, raw[u][1].y, raw[u][1].z, raw[u][1].w};
        float s8 = 0.f, sz8 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e = ex2_ftz(fmaf(__uint_as_float(xs[i]), kLog2e, -m));
          s8 += e;
          sz8 = fmaf(e, dbase + (float)i, sz8);
        }
        S += s8;
        Sx = fmaf(s8, fw, Sx);
        Sy = fmaf(s8, fh, Sy);
        Sz += sz8;
      }
    }
    for (; pix < ppc; pix += slots) {   // ragged tail
      const uint4 r0 = ld_stream(base + (size_t)pix * pstride), r1 = ld_stream(base + (size_t)pix * pstride + 1);
      const uint32_t xs[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      float vmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) vmax = fmaxf(vmax, __uint_as_float(xs[i]));
      vmax *= kLog2e;
      if (vmax > m) {
        const float f = exp2f(m - vmax);
        S *= f; Sx *= f; Sy *= f; Sz *= f;
        m = vmax;
      }
      const int gp = chunk * ppc + pix;
      const float fw = (float)(gp & 63), fh = (float)(gp >> 6);
      float s8 = 0.f, sz8 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float e = ex2_ftz(fmaf(__uint_as_float(xs[i]), kLog2e, -m));
        s8 += e;
        sz8 = fmaf(e, dbase + (float)i, sz8);
      }
      S += s8;
      Sx = fmaf(s8, fw, Sx);
      Sy = fmaf(s8, fh, Sy);
      Sz += sz8;
    }
  }
  head_chunk_finish(p, sm, b, chunk, nk, slots, slot, k, vec, kvalid, m, S, Sx, Sy, Sz);
}

// ------------------------------------------------------------------------------------------------------
// finish from partials written by the final conv's epilogue (soft-argmax fold): one CTA per image
// ------------------------------------------------------------------------------------------------------
constexpr int kFoldThreads = 256;

__global__ void __launch_bounds__(kFoldThreads) head_from_partials_kernel(const __grid_constant__ HeadParams p) {
  __shared__ HeadSmem sm;
  __shared__ float red[kFoldThreads / 32][5];
  const int b = blockIdx.x;
  const int nk = p.nkpt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < nk; ++k) {
    // chunk-parallel merge in a fixed order: strided per-thread pass, warp butterfly, then the 8 warps in order
    float m = -INFINITY, S = 0.f, Sx = 0.f, Sy = 0.f, Sz = 0.f;
    for (int c = threadIdx.x; c < p.chunks; c += kFoldThreads) {
      const float* pp = p.partials + (((size_t)b * p.chunks + c) * nk + k) * 5;
      softmax_merge(m, S, Sx, Sy, Sz, __ldcg(pp), __ldcg(pp + 1), __ldcg(pp + 2), __ldcg(pp + 3), __ldcg(pp + 4));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, off), S2 = __shfl_xor_sync(0xffffffffu, S, off);
      const float Sx2 = __shfl_xor_sync(0xffffffffu, Sx, off), Sy2 = __shfl_xor_sync(0xffffffffu, Sy, off);
      const float Sz2 = __shfl_xor_sync(0xffffffffu, Sz, off);
      softmax_merge(m, S, Sx, Sy, Sz, m2, S2, Sx2, Sy2, Sz2);
    }
    if (lane == 0) {
      red[warp][0] = m; red[warp][1] = S; red[warp][2] = Sx; red[warp][3] = Sy; red[warp][4] = Sz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float M = -INFINITY, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
      for (int w = 0; w < kFoldThreads / 32; ++w) softmax_merge(M, a1, a2, a3, a4, red[w][0], red[w][1], red[w][2], red[w][3], red[w][4]);
      float* dst = p.merged + ((size_t)b * nk + k) * 5;
      __stcg(dst, M);
      __stcg(dst + 1, a1);
      __stcg(dst + 2, a2);
      __stcg(dst + 3, a3);
      __stcg(dst + 4, a4);
    }
    __syncthreads();
  }
  __threadfence_block();
  HeadParams q = p;          // head_finalize merges `chunks` partials per keypoint: hand it the single merged one
  q.partials = p.merged;
  q.chunks = 1;
  head_finalize(q, sm, b);
}

int launch_head_from_partials(const HeadParams& p, cudaStream_t s) {
  HRP_REQUIRE(p.B > 0 && p.nkpt > 0 && p.nkpt <= kMaxKpt, "bad head dims");
  HRP_REQUIRE(p.K != nullptr && p.partials != nullptr && p.merged != nullptr && p.chunks > 0, "head: null tensor");
  HRP_REQUIRE(p.depth_in != nullptr || (p.feat != nullptr && p.depth_w != nullptr && p.k_value != nullptr),
              "head: a root-depth source is required");
  HRP_REQUIRE(p.ref_kpt >= 0 && p.ref_kpt < p.nkpt, "reference keypoint out of range");
  head_from_partials_kernel<<<p.B, kFoldThreads, 0, s>>>(p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

// ------------------------------------------------------------------------------------------------------
// backward of the heatmap integral w.r.t. the logits (first piece)
// ------------------------------------------------------------------------------------------------------
// u = sum_i p_i w_i / 64 - 0.5 with p = softmax over the 64^3 voxels of a keypoint (integral.py:97-135; the `/ sum`
// renormalisation of the resnet branch has zero net Jacobian at sum = 1), so
//   dL/dx_i = p_i * (g_u (w_i - E[w]) + g_v (h_i - E[h]) + g_d (d_i - E[d])) / 64.
// The softmax statistics are NOT recomputed: the forward left the per-chunk (max, sum) partials in the workspace.
// Same thread -> (pixel slot, 16-byte vector) mapping as the forward; one heatmap read, one gradient write.
struct HeadBwdParams {
  int B, nkpt, ref_kpt, fix_root, chunks;
  const bf16* heatmap;
  const float* partials;
  const float* uvd;
  const float* grad_uvd;
  void* grad_out;
};

template <bool F32, bool F32IN>
__global__ void __launch_bounds__(kHeadMaxThreads) head_bwd_heatmap_kernel(const HeadBwdParams p) {
  const int nk = p.nkpt;
  const int vpp = nk * 8;
  const int slots = (int)blockDim.x / vpp;
  const int slot = (int)threadIdx.x / vpp;
  if (slot >= slots) return;
  const int kv = (int)threadIdx.x - slot * vpp;
  const int k = kv >> 3, vec = kv & 7;
  const int b = blockIdx.x / p.chunks, chunk = blockIdx.x - b * p.chunks;
  const int ppc = 4096 / p.chunks;
  // global softmax statistics of (b, k) from the forward's per-chunk partials
  float M = -INFINITY;
  for (int c = 0; c < p.chunks; ++c) M = fmaxf(M, __ldg(p.partials + (((size_t)b * p.chunks + c) * nk + k) * 5));
  float S = 0.f;
  for (int c = 0; c < p.chunks; ++c) {
    const float* pp = p.partials + (((size_t)b * p.chunks + c) * nk + k) * 5;
    S = fmaf(__ldg(pp + 1), exp2f(__ldg(pp) - M), S);
  }
  const float invS = 1.0f / S;
  const float* gq = p.grad_uvd + ((size_t)b * nk + k) * 3;
  const float* uq = p.uvd + ((size_t)b * nk + k) * 3;
  const bool fixed_d = (p.fix_root != 0) && (k == p.ref_kpt);   // uvd[:, ref, 2] = 0 (integral.py:134): no gradient
  const float gu = gq[0] * (1.0f / 64.0f), gv = gq[1] * (1.0f / 64.0f), gd = fixed_d ? 0.0f : gq[2] * (1.0f / 64.0f);
  const float Ew = (uq[0] + 0.5f) * 64.0f, Eh = (uq[1] + 0.5f) * 64.0f, Ed = fixed_d ? 0.0f : (uq[2] + 0.5f) * 64.0f;
  const float dbase = (float)(vec * 8) - Ed;
  const size_t row0 = ((size_t)b * 4096 + (size_t)chunk * ppc);
  const uint4* src = F32IN ? reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.heatmap) + row0 * (size_t)(nk * 64)) + 2 * kv
                           : reinterpret_cast<const uint4*>(p.heatmap + row0 * (size_t)(nk * 64)) + kv;
  const size_t pstride = F32IN ? (size_t)vpp * 2 : (size_t)vpp;
  for (int pix = slot; pix < ppc; pix += slots) {
    const int gp = chunk * ppc + pix;
    const float cwh = fmaf(gu, (float)(gp & 63) - Ew, gv * ((float)(gp >> 6) - Eh));
    float g[8];
    if (F32IN) {
      const uint4 r0 = ld_stream(src + (size_t)pix * pstride), r1 = ld_stream(src + (size_t)pix * pstride + 1);
      const uint32_t xs[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
        g[i] = ex2_ftz(fmaf(__uint_as_float(xs[i]), kLog2e, -M)) * invS * fmaf(gd, dbase + (float)i, cwh);
    } else {
      const uint4 raw = ld_stream(src + (size_t)pix * pstride);
      const uint32_t xs[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p0 = ex2_ftz(fmaf(bf16lo_to_f32(xs[i]), kLog2e, -M)) * invS;
        const float p1 = ex2_ftz(fmaf(bf16hi_to_f32(xs[i]), kLog2e, -M)) * invS;
        g[2 * i] = p0 * fmaf(gd, dbase + (float)(2 * i), cwh);
        g[2 * i + 1] = p1 * fmaf(gd, dbase + (float)(2 * i + 1), cwh);
      }
    }
    const size_t vidx = (row0 + pix) * (size_t)vpp + kv;  // index of this 8-logit vector
    if (F32) {
      float4* dst = reinterpret_cast<float4*>(p.grad_out) + 2 * vidx;
      dst[0] = make_float4(g[0], g[1], g[2], g[3]);
      dst[1] = make_float4(g[4], g[5], g[6], g[7]);
    } else {
      uint4 o;
      o.x = pack_bf16x2(g[0], g[1]);
      o.y = pack_bf16x2(g[2], g[3]);
      o.z = pack_bf16x2(g[4], g[5]);
      o.w = pack_bf16x2(g[6], g[7]);
      reinterpret_cast<uint4*>(p.grad_out)[vidx] = o;
    }
  }
}

int launch_head_backward_heatmap(const bf16* heatmap, const float* partials, const float* uvd, const float* grad_uvd, int B,
                                 int nkpt, int ref_kpt, int fix_root, int chunks, bool out_fp32, void* grad_out,
                                 cudaStream_t s, bool in_fp32) {
  HRP_REQUIRE(B > 0 && nkpt > 0 && nkpt <= kMaxKpt && chunks > 0 && 4096 % chunks == 0, "bad head-backward dims");
  HRP_REQUIRE(ref_kpt >= 0 && ref_kpt < nkpt, "reference keypoint out of range");
  const int vpp = nkpt * 8;
  const int slots = std::max(1, std::min(4, kHeadMaxThreads / vpp));
  const int threads = (slots * vpp + 31) / 32 * 32;
  HeadBwdParams p;
  p.B = B;
  p.nkpt = nkpt;
  p.ref_kpt = ref_kpt;
  p.fix_root = fix_root;
  p.chunks = chunks;
  p.heatmap = heatmap;
  p.partials = partials;
  p.uvd = uvd;
  p.grad_uvd = grad_uvd;
  p.grad_out = grad_out;
  if (in_fp32) {
    HRP_REQUIRE(out_fp32, "fp32 logits give an fp32 gradient");
    head_bwd_heatmap_kernel<true, true><<<B * chunks, threads, 0, s>>>(p);
  } else if (out_fp32) {
    head_bwd_heatmap_kernel<true, false><<<B * chunks, threads, 0, s>>>(p);
  } else {
    head_bwd_heatmap_kernel<false, false><<<B * chunks, threads, 0, s>>>(p);
  }
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

int head_default_chunks(int B) {
  // enough CTAs to cover 148 SMs x ~8 resident CTAs a few times over, power of two <= 64
  int chunks = 64;
  while (chunks > 4 && (long)B * chunks > 148L * 8 * 4) chunks >>= 1;
  return chunks;
}

size_t head_partials_elems(int B, int nkpt, int chunks) { return (size_t)B * chunks * nkpt * 5; }

int launch_head(const HeadParams& p, cudaStream_t s) {
  HRP_REQUIRE(p.B > 0 && p.nkpt > 0 && p.nkpt <= kMaxKpt, "bad head dims");
  HRP_REQUIRE(p.heatmap != nullptr && p.K != nullptr && p.partials != nullptr && p.counters != nullptr,
              "head: null tensor");
  HRP_REQUIRE(p.chunks > 0 && 4096 % p.chunks == 0, "chunks must divide 4096");
  HRP_REQUIRE(p.depth_in != nullptr || (p.feat != nullptr && p.depth_w != nullptr && p.k_value != nullptr),
              "head: a root-depth source is required");
  HRP_REQUIRE(p.ref_kpt >= 0 && p.ref_kpt < p.nkpt, "reference keypoint out of range");
  const int vpp = p.nkpt * 8;
  const int slots = std::max(1, std::min(4, kHeadMaxThreads / vpp));
  const int threads = (slots * vpp + 31) / 32 * 32;
  HRP_REQUIRE(vpp <= kHeadMaxThreads && threads <= kHeadMaxThreads && slots <= kHeadMaxSlots,
              "too many keypoints for the head kernel");
  if (p.heatmap_f32) head_kernel_f32<<<p.B * p.chunks, threads, 0, s>>>(p);
  else head_kernel<<<p.B * p.chunks, threads, 0, s>>>(p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

}  // namespace hrp
