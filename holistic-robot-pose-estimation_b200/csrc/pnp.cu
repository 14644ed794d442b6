// Batched perspective-n-point: pose of a known 3-D keypoint set from its 2-D projections.
// Replaces BPnP_m3d.forward (lib/utils/BPnP.py:114-152): per sample `cv2.solvePnP(EPNP)` as the initial guess, then
// `cv2.solvePnP(ITERATIVE, useExtrinsicGuess)` -- i.e. the minimiser of the pixel reprojection error reached from a
// linear initialisation.  Here: one warp per sample, fp64 --
//   1. 32 starts (one per lane): a fixed rotation + the translation solving the linear collinearity equations for it;
//   2. 12 Levenberg-Marquardt iterations per lane on the 6-vector (left-multiplicative rotation update);
//   3. warp argmin of the reprojection error, winner polished to convergence;
//   4. angle-axis + translation in fp32, and the 6-D rotation the caller derives from it
//      (scripts/test.py:122-124: angle_axis_to_rotation_matrix -> rotmat_to_rot6d, geometries.py:164-232,117-132).
// Where OpenCV reaches the global minimum (every fixture), both agree to ~1e-9 before the fp32 rounding of the outputs
// (tests/test_eval_gpu.py).  Latency-bound, ~1 KB per sample.
#include "eval.h"
#include "launch_count.h"


namespace hrp {

namespace {

__device__ void rodrigues(const double* r, double* R) {
  const double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double th = sqrt(th2);
  double a, b;  // R = I + a [r]x + b [r]x^2
  if (th < 1e-8) {
    a = 1.0 - th2 / 6.0;
    b = 0.5 - th2 / 24.0;
  } else {
    a = sin(th) / th;
    b = (1.0 - cos(th)) / th2;
  }
  const double x = r[0], y = r[1], z = r[2];
  R[0] = 1.0 - b * (y * y + z * z);
  R[1] = -a * z + b * x * y;
  R[2] = a * y + b * x * z;
  R[3] = a * z + b * x * y;
  R[4] = 1.0 - b * (x * x + z * z);
  R[5] = -a * x + b * y * z;
  R[6] = -a * y + b * x * z;
  R[7] = a * x + b * y * z;
  R[8] = 1.0 - b * (x * x + y * y);
}

// rotation matrix -> angle-axis, robust near 0 and pi
__device__ void log_so3(const double* R, double* r) {
  const double tr = R[0] + R[4] + R[8];
  const double wx = R[7] - R[5], wy = R[2] - R[6], wz = R[3] - R[1];  // 2 sin(th) * axis
  const double s = 0.5 * sqrt(wx * wx + wy * wy + wz * wz);           // sin(th)
  const double c = fmin(fmax(0.5 * (tr - 1.0), -1.0), 1.0);
  const double th = atan2(s, c);
  if (s > 1e-6) {
    const double k = th / (2.0 * s);
    r[0] = k * wx;
    r[1] = k * wy;
    r[2] = k * wz;
  } else if (c > 0.0) {  // th ~ 0
    r[0] = 0.5 * wx;
    r[1] = 0.5 * wy;
    r[2] = 0.5 * wz;
  } else {  // th ~ pi: axis from the diagonal of (R + I) / 2, signs from the off-diagonal sums
    double ax = sqrt(fmax(0.5 * (R[0] + 1.0), 0.0)), ay = sqrt(fmax(0.5 * (R[4] + 1.0), 0.0)),
           az = sqrt(fmax(0.5 * (R[8] + 1.0), 0.0));
    if (ax >= ay && ax >= az) {
      ay = copysign(ay, R[1] + R[3]);
      az = copysign(az, R[2] + R[6]);
    } else if (ay >= az) {
      ax = copysign(ax, R[1] + R[3]);
      az = copysign(az, R[5] + R[7]);
    } else {
      ax = copysign(ax, R[2] + R[6]);
      ay = copysign(ay, R[5] + R[7]);
    }
    // orient with the (tiny) antisymmetric part when it is there
    if (ax * wx + ay * wy + az * wz < 0.0) {
      ax = -ax;
      ay = -ay;
      az = -az;
    }
    r[0] = th * ax;
    r[1] = th * ay;
    r[2] = th * az;
  }
}

__device__ bool solve6(double* A, double* b) {  // Gaussian elimination with partial pivoting, A 6x6 row-major; b <- x
  for (int i = 0; i < 6; ++i) {
    int piv = i;
    for (int r = i + 1; r < 6; ++r)
      if (fabs(A[r * 6 + i]) > fabs(A[piv * 6 + i])) piv = r;
    if (fabs(A[piv * 6 + i]) < 1e-300) return false;
    if (piv != i) {
      for (int k = 0; k < 6; ++k) {
        const double t = A[i * 6 + k];
        A[i * 6 + k] = A[piv * 6 + k];
        A[piv * 6 + k] = t;
      }
      const double t = b[i];
      b[i] = b[piv];
      b[piv] = t;
    }
    for (int r = i + 1; r < 6; ++r) {
      const double f = A[r * 6 + i] / A[i * 6 + i];
      for (int k = i; k < 6; ++k) A[r * 6 + k] -= f * A[i * 6 + k];
      b[r] -= f * b[i];
    }
  }
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < 6; ++k) s -= A[i * 6 + k] * b[k];
    b[i] = s / A[i * 6 + i];
  }
  return true;
}

}  // namespace

// cost (sum of squared pixel residuals) and, if H != nullptr, the Gauss-Newton system of the pose (R, t);
// +inf when a point is not in front of the camera
__device__ double pnp_system(const double* R, const double* t, const double* P3, const double* P2, int n, double fx, double fy,
                             double cx, double cy, double* H, double* g) {
  if (H != nullptr) {
    for (int k = 0; k < 36; ++k) H[k] = 0.0;
    for (int k = 0; k < 6; ++k) g[k] = 0.0;
  }
  double cost = 0.0;
  for (int i = 0; i < n; ++i) {
    const double X = P3[i * 3], Y = P3[i * 3 + 1], Z = P3[i * 3 + 2];
    const double rx = R[0] * X + R[1] * Y + R[2] * Z, ry = R[3] * X + R[4] * Y + R[5] * Z, rz = R[6] * X + R[7] * Y + R[8] * Z;
    const double xc = rx + t[0], yc = ry + t[1], zc = rz + t[2];
    if (!(zc > 1e-6)) return INFINITY;
    const double iz = 1.0 / zc;
    const double eu = fx * xc * iz + cx - P2[i * 2], ev = fy * yc * iz + cy - P2[i * 2 + 1];
    cost += eu * eu + ev * ev;
    if (H == nullptr) continue;
    // d(u,v)/d(Xc) = [a00 0 a02; 0 a11 a12];  d(Xc)/d(omega) = -[R X]x (left-multiplicative update), d(Xc)/dt = I
    const double a00 = fx * iz, a02 = -fx * xc * iz * iz, a11 = fy * iz, a12 = -fy * yc * iz * iz;
    const double Ju[6] = {a02 * ry, a00 * rz - a02 * rx, -a00 * ry, a00, 0.0, a02};
    const double Jv[6] = {-a11 * rz + a12 * ry, -a12 * rx, a11 * rx, 0.0, a11, a12};
    for (int r = 0; r < 6; ++r) {
      g[r] += Ju[r] * eu + Jv[r] * ev;
      for (int c = r; c < 6; ++c) H[r * 6 + c] += Ju[r] * Ju[c] + Jv[r] * Jv[c];
    }
  }
  if (H != nullptr)
    for (int r = 1; r < 6; ++r)
      for (int c = 0; c < r; ++c) H[r * 6 + c] = H[c * 6 + r];
  return cost;
}

// Levenberg-Marquardt from (R, t): multiplicative damping of the diagonal, step accepted only if the cost drops
__device__ double pnp_lm(double* R, double* t, const double* P3, const double* P2, int n, double fx, double fy, double cx,
                         double cy, int iters, double lam) {
  double H[36], g[6];
  double cost = pnp_system(R, t, P3, P2, n, fx, fy, cx, cy, H, g);
  if (!(cost < INFINITY)) return INFINITY;
  for (int it = 0; it < iters; ++it) {
    double A[36], d[6];
    for (int k = 0; k < 36; ++k) A[k] = H[k];
    for (int k = 0; k < 6; ++k) {
      A[k * 6 + k] += lam * H[k * 6 + k] + 1e-12;
      d[k] = -g[k];
    }
    if (!solve6(A, d)) break;
    double dR[9], Rn[9], tn[3];
    rodrigues(d, dR);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Rn[i * 3 + j] = dR[i * 3] * R[j] + dR[i * 3 + 1] * R[3 + j] + dR[i * 3 + 2] * R[6 + j];
    for (int k = 0; k < 3; ++k) tn[k] = t[k] + d[3 + k];
    const double cn = pnp_system(Rn, tn, P3, P2, n, fx, fy, cx, cy, nullptr, nullptr);
    if (cn < cost) {
      for (int k = 0; k < 9; ++k) R[k] = Rn[k];
      for (int k = 0; k < 3; ++k) t[k] = tn[k];
      lam = fmax(lam * 0.2, 1e-9);
      double nd = 0.0;
      for (int k = 0; k < 6; ++k) nd += d[k] * d[k];
      cost = pnp_system(R, t, P3, P2, n, fx, fy, cx, cy, H, g);
      if (nd < 1e-26) break;
    } else {
      lam = fmin(lam * 10.0, 1e6);
    }
  }
  return cost;
}

// start rotation of lane l: l < 24 -> the l-th proper signed axis permutation (the rotation group of the cube), else
// one of 8 fixed generic rotations
__device__ void pnp_start_rotation(int l, double* R) {
  if (l < 24) {
    const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    const int parity[6] = {0, 1, 1, 0, 0, 1};  // odd permutations need an odd number of -1 entries for det = +1
    const int pi = l >> 2, k = l & 3;
    // the four sign patterns with the required parity: enumerate s0, s1 freely, s2 fixes the parity
    const int s0 = k & 1, s1 = (k >> 1) & 1, s2 = (s0 ^ s1 ^ parity[pi]);
    const int sg[3] = {s0, s1, s2};
    for (int i = 0; i < 9; ++i) R[i] = 0.0;
    for (int i = 0; i < 3; ++i) R[i * 3 + perms[pi][i]] = sg[i] ? -1.0 : 1.0;
  } else {
    const double extra[8][3] = {{0.9, 0.4, -0.3},  {-0.5, 1.1, 0.7},  {1.3, -1.2, 0.6},  {-1.0, -0.9, -1.4},
                                {0.3, 2.0, -1.1},  {2.1, 0.5, 1.2},   {-1.7, 1.5, -0.8}, {0.6, -2.2, -1.5}};
    rodrigues(extra[l - 24], R);
  }
}

constexpr int kPnpWarps = 4;
constexpr int kPnpMaxPts = 64;

// One warp per sample, one lane per start.  The linear (DLT / EPnP-style) initialisations are degenerate for the nearly
// planar keypoint sets a serial arm produces (measured: the DLT start lands in a wrong basin for 6 of 32 Kuka samples),
// so the initial guess is replaced by a search: 32 fixed rotations covering SO(3) (the 24 axis permutations + 8 more),
// each with the translation that solves the linear collinearity equations for that rotation, 12 damped iterations per
// lane, the lowest reprojection error wins (warp argmin) and is polished to convergence.  On every fixture and on
// random stress cases this is the minimiser OpenCV reaches; on exactly planar sets, where OpenCV's EPnP start fails, it
// still finds the low-cost pose.
__global__ void __launch_bounds__(kPnpWarps * 32) pnp_kernel(PnpParams p) {
  __shared__ double sP3[kPnpWarps][kPnpMaxPts * 3];
  __shared__ double sP2[kPnpWarps][kPnpMaxPts * 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kPnpWarps + warp;
  if (b >= p.B) return;
  const int n = p.N;
  const float* K = p.K + (p.K_batched ? (size_t)b * 9 : 0);
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  double* P3 = sP3[warp];
  double* P2 = sP2[warp];
  for (int i = lane; i < n * 3; i += 32) P3[i] = (double)p.pts3d[(size_t)b * n * 3 + i];
  for (int i = lane; i < n * 2; i += 32) P2[i] = (double)p.pts2d[(size_t)b * n * 2 + i];
  __syncwarp();

  // ---- start of this lane: rotation from the table, translation from the linear collinearity equations
  double R[9], t[3];
  pnp_start_rotation(lane, R);
  {
    double sx = 0, sy = 0, sxy2 = 0, b0 = 0, b1 = 0, b2 = 0, zmin = INFINITY;
    for (int i = 0; i < n; ++i) {
      const double X = P3[i * 3], Y = P3[i * 3 + 1], Z = P3[i * 3 + 2];
      const double rx = R[0] * X + R[1] * Y + R[2] * Z, ry = R[3] * X + R[4] * Y + R[5] * Z, rz = R[6] * X + R[7] * Y + R[8] * Z;
      const double x = (P2[i * 2] - cx) / fx, y = (P2[i * 2 + 1] - cy) / fy;
      const double c1 = x * rz - rx, c2 = y * rz - ry;  // [1 0 -x] t = c1, [0 1 -y] t = c2
      sx += x;
      sy += y;
      sxy2 += x * x + y * y;
      b0 += c1;
      b1 += c2;
      b2 += -x * c1 - y * c2;
      zmin = fmin(zmin, rz);
    }
    // normal equations [[n 0 -sx] [0 n -sy] [-sx -sy sxy2]] t = b, eliminated by hand
    const double dn = (double)n;
    const double den = sxy2 - (sx * sx + sy * sy) / dn;
    double tz = (fabs(den) > 1e-12) ? (b2 + (sx * b0 + sy * b1) / dn) / den : 2.0;
    if (!(zmin + tz > 0.1)) tz = 0.6 - zmin;  // keep every point in front of the camera
    t[0] = (b0 + sx * tz) / dn;
    t[1] = (b1 + sy * tz) / dn;
    t[2] = tz;
  }
  double cost = pnp_lm(R, t, P3, P2, n, fx, fy, cx, cy, 12, 1e-3);
  if (!(cost == cost)) cost = INFINITY;
  // ---- warp argmin (ties -> lowest lane), broadcast the winner, polish
  double best = cost;
  int who = lane;
  for (int off = 16; off > 0; off >>= 1) {
    const double oc = __shfl_xor_sync(0xffffffffu, best, off);
    const int ow = __shfl_xor_sync(0xffffffffu, who, off);
    if (oc < best || (oc == best && ow < who)) {
      best = oc;
      who = ow;
    }
  }
  for (int k = 0; k < 9; ++k) R[k] = __shfl_sync(0xffffffffu, R[k], who);
  for (int k = 0; k < 3; ++k) t[k] = __shfl_sync(0xffffffffu, t[k], who);
  pnp_lm(R, t, P3, P2, n, fx, fy, cx, cy, p.max_iters, 1e-9);
  if (lane != 0) return;
  double r[3];
  log_so3(R, r);
  float* o = p.pose6 + (size_t)b * 6;
  for (int k = 0; k < 3; ++k) {
    o[k] = (float)r[k];
    o[3 + k] = (float)t[k];
  }
  if (p.rot6d != nullptr) {
    // geometries.py:164-232 on the fp32 angle-axis (eps = 1e-6 in the normalisation, first-order branch below it),
    // then rotmat_to_rot6d (:117-132): the first two ROWS of the matrix
    const float ax = o[0], ay = o[1], az = o[2];
    const float th2 = ax * ax + ay * ay + az * az;
    float m[6];
    if (th2 > 1e-6f) {
      const float th = sqrtf(th2);
      const float wx = ax / (th + 1e-6f), wy = ay / (th + 1e-6f), wz = az / (th + 1e-6f);
      const float c = cosf(th), s = sinf(th), k1 = 1.0f - c;
      m[0] = c + wx * wx * k1;
      m[1] = wx * wy * k1 - wz * s;
      m[2] = wy * s + wx * wz * k1;
      m[3] = wz * s + wx * wy * k1;
      m[4] = c + wy * wy * k1;
      m[5] = -wx * s + wy * wz * k1;
    } else {
      m[0] = 1.0f; m[1] = -az; m[2] = ay;
      m[3] = az; m[4] = 1.0f; m[5] = -ax;
    }
    float* q = p.rot6d + (size_t)b * 6;
    for (int k = 0; k < 6; ++k) q[k] = m[k];
  }
}

int launch_pnp(const PnpParams& p, cudaStream_t s) {
  pnp_kernel<<<(p.B + kPnpWarps - 1) / kPnpWarps, kPnpWarps * 32, 0, s>>>(p);
  count_launch();
  HRP_CUDA_CHECK(cudaGetLastError());
  return HRP_OK;
}

}  // namespace hrp
