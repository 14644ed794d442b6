"""CPU restatement of the callers either side of the hot path -- TEST INFRASTRUCTURE.

  f1  input pipeline: frame + bbox -> square crop -> bilinear resize to 256x256 -> uint8 CHW + updated intrinsics
      + the k_value scalar the depth head consumes.
  f2  evaluation metrics: per-batch ADD / 2-D error / joint / depth errors and the ADD / PCK AUC summary.
  f3  batched PnP: the reference's BPnP_m3d forward, whose arithmetic is OpenCV's cv2.solvePnP (opencv-python, a
      dependency of the reference that is not vendored in /root/reference; 4.13.0 in this image, the reference pins
      no version): EPnP start, Levenberg-Marquardt refinement of the pixel reprojection error.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(horopose_b200.preprocess / horopose_b200.metrics -> libhrp_b200.so) never does.  Every function cites the
reference lines it follows (paths relative to the reference root).  Pinned against the real reference by
tests/golden/make_golden.py (fixtures tests/golden/crop.npz, metrics_*.npz, pnp_*.npz; re-checked by
tests/test_eval_oracle.py).

Arithmetic notes that matter for parity:
  * the resize is torch's CPU `F.interpolate(mode="bilinear", align_corners=False)` on `uint8/255` in fp32, followed
    by `(x * 255).to(uint8)` (truncation) -- augmentations.py:181,200,231;
  * the intrinsics update runs in fp32 torch ops in the order of geometries.py:360-402 after an fp64 principal-point
    shift (roboutils.py:150-152);
  * the AUC scan compares fp32 errors with fp64 thresholds `i * delta` (numpy >= 2 promotion; the reference pins
    numpy 1.22 whose value-based casting compares in fp32 -- the oracle is "the reference's code on this box's
    numpy").
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------
# f1: crop + resize + intrinsics
# ------------------------------------------------------------------------------------------------------
def square_crop(frame: np.ndarray, bbox, K: np.ndarray):
    """roboutils.py:128-157 `resize_image`: paste frame[hmin:hmax, wmin:wmax] centred into a zero square whose side is
    the longer bbox edge; shift the principal point (fp64).  frame (H,W,3) uint8, bbox ints (wmin,hmin,wmax,hmax)."""
    wmin, hmin, wmax, hmax = (int(v) for v in bbox)
    bw, bh = wmax - wmin, hmax - hmin
    side = int(max(bw, bh))
    xo, yo = int((side - bw) // 2), int((side - bh) // 2)
    sq = np.zeros((side, side, 3), dtype=np.uint8)
    sq[yo:yo + bh, xo:xo + bw] = frame[hmin:hmax, wmin:wmax]
    K2 = np.array(K, dtype=np.float64, copy=True)
    K2[0, 2] -= (wmin - xo)
    K2[1, 2] -= (hmin - yo)
    return sq, K2


def intrinsics_crop_resize(K, side: int, out: int = 256) -> torch.Tensor:
    """geometries.py:360-402 `get_K_crop_resize` for the box (0, 0, side, side) the augmentation always passes
    (augmentations.py:193-203): fp32 torch arithmetic in the reference's operation order.  K (3,3) -> (3,3) fp32."""
    K = torch.as_tensor(np.asarray(K)).float()
    box = torch.tensor([side / 2 - side / 2, side / 2 - side / 2, side / 2 + side / 2, side / 2 + side / 2]).float()
    final = torch.tensor(float(out))
    cw, ch = box[2] - box[0], box[3] - box[1]
    cj, ci = (box[0] + box[2]) / 2, (box[1] + box[3]) / 2
    cx = K[0, 2] + (cw - 1) / 2 - cj
    cy = K[1, 2] + (ch - 1) / 2 - ci
    dx, dy = cx - (cw - 1) / 2, cy - (ch - 1) / 2
    sx, sy = final / cw, final / ch
    Kn = K.clone()
    Kn[0, 0] = sx * K[0, 0]
    Kn[1, 1] = sy * K[1, 1]
    Kn[0, 2] = (final - 1) / 2 + sx * dx
    Kn[1, 2] = (final - 1) / 2 + sy * dy
    return Kn


def crop_resize(frame: np.ndarray, bbox, K: np.ndarray, out: int = 256):
    """dream.py:297-310 / :350-362 (`_get_rootnet_data` / `_get_other_data`, no flip / padding):
    resize_image -> CropResizeToAspectAugmentation((out,out)) -> uint8 CHW.
    Returns (image uint8 (3,out,out), K fp32 (3,3)).  The augmentation is the identity when the square already has
    the target size (augmentations.py:174-176); K then stays fp64-shifted and is cast by `torch.FloatTensor`."""
    sq, K2 = square_crop(frame, bbox, K)
    side = sq.shape[0]
    if side == out:
        return torch.from_numpy(sq).permute(2, 0, 1).contiguous(), torch.as_tensor(K2).float()
    img = (torch.from_numpy(sq).float() / 255).unsqueeze(0).permute(0, 3, 1, 2)
    img = F.interpolate(img, size=(out, out), mode="bilinear", align_corners=False)
    u8 = (img[0].permute(1, 2, 0) * 255).to(torch.uint8)
    return u8.permute(2, 0, 1).contiguous(), intrinsics_crop_resize(K2, side, out)


def k_value(fx, fy, bboxes) -> torch.Tensor:
    """scripts/test.py:141-152: k = sqrt(fx * fy * 1000 * 1000 / max(|x2-x1|, |y2-y1|)^2), fp32."""
    fx, fy, bboxes = (torch.as_tensor(v).float() for v in (fx, fy, bboxes))
    real = torch.tensor([1000.0, 1000.0])
    area = torch.max(torch.abs(bboxes[:, 2] - bboxes[:, 0]), torch.abs(bboxes[:, 3] - bboxes[:, 1])) ** 2
    return torch.stack([torch.sqrt(fx[n] * fy[n] * real[0] * real[1] / area[n]) for n in range(fx.shape[0])]).float()


# ------------------------------------------------------------------------------------------------------
# f2: metrics
# ------------------------------------------------------------------------------------------------------
def project(K: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """transforms.py:7-15 (numpy variant): (K p)[:2] / (K p)[2] per sample."""
    out = []
    for Kb, pb in zip(K, pts):
        q = (Kb @ pb.T).T
        out.append(q[:, :2] / q[:, 2:3])
    return np.stack(out)


def metrics_batch(pred_kp3d, gt_kp3d, gt_kp2d, K_original, pred_joint, gt_joint, ref_id: int, panda: bool,
                  frame_wh=(640.0, 480.0)):
    """metrics.py:8-119 `compute_metrics_batch` given the predicted camera-frame keypoints (the reference obtains
    them from URDFRobot.get_keypoints[_root], :27-33).  All inputs fp32 numpy; results keep numpy's fp32.
    Returns a dict with the nine reference outputs (same names as the returned tuple, :119)."""
    p3, g3, g2 = (np.asarray(a, dtype=np.float32) for a in (pred_kp3d, gt_kp3d, gt_kp2d))
    p2 = project(np.asarray(K_original, dtype=np.float32), p3)
    e3 = np.linalg.norm(p3 - g3, ord=2, axis=2)
    e2 = np.linalg.norm(p2 - g2, ord=2, axis=2)
    ok = (g2[:, :, 0] <= frame_wh[0]) & (g2[:, :, 0] >= 0) & (g2[:, :, 1] <= frame_wh[1]) & (g2[:, :, 1] >= 0)
    e2v = e2 * ok
    res = {
        "error3d": e3.mean(axis=1),
        "error2d": e2v.sum(axis=1) / ok.sum(axis=1),
        "dis3d": e3.mean(axis=0),
        "dis2d": e2v.sum(axis=0) / ok.sum(axis=0),
    }
    if pred_joint is not None:
        ej = np.abs(np.asarray(gt_joint, dtype=np.float32) - np.asarray(pred_joint, dtype=np.float32))
        res["l1_jointerror"] = ej.mean(axis=0)
        res["mean_jointerror"] = (ej[:, :-1] if panda else ej).mean(axis=1)
    res["error_depth"] = np.abs(p3[:, ref_id, 2] - g3[:, ref_id, 2])
    pr = p3[:, :, 2] - p3[:, ref_id:ref_id + 1, 2]
    gr = g3[:, :, 2] - g3[:, ref_id:ref_id + 1, 2]
    res["batch_error_relative"] = np.abs(pr - gr).mean(axis=1)
    p3r, g3r = p3.copy(), g3.copy()
    p3r[:, :, 2], g3r[:, :, 2] = pr, gr
    res["error3d_relative"] = np.linalg.norm(p3r - g3r, ord=2, axis=2).mean(axis=1)
    return res


ADD_MM = (1, 5, 10, 20, 40, 60, 80, 100)
PCK_PX = (2.5, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0)


def _auc(d: np.ndarray, limit: float, delta: float) -> float:
    thr = np.arange(0.0, limit, delta)
    frac = [np.mean(d <= t) for t in thr]
    trapz = getattr(np, "trapezoid", None) or np.trapz
    return float(trapz(frac, dx=delta) / limit)


def summary_add_pck(dis3d, dis2d) -> dict:
    """metrics.py:122-162: AUC of the fraction-under-threshold curves (ADD: 0..0.1 m step 1e-5; PCK: 0..20 px step
    0.01), means, medians and the tabulated threshold fractions.  dis3d / dis2d: per-image fp32 errors."""
    d3, d2 = np.asarray(dis3d), np.asarray(dis2d)
    out = {"ADD/mean": np.mean(d3), "ADD/median": np.median(d3), "ADD/AUC": _auc(d3, 0.1, 0.00001),
           "ADD_2D/mean": np.mean(d2), "ADD_2D/median": np.median(d2), "PCK/AUC": _auc(d2, 20.0, 0.01)}
    for mm in ADD_MM:
        out[f"ADD_{mm}_mm"] = np.mean(d3 <= mm * 1e-3)
    for px in PCK_PX:
        out[f"PCK_{px}_pixel"] = np.mean(d2 <= px)
    return out


# ------------------------------------------------------------------------------------------------------
# f3: batched PnP (the arithmetic lives in OpenCV, a dependency of the reference: opencv-python, cv2.solvePnP)
# ------------------------------------------------------------------------------------------------------
def pnp_m3d(pts2d, pts3d, K) -> torch.Tensor:
    """BPnP.py:126-148 (`BPnP_m3d.forward`, ini_pose=None): per sample EPnP as the initial guess, then the iterative
    (Levenberg-Marquardt) refinement with useExtrinsicGuess; returns (B,6) fp32 = (angle-axis, translation)."""
    import cv2 as cv
    pts2d, pts3d = torch.as_tensor(pts2d), torch.as_tensor(pts3d)
    bs, n = pts2d.shape[:2]
    K_np = torch.as_tensor(K).detach().cpu().numpy().copy()
    out = torch.zeros(bs, 6)
    for i in range(bs):
        p2 = np.ascontiguousarray(pts2d[i].detach().cpu()).reshape((n, 1, 2))
        p3 = np.ascontiguousarray(pts3d[i].detach().cpu()).reshape((n, 3))
        _, r0, t0 = cv.solvePnP(objectPoints=p3, imagePoints=p2, cameraMatrix=K_np, distCoeffs=None, flags=cv.SOLVEPNP_EPNP)
        _, r, t = cv.solvePnP(objectPoints=p3, imagePoints=p2, cameraMatrix=K_np, distCoeffs=None,
                              flags=cv.SOLVEPNP_ITERATIVE, useExtrinsicGuess=True, rvec=r0, tvec=t0)
        out[i] = torch.cat((torch.tensor(r, dtype=torch.float).view(3), torch.tensor(t, dtype=torch.float).view(3)))
    return out


def angle_axis_to_rot6d(aa: torch.Tensor) -> torch.Tensor:
    """geometries.py:164-232 (`angle_axis_to_rotation_matrix`, eps 1e-6, first-order branch for tiny angles) followed by
    `rotmat_to_rot6d` (:117-132, first two rows) -- scripts/test.py:123-124."""
    aa = torch.as_tensor(aa).float()
    th2 = (aa * aa).sum(dim=1, keepdim=True)
    th = torch.sqrt(th2)
    w = aa / (th + 1e-6)
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    c, s = torch.cos(th), torch.sin(th)
    k1 = 1.0 - c
    normal = torch.cat([c + wx * wx * k1, wx * wy * k1 - wz * s, wy * s + wx * wz * k1,
                        wz * s + wx * wy * k1, c + wy * wy * k1, -wx * s + wy * wz * k1], dim=1)
    one = torch.ones_like(th)
    taylor = torch.cat([one, -aa[:, 2:3], aa[:, 1:2], aa[:, 2:3], one, -aa[:, 0:1]], dim=1)
    return torch.where(th2 > 1e-6, normal, taylor)
