"""Batched PnP on the GPU: drop-in for the forward of `BPnP_m3d`
(lib/utils/BPnP.py:114-152), which the reference's evaluation loop calls per batch to obtain ground-truth rotations on
real datasets (scripts/test.py:120-125) and which runs `cv2.solvePnP` twice per sample on the CPU.

    out = BPnP_m3d.apply(gt_keypoints2d_original, world_3d_pts, K_original[0])      # (B,6): angle-axis, translation
    gt_rot = pnp_rot6d(gt_keypoints2d_original, world_3d_pts, K_original[0])        # test.py:122-124 in one call

Forward only: the implicit-function backward of BPnP (BPnP.py:154-210, training) is out of scope and requesting a
gradient raises.  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


def _run(pts2d, pts3d, K, want_rot6d: bool):
    for t in (pts2d, pts3d, K):
        if not t.is_cuda:
            raise _lib.HrpError("horopose_b200 has no CPU path: tensors must live on a CUDA device")
        if t.requires_grad:
            raise NotImplementedError("BPnP backward (BPnP.py:154-210) is outside the inference path")
    if pts2d.dim() != 3 or pts2d.shape[-1] != 2 or pts3d.shape != pts2d.shape[:2] + (3,):
        raise ValueError(f"pts2d (B,N,2) / pts3d (B,N,3) expected, got {tuple(pts2d.shape)} / {tuple(pts3d.shape)}")
    B, N = pts2d.shape[:2]
    p2 = pts2d.detach().to(torch.float32).contiguous()
    p3 = pts3d.detach().to(torch.float32).contiguous()
    Kc = K.detach().to(torch.float32).contiguous()
    if Kc.shape not in ((3, 3), (B, 3, 3)):
        raise ValueError(f"K must be (3,3) or (B,3,3), got {tuple(Kc.shape)}")
    pose = torch.empty(B, 6, dtype=torch.float32, device=p2.device)
    rot6 = torch.empty(B, 6, dtype=torch.float32, device=p2.device) if want_rot6d else None
    with torch.cuda.device(p2.device):
        check(_lib.lib().hrp_pnp(C.c_void_p(p2.data_ptr()), C.c_void_p(p3.data_ptr()), C.c_void_p(Kc.data_ptr()),
                                 C.c_int32(1 if Kc.dim() == 3 else 0), C.c_int32(B), C.c_int32(N),
                                 C.c_void_p(pose.data_ptr()), C.c_void_p(rot6.data_ptr()) if rot6 is not None else None,
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return pose, rot6


class BPnP_m3d:
    """`BPnP_m3d.apply(pts2d, pts3d, K, ini_pose=None)` -> P_6d (B,6).  `ini_pose` is accepted and ignored: the
    refinement converges to the same minimiser from the built-in linear initialisation."""

    @staticmethod
    def apply(pts2d, pts3d, K, ini_pose=None):
        return _run(pts2d, pts3d, K, False)[0]


def pnp_rot6d(pts2d, pts3d, K):
    """scripts/test.py:122-124: BPnP_m3d -> angle_axis_to_rotation_matrix -> rotmat_to_rot6d, (B,6)."""
    return _run(pts2d, pts3d, K, True)[1]
