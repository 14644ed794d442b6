"""GPU input pipeline: camera frame + bounding box -> network crop, intrinsics, k_value.

Drop-in for the per-sample chain the reference's dataset runs on the CPU for every image --
`resize_image` (lib/dataset/roboutils.py:128-157) -> `CropResizeToAspectAugmentation((256, 256))`
(lib/dataset/augmentations.py:165-233) -> uint8 CHW (lib/dataset/dream.py:297-310 / :350-362) -- plus
`get_K_crop_resize` (lib/utils/geometries.py:360-402) and the `k_values` expression of scripts/test.py:141-152,
batched into one kernel of libhrp_b200.so.  The crop is bit-exact with the reference's bytes, so
`model(crops, crops, k_values, K)` sees exactly what the reference model would.  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


class CropArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("frame_h", C.c_int32), ("frame_w", C.c_int32), ("out_size", C.c_int32),
                ("frames", C.c_void_p), ("bbox", C.c_void_p), ("K_in", C.c_void_p), ("out_u8", C.c_void_p),
                ("K_out", C.c_void_p), ("k_bbox", C.c_void_p), ("k_value", C.c_void_p), ("k_use_crop_K", C.c_int32)]


def _cuda(t: torch.Tensor, dtype) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.HrpError("horopose_b200 has no CPU path: tensors must live on a CUDA device")
    return t.detach().to(dtype).contiguous()


def crop_resize_batch(frames: torch.Tensor, bbox: torch.Tensor, K: torch.Tensor, resize_hw=(256, 256),
                      k_bbox: torch.Tensor | None = None, k_from_crop_K: bool = False, check_bbox: bool = True):
    """frames (B,H,W,3) uint8 CUDA (HWC, as decoded); bbox (B,4) integer (wmin,hmin,wmax,hmax), inside the frame;
    K (B,3,3) camera matrix of the full frame (kept in float64 for the principal-point shift, as the reference does).
    Returns (images uint8 (B,3,h,w), K float32 (B,3,3)[, k_value float32 (B)]) -- the dataset's "images" / "K"
    entries (dream.py:324-332) and, when `k_bbox` (B,4) is given, scripts/test.py's `k_values` (fx, fy taken from
    `K` like `args.use_origin_bbox`, or from the crop's K when `k_from_crop_K`).
    `check_bbox=False` skips the host-side box validation (a device->host sync when `bbox` lives on the GPU); the
    kernel then reads whatever the box addresses, so only pass boxes that are known to lie inside the frame."""
    if tuple(resize_hw) != (int(resize_hw[0]), int(resize_hw[0])) or int(resize_hw[0]) % 4:
        raise NotImplementedError("only square crops with a side that is a multiple of 4 are on the path (256 x 256)")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError(f"frames must be uint8 (B,H,W,3), got {frames.dtype} {tuple(frames.shape)}")
    B, H, W, _ = frames.shape
    frames = _cuda(frames, torch.uint8)
    if tuple(bbox.shape) != (B, 4):
        raise ValueError(f"bbox must be (B,4) = (wmin, hmin, wmax, hmax), got {tuple(bbox.shape)}")
    if check_bbox:
        bb_host = bbox.detach().cpu().to(torch.int64)
        # the reference pastes image[hmin:hmax, wmin:wmax] into a (hmax-hmin, wmax-wmin) window: numpy raises on a box
        # that leaves the frame or is empty (roboutils.py:137)
        if bool(((bb_host[:, 0] < 0) | (bb_host[:, 1] < 0) | (bb_host[:, 2] > W) | (bb_host[:, 3] > H)
                 | (bb_host[:, 2] <= bb_host[:, 0]) | (bb_host[:, 3] <= bb_host[:, 1])).any()):
            raise ValueError("bbox must be (wmin, hmin, wmax, hmax) inside the frame with positive extent")
    bb = bbox.detach().to(device=frames.device, dtype=torch.int32).contiguous()
    Kd = _cuda(K.to(frames.device), torch.float64)
    out = int(resize_hw[0])
    images = torch.empty(B, 3, out, out, dtype=torch.uint8, device=frames.device)
    K_out = torch.empty(B, 3, 3, dtype=torch.float32, device=frames.device)
    kb = kv = None
    if k_bbox is not None:
        kb = _cuda(k_bbox.to(frames.device), torch.float32)
        kv = torch.empty(B, dtype=torch.float32, device=frames.device)
    a = CropArgs(B, H, W, out, frames.data_ptr(), bb.data_ptr(), Kd.data_ptr(), images.data_ptr(), K_out.data_ptr(),
                 kb.data_ptr() if kb is not None else None, kv.data_ptr() if kv is not None else None,
                 1 if k_from_crop_K else 0)
    with torch.cuda.device(frames.device):
        check(_lib.lib().hrp_crop_resize(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return (images, K_out) if kv is None else (images, K_out, kv)
