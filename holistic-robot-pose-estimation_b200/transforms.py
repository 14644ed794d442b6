"""Drop-ins for the geometry helpers of the reference's `lib/utils/transforms.py` that sit on the inference path
(`uvd_to_xyz` :33-73, `uvz2xyz_singlepoint` :133-143, `get_intrinsic_matrix_batch` :145-162 / lib/utils/integral.py:56-73,
`point_projection_from_3d[_tensor]` :11-21), each one CUDA kernel behind the C ABI (include/hrp.h).  Inside the full model
the fused head evaluates the same expressions on chip; these are the operator-level entry points.  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check
from .robot import point_projection_from_3d, point_projection_from_3d_tensor  # noqa: F401  (same module in the reference)


def _f32(t: torch.Tensor, dev=None) -> torch.Tensor:
    t = torch.as_tensor(t)
    if dev is not None:
        t = t.to(dev)
    if not t.is_cuda:
        raise _lib.HrpError("horopose_b200 has no CPU path: tensors must live on a CUDA device")
    return t.detach().to(torch.float32).contiguous()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def get_intrinsic_matrix_batch(f, c, bsz, inv=False):
    """f = (fx, fy), c = (cx, cy) as (bsz,) tensors -> (bsz,3,3) fp32 on the GPU; `inv=True` divides in float64."""
    fx, fy, cx, cy = (torch.as_tensor(v).reshape(-1) for v in (f[0], f[1], c[0], c[1]))
    dev = fx.device if fx.is_cuda else torch.device("cuda", torch.cuda.current_device())
    K = torch.zeros(bsz, 3, 3, dtype=torch.float32, device=dev)
    K[:, 0, 0], K[:, 1, 1] = fx.to(dev).float(), fy.to(dev).float()
    K[:, 0, 2], K[:, 1, 2] = cx.to(dev).float(), cy.to(dev).float()
    K[:, 2, 2] = 1.0
    if not inv:
        return K
    out = torch.empty_like(K)
    with torch.cuda.device(dev):
        check(_lib.lib().hrp_inv_intrinsics(C.c_void_p(K.data_ptr()), C.c_void_p(out.data_ptr()), bsz, _stream()))
    return out


def uvd_to_xyz(uvd_jts, image_size, intrinsic_matrix_inverse, root_trans, depth_factor, return_relative=False):
    assert uvd_jts.dim() == 3 and uvd_jts.shape[2] == 3, uvd_jts.shape
    uvd = _f32(uvd_jts)
    dev = uvd.device
    kinv, rt = _f32(intrinsic_matrix_inverse, dev), _f32(root_trans, dev)
    B, N = uvd.shape[0], uvd.shape[1]
    assert kinv.shape == (B, 3, 3) and rt.shape == (B, 3), (kinv.shape, rt.shape)
    out = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hrp_uvd_to_xyz(C.c_void_p(uvd.data_ptr()), C.c_void_p(kinv.data_ptr()), C.c_void_p(rt.data_ptr()),
                                        C.c_float(float(image_size)), C.c_float(float(depth_factor)),
                                        int(bool(return_relative)), B, N, C.c_void_p(out.data_ptr()), _stream()))
    return out


def uvz2xyz_singlepoint(uv, z, K):
    batch_size = uv.shape[0]
    assert uv.shape == (batch_size, 2) and z.shape == (batch_size, 1) and K.shape == (batch_size, 3, 3), \
        (uv.shape, z.shape, K.shape)
    uv = _f32(uv)
    dev = uv.device
    z, K = _f32(z, dev).view(-1), _f32(K, dev)
    out = torch.empty(batch_size, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hrp_uvz2xyz_singlepoint(C.c_void_p(uv.data_ptr()), C.c_void_p(z.data_ptr()),
                                                 C.c_void_p(K.data_ptr()), batch_size, C.c_void_p(out.data_ptr()),
                                                 _stream()))
    return out
