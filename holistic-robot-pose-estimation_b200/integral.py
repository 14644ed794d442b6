"""Drop-in for `HeatmapIntegralPose` (lib/utils/integral.py:76-186) on top of the fused head kernel.

forward(out, K=..., root_trans=...) takes the reference's heatmap-logit tensor (B, nkpt*64, 64, 64) (channel =
k*64 + d) and returns (pred_uvd_jts (B,nkpt,3), pred_xyz_jts (B,nkpt,3)).  The logits are bridged to the kernel's
pixel-major layout (one extra pass) as fp32 (default) or rounded to bf16 (`logits_dtype="bf16"`: the type the network's
final conv would store; inside the full model there is no hand-off, the fold consumes the fp32 accumulators).
When the logits require a gradient (training), the soft-argmax backward runs on the GPU as well (`_IntegralUVD`).
"""
from __future__ import annotations

import os

import ctypes as C

import torch

from . import _lib, ops
from ._lib import check


class HeadArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("nkpt", C.c_int32), ("ref_kpt", C.c_int32), ("fix_root", C.c_int32),
                ("image_size", C.c_float), ("depth_factor", C.c_float),
                ("heatmap", C.c_void_p), ("K", C.c_void_p), ("root_depth", C.c_void_p), ("robot", C.c_void_p),
                ("pose", C.c_void_p), ("rot", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("uvd", C.c_void_p), ("xyz_int", C.c_void_p), ("root_uv", C.c_void_p), ("trans", C.c_void_p),
                ("xyz_fk", C.c_void_p), ("uv_int", C.c_void_p), ("uv_fk", C.c_void_p), ("heatmap_fp32", C.c_int32)]


_workspaces = {}


def _workspace(device, B, nkpt):
    need = C.c_int64(0)
    check(_lib.lib().hrp_head_workspace_bytes(B, nkpt, C.byref(need)))
    # one workspace per (device, stream, shape): the kernel keeps per-image counters and partial sums there, so two
    # streams running the head concurrently must not share it
    key = (device, torch.cuda.current_stream(device).cuda_stream, B, nkpt)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(need.value, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def run_head(heatmap_nhwc, K, root_depth, *, nkpt, ref_kpt, fix_root=True, image_size=256.0, depth_factor=1.3,
             robot=None, pose=None, rot=None, want_uv=False, workspace=None):
    """heatmap_nhwc: (B,64,64,nkpt*64) CUDA, bf16 (the layout the final convolution writes) or fp32 (a caller's logits,
    no rounding).  Returns dict of fp32 CUDA tensors."""
    assert heatmap_nhwc.is_cuda and heatmap_nhwc.dtype in (torch.bfloat16, torch.float32) and heatmap_nhwc.is_contiguous()
    B = heatmap_nhwc.shape[0]
    dev = heatmap_nhwc.device
    K = K.detach().to(device=dev, dtype=torch.float32).contiguous()
    root_depth = root_depth.detach().to(device=dev, dtype=torch.float32).contiguous().view(-1)
    out = {n: torch.empty(B, *shape, dtype=torch.float32, device=dev) for n, shape in
           (("uvd", (nkpt, 3)), ("xyz_int", (nkpt, 3)), ("root_uv", (2,)), ("trans", (3,)))}
    a = HeadArgs()
    a.B, a.nkpt, a.ref_kpt, a.fix_root = B, nkpt, ref_kpt, int(fix_root)
    a.image_size, a.depth_factor = float(image_size), float(depth_factor)
    a.heatmap, a.K, a.root_depth = heatmap_nhwc.data_ptr(), K.data_ptr(), root_depth.data_ptr()
    a.heatmap_fp32 = int(heatmap_nhwc.dtype == torch.float32)
    keep = [K, root_depth]
    if robot is not None and pose is not None:
        pose = pose.detach().to(device=dev, dtype=torch.float32).contiguous()
        rot = rot.detach().to(device=dev, dtype=torch.float32).contiguous()
        keep += [pose, rot]
        a.robot, a.pose, a.rot = robot.handle(dev), pose.data_ptr(), rot.data_ptr()
        out["xyz_fk"] = torch.empty(B, nkpt, 3, dtype=torch.float32, device=dev)
        a.xyz_fk = out["xyz_fk"].data_ptr()
        if want_uv:
            out["uv_fk"] = torch.empty(B, nkpt, 2, dtype=torch.float32, device=dev)
            a.uv_fk = out["uv_fk"].data_ptr()
    if want_uv:
        out["uv_int"] = torch.empty(B, nkpt, 2, dtype=torch.float32, device=dev)
        a.uv_int = out["uv_int"].data_ptr()
    ws = workspace if workspace is not None else _workspace(dev, B, nkpt)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.uvd, a.xyz_int = out["uvd"].data_ptr(), out["xyz_int"].data_ptr()
    a.root_uv, a.trans = out["root_uv"].data_ptr(), out["trans"].data_ptr()
    with torch.cuda.device(dev):
        check(_lib.lib().hrp_head(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


class _IntegralUVD(torch.autograd.Function):
    """uvd = soft-argmax(heatmap logits) with a gradient to the logits (first piece): the
    forward is the fused head kernel, the backward `hrp_head_backward_heatmap`, which re-uses the softmax statistics
    the forward left in its (private, saved) workspace -- one heatmap read and one gradient write."""

    @staticmethod
    def forward(ctx, out, K, root_depth, nkpt, rootid, fixroot, image_size, depth_factor, fp32_logits=True):
        hm = _pixel_major(out, fp32_logits)
        need = C.c_int64(0)
        check(_lib.lib().hrp_head_workspace_bytes(out.shape[0], nkpt, C.byref(need)))
        ws = torch.zeros(need.value, dtype=torch.uint8, device=out.device)
        r = run_head(hm, K, root_depth, nkpt=nkpt, ref_kpt=rootid, fix_root=fixroot, image_size=image_size,
                     depth_factor=depth_factor, workspace=ws)
        ctx.save_for_backward(hm, r["uvd"], ws)
        ctx.cfg = (nkpt, rootid, fixroot, tuple(out.shape))
        return r["uvd"]

    @staticmethod
    def backward(ctx, grad_uvd):
        hm, uvd, ws = ctx.saved_tensors
        nkpt, rootid, fixroot, shape = ctx.cfg
        B = hm.shape[0]
        g = grad_uvd.detach().to(torch.float32).contiguous()
        grad = torch.empty(B, 64, 64, nkpt * 64, dtype=torch.float32, device=hm.device)
        with torch.cuda.device(hm.device):
            check(_lib.lib().hrp_head_backward_heatmap(
                C.c_void_p(hm.data_ptr()), C.c_void_p(uvd.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(ws.data_ptr()),
                C.c_int64(ws.numel()), C.c_int32(B), C.c_int32(nkpt), C.c_int32(rootid), C.c_int32(int(fixroot)),
                C.c_int32(3 if hm.dtype == torch.float32 else 1), C.c_void_p(grad.data_ptr()),
                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return grad.permute(0, 3, 1, 2).reshape(shape), None, None, None, None, None, None, None, None


def _pixel_major(out, fp32_logits):
    """The reference's heatmap (B, nkpt*64, 64, 64) -> pixel-major (B, 64, 64, nkpt*64): fp32 as given (one transposing
    copy), or rounded to bf16 -- the layout and type the network's final convolution would write."""
    x = out.detach().float()
    if fp32_logits:
        return x.reshape(x.shape[0], -1, 64, 64).permute(0, 2, 3, 1).contiguous()
    return ops.nchw_to_nhwc_bf16(x.contiguous(), cpad=out.shape[1])


def _inv_intrinsics(K):
    """integral.py:56-73: the divisions run in float64, the result is stored as float32."""
    Kd = K.double()
    inv = torch.zeros_like(Kd)
    inv[:, 0, 0], inv[:, 1, 1] = 1.0 / Kd[:, 0, 0], 1.0 / Kd[:, 1, 1]
    inv[:, 0, 2], inv[:, 1, 2] = -Kd[:, 0, 2] / Kd[:, 0, 0], -Kd[:, 1, 2] / Kd[:, 1, 1]
    inv[:, 2, 2] = 1.0
    return inv.float()


def _uvd_to_xyz(uvd, image_size, inv_k, root_trans, depth_factor):
    """transforms.py:33-73 on (B,nkpt,3) tensors (a few hundred bytes per image: left to torch so that autograd carries
    the gradient from xyz back to uvd, K and root_trans)."""
    u = (uvd[:, :, 0:1] + 0.5) * image_size
    v = (uvd[:, :, 1:2] + 0.5) * image_size
    z = uvd[:, :, 2:3] * depth_factor + root_trans[:, None, 2:3]
    pix = torch.cat((u, v, torch.ones_like(u)), dim=2)
    return torch.matmul(inv_k[:, None], pix[..., None]).squeeze(-1) * z


class HeatmapIntegralPose(torch.nn.Module):
    def __init__(self, backbone, **kwargs):
        super().__init__()
        self.backbone_name = backbone
        self.norm_type = kwargs["norm_type"]
        if self.norm_type != "softmax":
            raise NotImplementedError("only norm_type='softmax' is on the hot path (integral.py:45-54)")
        self.num_joints = kwargs["num_joints"]
        self.depth_dim, self.height_dim, self.width_dim = kwargs["depth_dim"], kwargs["height_dim"], kwargs["width_dim"]
        if (self.depth_dim, self.height_dim, self.width_dim) != (64, 64, 64):
            raise NotImplementedError("the fused head is specialised for 64x64x64 heatmaps (full_net.py:64-66)")
        self.rootid = kwargs.get("rootid", 0)
        self.fixroot = kwargs.get("fixroot", False)
        bbox_3d_shape = kwargs.get("bbox_3d_shape", (2300, 2300, 2300))
        self.bbox_3d_shape = torch.tensor(bbox_3d_shape).float()
        self.depth_factor = self.bbox_3d_shape[2] * 1e-3   # integral.py:91-93 (an fp32 tensor product)
        self.image_size = kwargs["image_size"]
        # fp32 logits are consumed as fp32 (exactly the reference's arithmetic up to the order of the sums); "bf16" rounds
        # them first to what the network's final convolution stores and halves the bytes the kernel reads
        self.logits_dtype = kwargs.get("logits_dtype", os.environ.get("HRP_INTEGRAL_LOGITS", "fp32"))
        if self.logits_dtype not in ("fp32", "bf16"):
            raise ValueError("logits_dtype must be 'fp32' or 'bf16'")

    def forward(self, out, flip_test=False, **kwargs):
        K, root_trans = kwargs["K"], kwargs["root_trans"]
        if not out.is_cuda:
            raise _lib.HrpError("horopose_b200 has no CPU path: tensors must live on a CUDA device")
        if torch.is_grad_enabled() and out.requires_grad:
            # training (lib/core/function.py:253-311): gradient to the logits through the backward kernel; xyz from uvd in
            # torch so that K / root_trans gradients flow as in the reference
            K = K.to(out.device).float()
            root_trans = root_trans.to(out.device).float()
            uvd = _IntegralUVD.apply(out, K.detach(), root_trans[:, 2].detach(), self.num_joints, self.rootid,
                                     bool(self.fixroot), float(self.image_size), float(self.depth_factor),
                                     self.logits_dtype == "fp32")
            xyz = _uvd_to_xyz(uvd, float(self.image_size), _inv_intrinsics(K), root_trans, float(self.depth_factor))
            return uvd, xyz
        hm = _pixel_major(out, self.logits_dtype == "fp32")
        r = run_head(hm, K, root_trans[:, 2].to(out.device), nkpt=self.num_joints, ref_kpt=self.rootid,
                     fix_root=self.fixroot, image_size=float(self.image_size), depth_factor=float(self.depth_factor))
        return r["uvd"], r["xyz_int"]
