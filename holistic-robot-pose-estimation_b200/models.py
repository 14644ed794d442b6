"""Drop-in model classes: `RootNetwithRegInt` / `get_rootNetwithRegInt_model` (lib/models/full_net.py:18-435) and
`RootNet` / `get_rootnet` (lib/models/depth_net.py:11-170), executed by libhrp_b200.so.

Same constructor arguments, `load_state_dict` keys, `forward` signature and return tuple as the reference.  The
modules hold no torch parameters: `load_state_dict` hands the reference-keyed tensors to the C ABI, which folds
BatchNorm, repacks the weights to bf16 and plans the network; `forward` passes raw device pointers.
Only the shipped configuration family is on the hot path (resnet50 + hrnet32, rotation_dim 6, fix_root, no
add_fc / multi_kp / reg_joint_map / direct_reg_rot / rot_iterative_matmul); anything else raises.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import _lib, arch, tables
from ._lib import check
from .robot import URDFRobot

MODEL_FULL, MODEL_DEPTHNET = 0, 1


class ModelDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dof", C.c_int32), ("nkpt", C.c_int32), ("ref_kpt", C.c_int32),
                ("n_iter", C.c_int32), ("fix_root", C.c_int32), ("image_size", C.c_float), ("depth_factor", C.c_float),
                ("chunk", C.c_int32), ("inflight", C.c_int32)]


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("pose", "rot", "trans", "root_uv", "depth", "uvd", "xyz_int", "xyz_fk", "uv_int", "uv_fk")]


def _get(args, name, default=None):
    if isinstance(args, dict):
        return args.get(name, default)
    return getattr(args, name, default)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


TUNING_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuning", "b200.txt")


def _tuning_table() -> str:
    """Committed kernel-variant table (produced on B200 by tools/make_tuning.py).  $HRP_TUNING=0 disables it (shape
    heuristics only), $HRP_TUNING=<path> selects another file."""
    sel = os.environ.get("HRP_TUNING", "")
    if sel == "0":
        return ""
    path = sel if sel else TUNING_PATH
    try:
        with open(path) as f:
            return f.read()
    except FileNotFoundError:
        if sel:
            raise
        return ""


class _EngineModule(nn.Module):
    """Common plumbing: weight hand-over and lazy finalisation on the first CUDA forward."""

    def __init__(self):
        super().__init__()
        self._sd = None
        self._handle = None
        self._device = None
        self.chunk = int(os.environ.get("HRP_CHUNK", "512"))
        self.inflight = int(os.environ.get("HRP_INFLIGHT", "1"))

    # -- nn.Module surface the reference callers use -----------------------------------------------------
    def load_state_dict(self, state_dict, strict=True):
        spec = self._spec()
        sd = OrderedDict()
        for k, v in state_dict.items():
            sd[k] = v.detach().cpu()
        missing = [k for k in spec if k not in sd and not k.endswith("num_batches_tracked")]
        unexpected = [k for k in sd if k not in spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        for k, shp in spec.items():
            if k in sd and tuple(sd[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(shp)}")
        if self._sd is not None and not strict:
            merged = OrderedDict(self._sd)
            merged.update(sd)
            sd = merged
        self._sd = sd
        self._release()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def state_dict(self, *a, **k):
        return OrderedDict(self._sd or {})

    def float(self):
        return self

    def to(self, *a, **k):
        return self

    def cuda(self, *a, **k):
        return self

    def _release(self):
        if self._handle is not None:
            _lib.lib().hrp_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure(self, device):
        if not torch.cuda.is_available():
            raise _lib.HrpError("horopose_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if self._handle is not None and self._device == device:
            return
        if self._sd is None:
            raise _lib.HrpError("load_state_dict() must be called before forward()")
        self._release()
        L = _lib.lib()
        with torch.cuda.device(device):
            desc = self._desc()
            h = C.c_void_p(0)
            check(L.hrp_model_create(C.byref(desc), C.byref(h)))
            try:
                for k, v in self._sd.items():
                    if v.dtype in (torch.int64, torch.int32) or k.endswith("num_batches_tracked"):
                        continue
                    a = np.ascontiguousarray(v.float().numpy())
                    shape = (C.c_int64 * a.ndim)(*a.shape)
                    check(L.hrp_model_set_tensor(h, k.encode(), a.ctypes.data_as(C.POINTER(C.c_float)), shape, a.ndim))
                self._pre_finalize(h)
                table = _tuning_table()
                if table:
                    check(L.hrp_model_set_tuning(h, table.encode()))
                check(L.hrp_model_finalize(h))
            except Exception:
                L.hrp_model_destroy(h)
                raise
        self._handle, self._device = h, device

    def _pre_finalize(self, h):
        pass

    def stats(self, batch: int) -> dict:
        """FLOPs, kernel count and activation-arena bytes of the plan built for `batch` images."""
        fl, kn, ab = C.c_double(0), C.c_int32(0), C.c_int64(0)
        check(_lib.lib().hrp_model_stats(self._handle, int(batch), C.byref(fl), C.byref(kn), C.byref(ab)))
        return {"flops": fl.value, "kernels": kn.value, "activation_bytes": ab.value}

    def tuning(self) -> str:
        """Current kernel-variant table of the handle (committed entries + anything tuned with HRP_AUTOTUNE=1)."""
        buf = C.create_string_buffer(1 << 20)
        check(_lib.lib().hrp_model_get_tuning(self._handle, buf, len(buf)))
        return buf.value.decode()

    def activation(self, name: str) -> torch.Tensor:
        """Debug/test hook: copy of a named intermediate activation as fp32 NCHW (or (B,2048) for feat / xf)."""
        ptr, B, H, W, Cc = C.c_void_p(0), C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(_lib.lib().hrp_model_activation(self._handle, name.encode(), C.byref(ptr), C.byref(B), C.byref(H),
                                              C.byref(W), C.byref(Cc)))
        if Cc.value < 0:
            n = B.value * (-Cc.value)
            tmp = torch.empty(n, dtype=torch.float32, device=self._device)
            _memcpy_d2d(tmp.data_ptr(), ptr.value, n * 4)
            return tmp.view(B.value, -Cc.value)
        n = B.value * H.value * W.value * Cc.value
        tmp = torch.empty(n, dtype=torch.bfloat16, device=self._device)
        _memcpy_d2d(tmp.data_ptr(), ptr.value, n * 2)
        return tmp.view(B.value, H.value, W.value, Cc.value).permute(0, 3, 1, 2).float().contiguous()


def _memcpy_d2d(dst, src, nbytes):
    check(_lib.lib().hrp_copy_device(C.c_void_p(dst), C.c_void_p(src), C.c_int64(nbytes), _stream()))
    torch.cuda.synchronize()


class RootNetwithRegInt(_EngineModule):
    def __init__(self, init_param_dict, args, **kwargs):
        super().__init__()
        robot_type = init_param_dict["robot_type"]
        if robot_type not in arch.ROBOTS:
            raise ValueError(f"Robot type {robot_type} is not supported.")
        self.robot_type = robot_type
        self.dof, self.num_joints, _ = arch.ROBOTS[robot_type]
        self.backbone_name = _get(args, "backbone_name", "resnet50")
        self.rootnet_backbone_name = _get(args, "rootnet_backbone_name", "hrnet32")
        if self.backbone_name not in ("resnet", "resnet50") or self.rootnet_backbone_name not in ("hrnet", "hrnet32"):
            raise NotImplementedError("hot path = backbone_name resnet50 + rootnet_backbone_name hrnet32 "
                                      "(configs/*/full.yaml:17-18)")
        for flag in ("reg_joint_map", "direct_reg_rot", "rot_iterative_matmul", "multi_kp", "add_fc", "use_rpmg"):
            if _get(args, flag, False):
                raise NotImplementedError(f"args.{flag}=True is outside the hot path (no shipped config enables it)")
        if int(_get(args, "rotation_dim", 6)) != 6:
            raise NotImplementedError("rotation_dim must be 6 (configs/*/full.yaml)")
        self.n_iter = int(_get(args, "n_iter", 4))
        self.image_size = float(np.asarray(_get(args, "other_image_size", 256.0)).reshape(-1)[0])
        if self.image_size != 256.0:
            raise NotImplementedError("the engine is specialised for 256x256 crops (full.yaml image sizes)")
        self.depth_dim = 64
        self.bbox_3d_shape = _get(args, "bbox_3d_shape", [1300, 1300, 1300])
        self.reference_keypoint_id = int(_get(args, "reference_keypoint_id", 3))
        self.fix_root = bool(_get(args, "fix_root", True))
        self.depth_factor = float(torch.tensor(self.bbox_3d_shape).float()[2] * 1e-3)   # integral.py:91-93
        self.robot = URDFRobot(robot_type, urdf_path=kwargs.get("urdf_path"))
        pose_params = init_param_dict.get("pose_params")
        from_mean = init_param_dict.get("init_pose_from_mean", True)
        if pose_params is not None:
            tbl = pose_params["mean" if from_mean else "zero"][robot_type]
            init_pose = [tbl[k] for k in tables.JOINT_NAMES[robot_type]]
        else:
            init_pose = (tables.INIT_POSE_MEAN if from_mean else tables.INIT_POSE_ZERO)[robot_type]
        cam = np.asarray(init_param_dict.get("cam_params", np.eye(4)), dtype=np.float64)
        self.init_pose = torch.tensor([init_pose], dtype=torch.float32)
        self.init_rot = torch.tensor(cam[:2, :3].reshape(1, 6), dtype=torch.float32)   # rotmat_to_rot6d: first 2 rows

    def _spec(self):
        return arch.full_model_spec(self.robot_type)

    def _desc(self):
        return ModelDesc(MODEL_FULL, self.dof, self.num_joints, self.reference_keypoint_id, self.n_iter,
                         int(self.fix_root), self.image_size, self.depth_factor, self.chunk, self.inflight)

    def _pre_finalize(self, h):
        check(_lib.lib().hrp_model_set_robot(h, self.robot.handle()))

    def forward(self, x_reg_input, x_root_input, k_value, K, init_pose=None, init_rot=None, test_fps=False):
        if not x_reg_input.is_cuda:
            raise _lib.HrpError("horopose_b200 has no CPU path: inputs must live on a CUDA device")
        dev = x_reg_input.device
        self._ensure(dev)
        t0 = time.time()
        B = x_reg_input.shape[0]
        # uint8 images (0..255) are accepted as-is: the engine fuses the caller-side `.float() / 255.`
        # (scripts/test.py:83-86) into its input-packing kernel
        u8 = (x_reg_input.dtype == torch.uint8 and x_root_input.dtype == torch.uint8)
        dt = torch.uint8 if u8 else torch.float32
        x_reg = x_reg_input.detach().to(dt).contiguous()
        x_root = x_root_input.detach().to(device=dev, dtype=dt).contiguous()
        k_value = k_value.detach().to(device=dev, dtype=torch.float32).contiguous().view(-1)
        K = K.detach().to(device=dev, dtype=torch.float32).contiguous()
        assert x_reg.shape == (B, 3, 256, 256) and x_root.shape == (B, 3, 256, 256), (x_reg.shape, x_root.shape)
        assert k_value.shape == (B,) and K.shape == (B, 3, 3)
        ip = ir = None
        if init_pose is not None or init_rot is not None:
            ip = (self.init_pose.expand(B, -1) if init_pose is None else init_pose).to(dev).float().contiguous()
            ir = (self.init_rot.expand(B, -1) if init_rot is None else init_rot).to(dev).float().contiguous()
        nk = self.num_joints
        shapes = dict(pose=(B, self.dof), rot=(B, 6), trans=(B, 3), root_uv=(B, 2), depth=(B, 1), uvd=(B, nk, 3),
                      xyz_int=(B, nk, 3), xyz_fk=(B, nk, 3))
        outs = {n: torch.empty(*s, dtype=torch.float32, device=dev) for n, s in shapes.items()}
        o = Outputs()
        for n, t in outs.items():
            setattr(o, n, t.data_ptr())
        with torch.cuda.device(dev):
            fwd = _lib.lib().hrp_model_forward_u8 if u8 else _lib.lib().hrp_model_forward
            check(fwd(self._handle, C.c_void_p(x_reg.data_ptr()), C.c_void_p(x_root.data_ptr()),
                      C.c_void_p(k_value.data_ptr()), C.c_void_p(K.data_ptr()),
                      C.c_void_p(ip.data_ptr() if ip is not None else 0),
                      C.c_void_p(ir.data_ptr() if ir is not None else 0), B, C.byref(o), _stream()))
        res = (outs["pose"], outs["rot"], outs["trans"], outs["root_uv"], outs["depth"], outs["uvd"], outs["xyz_int"],
               outs["xyz_fk"])
        if test_fps:  # full_net.py:253-289,385-392: wall-clock with a stream sync; the fused engine has one phase
            torch.cuda.current_stream().synchronize()
            t = time.time() - t0
            return (*res, (0.0, t, t))
        return res


def get_rootNetwithRegInt_model(init_params_dict, args, **kwargs):
    """full_net.py:401-435 (pretrained-depthnet remap included)."""
    if _get(args, "backbone_name") not in ["resnet", "resnet50", "resnet34", "resnet101", "hrnet", "hrnet32"]:
        raise NotImplementedError
    if _get(args, "rootnet_backbone_name") not in ["resnet", "resnet50", "resnet34", "hrnet", "hrnet32"]:
        raise NotImplementedError
    model = RootNetwithRegInt(init_params_dict, args, **kwargs)
    pretrained = _get(args, "pretrained_rootnet")
    if pretrained is not None:
        ckpt = torch.load(pretrained, map_location="cpu")
        weights = {(k.replace("backbone", "rootnet_backbone") if k.startswith("backbone") else k): v
                   for k, v in ckpt["model_state_dict"].items()}
        model.load_state_dict(weights, strict=False)
    return model


class RootNet(_EngineModule):
    def __init__(self, backbone, pred_xy=False, use_offset=False, add_fc=False, input_shape=(256, 256), **kwargs):
        super().__init__()
        if backbone not in ("hrnet", "hrnet32"):
            raise NotImplementedError("hot path = RootNet('hrnet32') (configs/*/depthnet.yaml)")
        if pred_xy or use_offset or add_fc:
            raise NotImplementedError("pred_xy / use_offset / add_fc are off in every shipped config")
        self.backbone_name = backbone

    def _spec(self):
        return arch.depthnet_spec()

    def _desc(self):
        return ModelDesc(MODEL_DEPTHNET, 0, 0, 0, 0, 0, 256.0, 1.3, self.chunk, self.inflight)

    def init_weights(self):
        pass

    def forward(self, x, k_value):
        if not x.is_cuda:
            raise _lib.HrpError("horopose_b200 has no CPU path: inputs must live on a CUDA device")
        dev = x.device
        self._ensure(dev)
        B = x.shape[0]
        x = x.detach().to(torch.float32).contiguous()
        k_value = k_value.detach().to(device=dev, dtype=torch.float32).contiguous().view(-1)
        out = torch.empty(B, 1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().hrp_model_depthnet_forward(self._handle, C.c_void_p(x.data_ptr()),
                                                        C.c_void_p(k_value.data_ptr()), B, C.c_void_p(out.data_ptr()),
                                                        _stream()))
        return out


def get_rootnet(backbone, pred_xy=False, use_offset=False, add_fc=False, input_shape=(256, 256), **kwargs):
    return RootNet(backbone, pred_xy, use_offset, add_fc, input_shape=(256, 256), **kwargs)
