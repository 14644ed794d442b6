"""Drop-in for the reference's `URDFRobot` (lib/utils/urdf_robot.py:22-199) and projection helpers
(lib/utils/transforms.py:7-21), backed by the CUDA forward-kinematics / projection kernels of libhrp_b200.so.

Same method names, argument meaning and return shapes as the reference; tensors are fp32 CUDA tensors.
Rendering members (`robot_for_render`, `urdf_path_visual`, urdf_robot.py:201-387) are out of scope.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib, tables, urdf
from ._lib import check

_DOF = {"panda": 8, "kuka": 7, "baxter": 15}


class LinkRow(C.Structure):
    _fields_ = [("parent", C.c_int32), ("jtype", C.c_int32), ("qcol", C.c_int32), ("pad", C.c_int32),
                ("qmul", C.c_double), ("qoff", C.c_double), ("origin", C.c_double * 16), ("axis", C.c_double * 3)]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.HrpError("horopose_b200 has no CPU path: tensors must live on a CUDA device")
    return t.detach().to(torch.float32).contiguous()


# where the reference looks for the robot descriptions, relative to the working directory (lib/config.py:21,33-36;
# its Baxter path is an absolute path on the authors' machine, the deps/ layout is used here instead)
_REFERENCE_URDF = {
    "panda": "data/deps/panda-description/panda.urdf",
    "kuka": "data/deps/kuka-description/iiwa_description/urdf/iiwa7.urdf",
    "baxter": "data/deps/baxter-description/baxter_description/urdf/baxter.urdf",
}


def resolve_urdf_path(robot_type: str, urdf_path: str | None = None) -> str:
    """Explicit argument > $HRP_URDF_<ROBOT> > the reference's `data/deps/...` location under the working directory.
    `urdf_path="synthetic"` (or `synth.use_synthetic_urdfs()`, which sets the environment variables) selects the
    mesh-free fixtures shipped for tests and benchmarks.  There is NO silent fallback: a real checkpoint evaluated on
    approximate kinematics would give silently wrong keypoints."""
    if urdf_path == "synthetic":
        from . import synth
        return str(synth.URDF_PATHS[robot_type])
    if urdf_path is not None:
        return str(urdf_path)
    env = os.environ.get(f"HRP_URDF_{robot_type.upper()}")
    if env:
        return env
    if os.path.exists(_REFERENCE_URDF[robot_type]):
        return os.path.abspath(_REFERENCE_URDF[robot_type])
    raise _lib.HrpError(
        f"no URDF for robot {robot_type!r}: pass urdf_path=..., set HRP_URDF_{robot_type.upper()}, or place the "
        f"description at ./{_REFERENCE_URDF[robot_type]} as the reference does (lib/config.py:33-36); "
        "urdf_path='synthetic' selects the test fixture explicitly")


class _FKFunction(torch.autograd.Function):
    """Keypoints from (q, rot6d, trans) with gradients: forward = the FK kernel, backward =
    `hrp_fk_backward` (analytic geometric Jacobian, rigid root inverse, 6-D rotation adjoint) -- what torch autograd
    computes through URDFRobot.get_keypoints[_root] in the reference's training losses (lib/core/function.py:253-311)."""

    @staticmethod
    def forward(ctx, robot, root, q, rot, trans):
        use_b2c = rot is not None
        with torch.no_grad():
            pts = robot._fk_raw(q, rot, trans, root=root)[0]
        ctx.robot, ctx.root, ctx.use_b2c = robot, int(root), use_b2c
        ctx.save_for_backward(q, rot if use_b2c else q.new_zeros(0), trans if use_b2c else q.new_zeros(0))
        return pts

    @staticmethod
    def backward(ctx, grad_pts):
        q, rot, trans = ctx.saved_tensors
        robot, use_b2c = ctx.robot, ctx.use_b2c
        q32 = _f32(q)
        g = _f32(grad_pts)
        B = q32.shape[0]
        gq = torch.empty(B, robot.dof, dtype=torch.float32, device=q32.device)
        grot = torch.empty(B, 6, dtype=torch.float32, device=q32.device) if use_b2c else None
        gtr = torch.empty(B, 3, dtype=torch.float32, device=q32.device) if use_b2c else None
        r32 = _f32(rot) if use_b2c else None
        t32 = _f32(trans) if use_b2c else None
        with torch.cuda.device(q32.device):
            check(_lib.lib().hrp_fk_backward(
                robot.handle(q32.device), C.c_void_p(q32.data_ptr()), C.c_void_p(r32.data_ptr() if use_b2c else 0),
                C.c_void_p(t32.data_ptr() if use_b2c else 0), ctx.root, int(use_b2c), C.c_void_p(g.data_ptr()),
                C.c_void_p(gq.data_ptr()), C.c_void_p(grot.data_ptr() if use_b2c else 0),
                C.c_void_p(gtr.data_ptr() if use_b2c else 0), B, _stream()))
        return None, None, gq, grot, gtr


class _ProjectFunction(torch.autograd.Function):
    """point_projection_from_3d_tensor with a gradient to the points (hrp_project_backward)."""

    @staticmethod
    def forward(ctx, K, pts):
        ctx.save_for_backward(K, pts)
        with torch.no_grad():
            return _project_raw(K, pts)

    @staticmethod
    def backward(ctx, grad_uv):
        K, pts = ctx.saved_tensors
        K32, p32, g = _f32(K), _f32(pts), _f32(grad_uv)
        B, N = p32.shape[0], p32.shape[1]
        gp = torch.empty(B, N, 3, dtype=torch.float32, device=p32.device)
        with torch.cuda.device(p32.device):
            check(_lib.lib().hrp_project_backward(C.c_void_p(K32.data_ptr()), C.c_void_p(p32.data_ptr()),
                                                  C.c_void_p(g.data_ptr()), C.c_void_p(gp.data_ptr()), B, N, _stream()))
        return None, gp


class _LinkFkView:
    """`URDFRobot.robot.link_fk_batch(cfgs, use_names=True)` of the reference (urdf_robot.py:108 ->
    urdfpytorch/urdf.py:3061-3149): transforms of EVERY link of the description, computed by one CUDA kernel."""

    def __init__(self, owner):
        self._owner = owner

    def link_fk_batch(self, cfgs, use_names=True):
        T = self._owner._link_fk_all(cfgs)
        names = self._owner.tree.link_names
        if use_names:
            return {n: T[:, i] for i, n in enumerate(names)}
        raise NotImplementedError("link_fk_batch(use_names=False) returns Link objects in the reference (off-path)")


class URDFRobot:
    def __init__(self, robot_type: str, urdf_path: str | None = None):
        if robot_type not in _DOF:
            raise NotImplementedError(f"robot type {robot_type!r} is outside the hot path (panda / kuka / baxter)")
        self.robot_type = robot_type
        self.urdf_path = resolve_urdf_path(robot_type, urdf_path)
        self.dof = _DOF[robot_type]
        self.tree = urdf.load_urdf(self.urdf_path)
        self.actuated_joint_names = tables.JOINT_NAMES[robot_type]
        if self.tree.actuated_joint_names != self.actuated_joint_names:
            # q[:, i] drives tree.actuated_joint_names[i] (urdf.py:3933-3934); the model was trained in
            # JOINT_NAMES order, so a URDF whose depth-sorted order differs cannot be used silently.
            raise ValueError(f"actuated joints of {self.urdf_path} are {self.tree.actuated_joint_names}, expected "
                             f"{self.actuated_joint_names}")
        self.global_scale = 1.0
        self.link_names, offsets = self._link_names_and_offsets()
        self._offsets_np = offsets
        self.offsets = torch.as_tensor(offsets, dtype=torch.float32).unsqueeze(0).unsqueeze(-1)  # (1,nkpt,3,1)
        self._handles = {}   # device index -> hrp_robot handle (tables live in that device's memory)
        self.robot = _LinkFkView(self)

    # urdf_robot.py:52-80
    def _link_names_and_offsets(self):
        if self.robot_type in ("panda", "kuka"):
            names = list(tables.LINK_NAMES[self.robot_type])
            return names, np.zeros((len(names), 3), dtype=np.float64)
        names, offs = [], []
        for jn in tables.BAXTER_KEYPOINT_JOINTS:
            j = self.tree.joints[jn]
            names.append(j.parent)
            offs.append(j.origin[:3, 3])
        return names, np.stack(offs)

    # ---- table upload ---------------------------------------------------------------------------------
    def _pruned_rows(self):
        t = self.tree
        keep = set()
        for n in self.link_names:
            i = t.link_index(n)
            while i >= 0 and i not in keep:
                keep.add(i)
                i = t.parent[i]
        rows, remap = self._rows(sorted(keep))
        kp = (C.c_int32 * len(self.link_names))(*[remap[t.link_index(n)] for n in self.link_names])
        return rows, kp

    def _rows(self, order):
        t = self.tree
        remap = {old: new for new, old in enumerate(order)}
        rows = (LinkRow * len(order))()
        for new, old in enumerate(order):
            r = rows[new]
            r.parent = remap[t.parent[old]] if t.parent[old] >= 0 else -1
            r.jtype, r.qcol, r.qmul, r.qoff = t.jtype[old], t.qcol[old], t.qmul[old], t.qoff[old]
            r.origin[:] = list(np.asarray(t.origin[old], dtype=np.float64).reshape(-1))
            r.axis[:] = list(np.asarray(t.axis[old], dtype=np.float64))
        return rows, remap

    def handle(self, device=None):
        """hrp_robot handle whose tables live on `device` (default: the current CUDA device)."""
        dev = torch.cuda.current_device() if device is None else torch.device(device).index
        if dev is None:
            dev = torch.cuda.current_device()
        h = self._handles.get(dev)
        if h is None:
            rows, kp = self._pruned_rows()
            off = np.ascontiguousarray(self._offsets_np * self.global_scale, dtype=np.float64)
            h = C.c_void_p(0)
            with torch.cuda.device(dev):
                check(_lib.lib().hrp_robot_create(rows, len(rows), kp, off.ctypes.data_as(C.POINTER(C.c_double)),
                                                  len(self.link_names), self.dof, C.byref(h)))
                full, _ = self._rows(list(range(len(self.tree.link_names))))
                check(_lib.lib().hrp_robot_set_full_tree(h, full, len(full)))
            self._handles[dev] = h
        return h

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().hrp_robot_destroy(h)
            self._handles = {}
        except Exception:
            pass

    # ---- kernels --------------------------------------------------------------------------------------
    def _fk(self, q, rot=None, trans=None, root=0, want_pts=True, want_rot=False):
        needs_grad = torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad
                                                     for t in (q, rot, trans))
        if needs_grad and want_pts and not want_rot:
            if rot is not None and rot.shape[1] != 6:
                raise NotImplementedError("gradients are implemented for the 6-D rotation representation (rotation_dim 6)")
            return _FKFunction.apply(self, root, q, rot, trans), None
        return self._fk_raw(q, rot, trans, root=root, want_pts=want_pts, want_rot=want_rot)

    def _fk_raw(self, q, rot=None, trans=None, root=0, want_pts=True, want_rot=False):
        q = _f32(q)
        B = q.shape[0]
        assert q.shape[1] == self.dof, (q.shape, self.dof)
        nk = len(self.link_names)
        use_b2c = rot is not None
        rot_dim = 0
        if use_b2c:
            rot, trans = _f32(rot), _f32(trans)
            rot_dim = rot.shape[1]
            if rot_dim not in (4, 6, 9):
                raise NotImplementedError
        pts = torch.empty(B, nk, 3, dtype=torch.float32, device=q.device) if want_pts else None
        rout = torch.empty(B, rot_dim, dtype=torch.float32, device=q.device) if want_rot else None
        with torch.cuda.device(q.device):
            check(_lib.lib().hrp_fk(self.handle(q.device), C.c_void_p(q.data_ptr()),
                                    C.c_void_p(rot.data_ptr() if use_b2c else 0), rot_dim,
                                    C.c_void_p(trans.data_ptr() if use_b2c else 0), int(root), int(use_b2c),
                                    C.c_void_p(pts.data_ptr() if pts is not None else 0),
                                    C.c_void_p(rout.data_ptr() if rout is not None else 0), B, _stream()))
        return pts, rout

    def _link_fk(self, q, all_links: bool):
        q = _f32(q)
        B = q.shape[0]
        assert q.shape[1] == self.dof, (q.shape, self.dof)
        L = len(self.tree.link_names) if all_links else len(self.link_names)
        out = torch.empty(B, L, 4, 4, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            check(_lib.lib().hrp_link_fk(self.handle(q.device), C.c_void_p(q.data_ptr()), B, int(all_links),
                                         C.c_float(self.global_scale), C.c_void_p(out.data_ptr()), _stream()))
        return out

    def _link_fk_all(self, q):
        return self._link_fk(q, True)

    # ---- reference API (urdf_robot.py) ----------------------------------------------------------------
    def get_TWL(self, cfgs):  # :107-111 -> (B, nkpt, 4, 4)
        return self._link_fk(cfgs, False)

    def get_keypoints(self, jointcfgs, b2c_rot, b2c_trans):  # :82-105
        return self._fk(jointcfgs, b2c_rot, b2c_trans, root=0)[0]

    def get_keypoints_root(self, jointcfgs, b2c_rot, b2c_trans, root=0):  # :169-199
        assert 0 <= root < len(self.link_names)
        return self._fk(jointcfgs, b2c_rot, b2c_trans, root=root)[0]

    def get_keypoints_only_fk(self, jointcfgs):  # :141-149
        return self._fk(jointcfgs)[0]

    def get_keypoints_only_fk_at_specific_root(self, jointcfgs, root=0):  # :151-166
        assert 0 <= root < len(self.link_names)
        return self._fk(jointcfgs, root=root)[0]

    def get_rotation_at_specific_root(self, jointcfgs, b2c_rot, b2c_trans, root=0):  # :113-138
        if root == 0:
            return b2c_rot
        assert root < len(self.link_names), (root, len(self.link_names))
        return self._fk(jointcfgs, b2c_rot, b2c_trans, root=root, want_pts=False, want_rot=True)[1]


def point_projection_from_3d_tensor(camera_K, points):
    """transforms.py:17-21: (B,3,3), (B,N,3) -> (B,N,2) on the GPU (differentiable w.r.t. the points)."""
    if torch.is_grad_enabled() and torch.is_tensor(points) and points.requires_grad:
        return _ProjectFunction.apply(camera_K, points)
    return _project_raw(camera_K, points)


def _project_raw(camera_K, points):
    K, pts = _f32(camera_K), _f32(points)
    B, N = pts.shape[0], pts.shape[1]
    uv = torch.empty(B, N, 2, dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        check(_lib.lib().hrp_project(C.c_void_p(K.data_ptr()), C.c_void_p(pts.data_ptr()), C.c_void_p(uv.data_ptr()),
                                     B, N, _stream()))
    return uv


def point_projection_from_3d(camera_K, points):
    """transforms.py:11-15 (numpy in, numpy out) -- computed by the same CUDA kernel."""
    K = torch.as_tensor(np.asarray(camera_K), dtype=torch.float32).cuda()
    pts = torch.as_tensor(np.asarray(points), dtype=torch.float32).cuda()
    return point_projection_from_3d_tensor(K, pts).cpu().numpy()
