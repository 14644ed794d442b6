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


class URDFRobot:
    def __init__(self, robot_type: str, urdf_path: str | None = None):
        if robot_type not in _DOF:
            raise NotImplementedError(f"robot type {robot_type!r} is outside the hot path (panda / kuka / baxter)")
        self.robot_type = robot_type
        if urdf_path is None:
            urdf_path = os.environ.get(f"HRP_URDF_{robot_type.upper()}")
        if urdf_path is None:
            from . import synth
            urdf_path = str(synth.URDF_PATHS[robot_type])  # synthetic fixture (the reference ships no URDFs)
        self.urdf_path = str(urdf_path)
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
        self._handle = None
        self._device = None

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
        order = sorted(keep)
        remap = {old: new for new, old in enumerate(order)}
        rows = (LinkRow * len(order))()
        for new, old in enumerate(order):
            r = rows[new]
            r.parent = remap[t.parent[old]] if t.parent[old] >= 0 else -1
            r.jtype, r.qcol, r.qmul, r.qoff = t.jtype[old], t.qcol[old], t.qmul[old], t.qoff[old]
            r.origin[:] = list(np.asarray(t.origin[old], dtype=np.float64).reshape(-1))
            r.axis[:] = list(np.asarray(t.axis[old], dtype=np.float64))
        kp = (C.c_int32 * len(self.link_names))(*[remap[t.link_index(n)] for n in self.link_names])
        return rows, kp

    def handle(self):
        dev = torch.cuda.current_device()
        if self._handle is None or self._device != dev:
            rows, kp = self._pruned_rows()
            off = np.ascontiguousarray(self._offsets_np * self.global_scale, dtype=np.float64)
            h = C.c_void_p(0)
            check(_lib.lib().hrp_robot_create(rows, len(rows), kp, off.ctypes.data_as(C.POINTER(C.c_double)),
                                              len(self.link_names), self.dof, C.byref(h)))
            self._handle, self._device = h, dev
        return self._handle

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().hrp_robot_destroy(self._handle)
        except Exception:
            pass

    # ---- kernels --------------------------------------------------------------------------------------
    def _fk(self, q, rot=None, trans=None, root=0, want_pts=True, want_rot=False):
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
            check(_lib.lib().hrp_fk(self.handle(), C.c_void_p(q.data_ptr()),
                                    C.c_void_p(rot.data_ptr() if use_b2c else 0), rot_dim,
                                    C.c_void_p(trans.data_ptr() if use_b2c else 0), int(root), int(use_b2c),
                                    C.c_void_p(pts.data_ptr() if pts is not None else 0),
                                    C.c_void_p(rout.data_ptr() if rout is not None else 0), B, _stream()))
        return pts, rout

    # ---- reference API (urdf_robot.py) ----------------------------------------------------------------
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
    """transforms.py:17-21: (B,3,3), (B,N,3) -> (B,N,2) on the GPU."""
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
