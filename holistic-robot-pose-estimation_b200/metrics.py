"""On-device evaluation metrics: drop-in for lib/utils/metrics.py.

`compute_metrics_batch` keeps the reference signature and the nine returned values (metrics.py:8-119) but leaves
them on the GPU as tensors (the reference returns numpy arrays / lists after a device sync per batch);
`summary_add_pck` (metrics.py:122-162) runs the two 10 000- / 2 000-threshold AUC scans, the medians and the
threshold table as one kernel.  `MetricAccumulator` is the device-resident `alldis` of scripts/test.py:199-232.
No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check

ADD_MM = (1, 5, 10, 20, 40, 60, 80, 100)
PCK_PX = (2.5, 5.0, 7.5, 10.0, 12.5, 15.0, 17.5, 20.0)


class MetricsArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("nkpt", C.c_int32), ("dof", C.c_int32), ("ref_kpt", C.c_int32),
                ("drop_last_joint", C.c_int32), ("frame_w", C.c_float), ("frame_h", C.c_float),
                ("pred_kp3d", C.c_void_p), ("gt_kp3d", C.c_void_p), ("gt_kp2d", C.c_void_p), ("K_original", C.c_void_p),
                ("pred_joint", C.c_void_p), ("gt_joint", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_int64), ("per_image", C.c_void_p), ("dis3d", C.c_void_p), ("dis2d", C.c_void_p),
                ("l1_jointerror", C.c_void_p)]


def _f32(t) -> torch.Tensor:
    t = torch.as_tensor(t)
    if not t.is_cuda:
        raise _lib.HrpError("horopose_b200 has no CPU path: tensors must live on a CUDA device")
    return t.detach().to(torch.float32).contiguous()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def compute_metrics_batch(robot, gt_keypoints3d, gt_keypoints2d, K_original, gt_joint, **pred_kwargs):
    """metrics.py:8-119.  `robot` is a horopose_b200.robot.URDFRobot.  Returns the reference's 9-tuple
    (error3d, error2d, dis3d, dis2d, l1_jointerror, mean_jointerror, error_depth, batch_error_relative,
    error3d_relative) as CUDA tensors."""
    pred_joint, pred_rot, pred_trans = pred_kwargs["pred_joint"], pred_kwargs["pred_rot"], pred_kwargs["pred_trans"]
    if pred_kwargs.get("pred_xy") is not None and pred_kwargs.get("pred_depth") is not None:
        pred_trans = torch.cat((pred_kwargs["pred_xy"], pred_kwargs["pred_depth"]), dim=-1)
    ref_id = int(pred_kwargs["reference_keypoint_id"])
    if pred_joint is None or pred_rot is None or pred_trans is None:
        assert pred_kwargs["pred_xyz_integral"] is not None
        pred_kp3d = _f32(pred_kwargs["pred_xyz_integral"])
        pred_joint = None
    elif ref_id == 0:
        pred_kp3d = robot.get_keypoints(pred_joint, pred_rot, pred_trans)
    else:
        pred_kp3d = robot.get_keypoints_root(pred_joint, pred_rot, pred_trans, root=ref_id)
    pred_kp3d = _f32(pred_kp3d)
    B, nkpt, dof = pred_kp3d.shape[0], len(robot.link_names), robot.dof
    gt3, gt2, Ko = _f32(gt_keypoints3d), _f32(gt_keypoints2d), _f32(K_original)
    assert pred_kp3d.shape == (B, nkpt, 3), f"{pred_kp3d.shape}"
    assert gt3.shape == (B, nkpt, 3), f"{gt3.shape}"
    assert gt2.shape == (B, nkpt, 2), f"{gt2.shape}"
    dev = pred_kp3d.device
    pj = gj = None
    if pred_joint is not None:
        pj, gj = _f32(pred_joint), _f32(gt_joint)
        assert gj.shape == pj.shape == (B, dof), f"{pj.shape},{gj.shape}"
    lib = _lib.lib()
    need = C.c_int64()
    check(lib.hrp_metrics_workspace_bytes(B, nkpt, dof, C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    per_image = torch.empty(6, B, dtype=torch.float32, device=dev)
    dis3d = torch.empty(nkpt, dtype=torch.float32, device=dev)
    dis2d = torch.empty(nkpt, dtype=torch.float32, device=dev)
    l1 = torch.zeros(dof, dtype=torch.float32, device=dev)
    a = MetricsArgs(B, nkpt, dof, ref_id, 1 if robot.robot_type == "panda" else 0, 640.0, 480.0,
                    pred_kp3d.data_ptr(), gt3.data_ptr(), gt2.data_ptr(), Ko.data_ptr(),
                    pj.data_ptr() if pj is not None else None, gj.data_ptr() if gj is not None else None,
                    ws.data_ptr(), need.value, per_image.data_ptr(), dis3d.data_ptr(), dis2d.data_ptr(),
                    l1.data_ptr() if pj is not None else None)
    with torch.cuda.device(dev):
        check(lib.hrp_metrics_batch(C.byref(a), _stream()))
    if pj is None:  # metrics.py:89-91
        per_image[2].zero_()
    return (per_image[0], per_image[1], dis3d, dis2d, l1, per_image[2], per_image[3], per_image[4], per_image[5])


def summary_add_pck(alldis) -> dict:
    """metrics.py:122-162.  alldis['dis3d'] / ['dis2d']: per-image errors (CUDA tensors, or lists of them)."""
    def cat(v):
        # the reference accumulates python floats / numpy values (scripts/test.py:199-232): those are uploaded once here;
        # CUDA tensors (or lists of them) stay on the device; CPU *tensors* are refused like everywhere else (no CPU path)
        def up(x):
            if torch.is_tensor(x):
                return _f32(x)
            if not torch.cuda.is_available():
                raise _lib.HrpError("horopose_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            return torch.as_tensor(x, dtype=torch.float32).to(torch.device("cuda", torch.cuda.current_device()))
        if isinstance(v, (list, tuple)):
            v = torch.cat([up(x).reshape(-1) for x in v])
        return _f32(up(v)).reshape(-1)
    d3, d2 = cat(alldis["dis3d"]), cat(alldis["dis2d"])
    assert d3.shape[0] == d2.shape[0]
    out = torch.empty(22, dtype=torch.float64, device=d3.device)
    with torch.cuda.device(d3.device):
        check(_lib.lib().hrp_metrics_summary(C.c_void_p(d3.data_ptr()), C.c_void_p(d2.data_ptr()), C.c_int64(d3.shape[0]),
                                             C.c_void_p(out.data_ptr()), _stream()))
    o = out.cpu().tolist()
    summary = {"ADD/mean": o[0], "ADD/median": o[1], "ADD/AUC": o[2],
               "ADD_2D/mean": o[11], "ADD_2D/median": o[12], "PCK/AUC": o[13]}
    for i, mm in enumerate(ADD_MM):
        summary[f"ADD_{mm}_mm"] = o[3 + i]
    for i, px in enumerate(PCK_PX):
        summary[f"PCK_{px}_pixel"] = o[14 + i]
    return summary


class MetricAccumulator:
    """Device-resident `alldis` (scripts/test.py:199-232): append per-batch results without a host sync."""

    def __init__(self):
        self.dis3d, self.dis2d, self.dis3d_relative = [], [], []

    def add(self, batch_result):
        self.dis3d.append(batch_result[0])
        self.dis2d.append(batch_result[1])
        self.dis3d_relative.append(batch_result[8])

    def summary(self):
        return summary_add_pck({"dis3d": self.dis3d, "dis2d": self.dis2d})

    def summary_relative(self):
        return summary_add_pck({"dis3d": self.dis3d_relative, "dis2d": self.dis2d})
