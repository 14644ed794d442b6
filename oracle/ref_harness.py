"""Run the UNMODIFIED reference (read-only /root/reference) on CPU -- TEST INFRASTRUCTURE.

The reference is pure Python but cannot be imported as-is: it hard-codes `.cuda()`,
downloads ImageNet weights, opens cwd-relative files, and imports packages absent from this image.  This module
builds a throw-away sandbox directory with symlinks to the reference's `lib/` and `configs/`, installs stub
modules for the missing imports, and applies the minimal shims listed below.  No reference source is copied.

It exists in the build container only (the GPU box has no /root/reference): it validates oracle/horopose_oracle.py
and generates tests/golden/*.npz (see tests/golden/make_golden.py).  `available()` tells callers whether the
reference is present.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("HRP_REFERENCE_ROOT", "/root/reference"))
_state = {}


def available() -> bool:
    return (REFERENCE_ROOT / "lib" / "models" / "full_net.py").exists()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def setup(urdf_paths: dict):
    """Create the sandbox, chdir into it and import the reference.  `urdf_paths`: robot_type -> URDF file.
    Returns a namespace with the reference's entry points."""
    if _state:
        return _state["ns"]
    import numpy as np
    import torch

    assert available(), f"reference not found under {REFERENCE_ROOT}"
    sandbox = Path(tempfile.mkdtemp(prefix="hrp_ref_sandbox_"))
    os.symlink(REFERENCE_ROOT / "lib", sandbox / "lib")
    os.symlink(REFERENCE_ROOT / "configs", sandbox / "configs")
    (sandbox / "data").mkdir()
    (sandbox / "models").mkdir()
    # URDF locations the reference expects (lib/config.py:33-36)
    deps = sandbox / "data" / "deps"
    (deps / "panda-description" / "patched_urdf").mkdir(parents=True)
    (deps / "kuka-description" / "iiwa_description" / "urdf").mkdir(parents=True)
    (deps / "baxter-description").mkdir(parents=True)
    import shutil
    shutil.copy(urdf_paths["panda"], deps / "panda-description" / "panda.urdf")
    shutil.copy(urdf_paths["panda"], deps / "panda-description" / "patched_urdf" / "panda.urdf")
    shutil.copy(urdf_paths["kuka"], deps / "kuka-description" / "iiwa_description" / "urdf" / "iiwa7.urdf")
    shutil.copy(urdf_paths["baxter"], deps / "baxter-description" / "baxter.urdf")
    os.chdir(sandbox)
    sys.path[:0] = [str(sandbox / "lib"), str(sandbox)]

    # --- stubs for imports that are absent here and unused by the hot path -------------------------------
    class EasyDict(dict):  # HRnet.py:16, core/config.py
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

    _stub("easydict", EasyDict=EasyDict)
    import xml.etree.ElementTree as _ET

    class _XMLParser(_ET.XMLParser):  # lxml accepts remove_comments/remove_blank_text (urdf.py:3737-3739)
        def __init__(self, **kw):
            super().__init__()

    lxml = _stub("lxml")
    lxml.etree = _stub("lxml.etree", **{k: getattr(_ET, k) for k in dir(_ET) if not k.startswith("_")})
    lxml.etree.XMLParser = _XMLParser
    _stub("pyrender")
    _stub("trimesh")

    class _Fake:  # urdf_robot.py:29,34,39 instantiates PandaArm for rendering only
        def __init__(self, *a, **k):
            pass

    _stub("utils.mesh_renderer", PandaArm=_Fake, RobotMeshRenderer=_Fake)
    # --- shims -------------------------------------------------------------------------------------------
    torch.Tensor.cuda = lambda self, *a, **k: self            # SURVEY.md fact 3 (hard-coded .cuda())
    if not hasattr(np, "infty"):
        np.infty = np.inf

    import utils  # noqa: F401  (package lib/utils; our stub utils.mesh_renderer must be visible as attribute)
    utils.mesh_renderer = sys.modules["utils.mesh_renderer"]
    import utils.urdf_robot as urdf_robot
    urdf_robot.BAXTER_DESCRIPTION_PATH = str(deps / "baxter-description" / "baxter.urdf")  # lib/config.py:36
    from utils.urdfpytorch import urdf as urdfmod

    def _stable_sort(self, joints):  # urdf.py:3795-3813 with kind="stable" 
        lens = [len(self._paths_to_base[self._link_map[j.child]]) for j in joints]
        order = np.argsort(lens, kind="stable")
        return np.array(joints)[order].tolist()

    urdfmod.URDF._sort_joints = _stable_sort

    import models.backbones.Resnet as resnet_mod
    import models.backbones.HRnet as hrnet_mod
    resnet_mod.ResNet.init_weights = lambda self, name: None   # no ImageNet download (Resnet.py:69-92)
    _orig_init = hrnet_mod.PoseHighResolutionNet.init_weights
    hrnet_mod.PoseHighResolutionNet.init_weights = lambda self, pretrained="": _orig_init(self, "")

    import models.full_net as full_net
    import models.depth_net as depth_net
    import utils.integral as integral
    import utils.transforms as transforms
    import utils.geometries as geometries
    from dataset.const import INITIAL_JOINT_ANGLE, JOINT_NAMES, LINK_NAMES, JOINT_BOUNDS

    ns = types.SimpleNamespace(
        sandbox=sandbox, full_net=full_net, depth_net=depth_net, urdf_robot=urdf_robot, integral=integral,
        transforms=transforms, geometries=geometries, INITIAL_JOINT_ANGLE=INITIAL_JOINT_ANGLE,
        JOINT_NAMES=JOINT_NAMES, LINK_NAMES=LINK_NAMES, JOINT_BOUNDS=JOINT_BOUNDS, EasyDict=EasyDict)
    _state["ns"] = ns
    return ns


def full_args(ns, robot_type: str):
    """The model-relevant fields of configs/<robot>/full.yaml merged over lib/core/config.py defaults."""
    import yaml
    cfg_file = REFERENCE_ROOT / "configs" / robot_type / "full.yaml"
    y = yaml.safe_load(open(cfg_file))
    a = ns.EasyDict(
        backbone_name=y["backbone_name"], rootnet_backbone_name=y["rootnet_backbone_name"], use_rpmg=False,
        n_iter=y.get("n_iter", 4), other_image_size=y["other_image_size"], bbox_3d_shape=y["bbox_3d_shape"],
        reference_keypoint_id=y["reference_keypoint_id"], fix_root=y.get("fix_root", True), rotation_dim=6,
        reg_joint_map=False, joint_conv_dim=[], p_dropout=0.5, direct_reg_rot=False, rot_iterative_matmul=False,
        multi_kp=False, kps_need_depth=None, add_fc=False, pretrained_rootnet=None)
    return a


def build_full_model(ns, robot_type: str):
    import numpy as np
    init = {"robot_type": robot_type, "pose_params": ns.INITIAL_JOINT_ANGLE, "cam_params": np.eye(4, dtype=float),
            "init_pose_from_mean": True}  # scripts/test.py:58-63
    model = ns.full_net.get_rootNetwithRegInt_model(init, full_args(ns, robot_type))
    return model.eval()


def build_depthnet(ns):
    return ns.depth_net.get_rootnet("hrnet32").eval()
