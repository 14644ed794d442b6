"""Deterministic synthetic weights and inputs ("random-init weights of that architecture, synthetic inputs of
the config's shape" -- BASELINE.json north_star; recipe per SURVEY.md section 8d / 9).

Everything is generated from integer bit streams (numpy PCG64 `random_raw`) and single IEEE fp32 operations, so
the container that produced tests/golden/*.npz and the GPU box regenerate bit-identical tensors.  There is no
network access for checkpoints or datasets; `data: "synthetic"` in bench.py refers to this module.

Recipe (one per robot, backbones shared):
  * conv / deconv / linear weights: uniform, zero mean, std = gain*sqrt(2/fan_in) (He) -- He-normal in the
    reference (full_net.py:167-173) differs only in the shape of the distribution;
  * BatchNorm gamma ~ U[0.5,1.5] (x gamma_res on the last BN of every residual block),
    beta ~ U[-0.2,0.2];
  * BatchNorm running statistics: calibrated ONCE on seeded random images with the oracle and committed as
    fixtures/bn_stats.npz (tests/golden/make_golden.py) -- random-init nets explode in eval mode otherwise
    (SURVEY.md fact 7);
  * regression heads with the reference's own init scales (nn.Linear default for fc_*, xavier gain 0.01 for
    decpose / decrot, full_net.py:95-100,129-134; depth_layer normal std 1e-3, :174-177): the bf16 noise of the
    2048-d features reaches pose / rot / depth through these weights, so their scale sets the parity margin;
  * depth_layer.bias = 2.0 so that depth = gamma*k/1000 is a plausible 0.7..2.5 m.
"""
from __future__ import annotations

import zlib
from pathlib import Path

import numpy as np
import torch

from . import arch

FIXTURES = Path(__file__).resolve().parent / "fixtures"
BN_STATS_PATH = FIXTURES / "bn_stats.npz"
URDF_PATHS = {
    "panda": FIXTURES / "urdf" / "panda.urdf",
    "kuka": FIXTURES / "urdf" / "iiwa7.urdf",
    "baxter": FIXTURES / "urdf" / "baxter.urdf",
}
GAMMA_RES = 0.25


def use_synthetic_urdfs() -> None:
    """Point $HRP_URDF_<ROBOT> at the mesh-free URDF fixtures (tests, bench.py, smoke): `URDFRobot(robot_type)` and
    `RootNetwithRegInt(init, args)` with the reference signatures then resolve them.  An explicit opt-in: without it
    (and without a real description) constructing a robot raises instead of silently using approximate kinematics."""
    import os
    for rt, path in URDF_PATHS.items():
        os.environ.setdefault(f"HRP_URDF_{rt.upper()}", str(path))


def uniform01(key: str, n: int, seed: int = 0) -> np.ndarray:
    """n floats in [0,1) with 24 random bits each; depends only on (key, seed)."""
    s = (zlib.crc32(key.encode()) << 32) | (seed & 0xFFFFFFFF)
    raw = np.random.PCG64(s).random_raw(n)
    return (raw >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))


def sym_uniform(key: str, shape, std: float, seed: int = 0) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    u = uniform01(key, n, seed)
    a = np.float32(std * np.sqrt(3.0))
    v = (u * np.float32(2.0) - np.float32(1.0)) * a
    return torch.from_numpy(v.reshape(shape))


def range_uniform(key: str, shape, lo: float, hi: float, seed: int = 0) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    u = uniform01(key, n, seed)
    v = u * np.float32(hi - lo) + np.float32(lo)
    return torch.from_numpy(v.reshape(shape))


def _fill(spec, seed: int, with_bn_stats: bool):
    sd = {}
    res_last = set(arch.residual_last_bn_names(spec))
    stats = None
    if with_bn_stats:
        if not BN_STATS_PATH.exists():
            raise FileNotFoundError(f"{BN_STATS_PATH} missing: run tests/golden/make_golden.py --calibrate")
        stats = np.load(BN_STATS_PATH)
    for key, shape in spec.items():
        base, _, leaf = key.rpartition(".")
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif leaf in ("running_mean", "running_var"):
            if stats is not None:
                sd[key] = torch.from_numpy(stats[_stats_key(key)].astype(np.float32))
            else:
                sd[key] = torch.zeros(shape) if leaf == "running_mean" else torch.ones(shape)
        elif len(shape) == 4:  # conv / deconv weight
            if key.startswith("deconv_layers."):
                fan_in = shape[0] * 4  # ConvTranspose2d (Cin,Cout,4,4): 2x2 taps reach each output pixel
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            gain = 1.0
            if key == "depth_layer.weight":  # reference: normal(std=0.001), full_net.py:174-177
                sd[key] = sym_uniform(key, shape, 0.001, seed)
                continue
            sd[key] = sym_uniform(key, shape, gain * float(np.sqrt(2.0 / fan_in)), seed)
        elif len(shape) == 2 and leaf == "weight":  # linear: the reference's own init scales
            if base in ("decpose", "decrot"):  # xavier_uniform(gain=0.01), full_net.py:100,134
                std = 0.01 * float(np.sqrt(2.0 / (shape[0] + shape[1])))
            else:  # nn.Linear default: U(+-1/sqrt(fan_in))
                std = float(np.sqrt(1.0 / (3.0 * shape[1])))
            sd[key] = sym_uniform(key, shape, std, seed)
        elif leaf == "weight":  # BN gamma
            g = range_uniform(key, shape, 0.5, 1.5, seed)
            sd[key] = g * GAMMA_RES if base in res_last else g
        elif leaf == "bias":
            if key == "depth_layer.bias":
                sd[key] = torch.full(shape, 2.0)
            elif (base + ".running_mean") in spec:  # BN beta
                sd[key] = range_uniform(key, shape, -0.2, 0.2, seed)
            elif len(spec.get(base + ".weight", ())) == 2:  # linear bias: U(+-1/sqrt(fan_in)) like nn.Linear
                sd[key] = sym_uniform(key, shape, float(np.sqrt(1.0 / (3.0 * spec[base + ".weight"][1]))), seed)
            else:  # conv bias
                sd[key] = sym_uniform(key, shape, 0.05, seed)
        else:
            raise KeyError(key)
    return sd


def _stats_key(key: str) -> str:
    # the two backbones are shared by every robot and by the depthnet: strip the owner prefix variations
    return key.replace("backbone.", "B.", 1) if key.startswith("backbone.") else key.replace(
        "rootnet_backbone.", "B.", 1)


def full_state_dict(robot_type: str, seed: int = 0, with_bn_stats: bool = True):
    """Reference-keyed `state_dict` of RootNetwithRegInt (resnet50 + hrnet32) for `robot_type`."""
    spec = arch.full_model_spec(robot_type)
    sd = _fill({k: v for k, v in spec.items() if k not in ("init_pose", "init_rot")}, seed, with_bn_stats)
    from .tables import INIT_POSE_MEAN
    sd["init_pose"] = torch.tensor([INIT_POSE_MEAN[robot_type]], dtype=torch.float32)
    sd["init_rot"] = torch.tensor([[1.0, 0.0, 0.0, 0.0, 1.0, 0.0]], dtype=torch.float32)
    return sd


def depthnet_state_dict(seed: int = 0, with_bn_stats: bool = True):
    """`RootNet('hrnet32')` state dict; backbone weights equal the full model's rootnet_backbone (the reference
    remaps backbone.* -> rootnet_backbone.*, full_net.py:423-427)."""
    spec = arch.depthnet_spec()
    renamed = {k.replace("backbone.", "rootnet_backbone.", 1) if k.startswith("backbone.") else k: v
               for k, v in spec.items()}
    sd = _fill(renamed, seed, with_bn_stats)
    return {(k.replace("rootnet_backbone.", "backbone.", 1) if k.startswith("rootnet_backbone.") else k): v
            for k, v in sd.items()}


def inputs(batch: int, seed: int = 1, k_range=(500.0, 1500.0)):
    """x_reg, x_root (B,3,256,256) in [0,1); k_value (B,); K (B,3,3)  -- SURVEY.md section 8d C1."""
    x_reg = torch.from_numpy(uniform01("x_reg", batch * 3 * 256 * 256, seed).reshape(batch, 3, 256, 256))
    x_root = torch.from_numpy(uniform01("x_root", batch * 3 * 256 * 256, seed).reshape(batch, 3, 256, 256))
    k_value = range_uniform("k_value", (batch,), k_range[0], k_range[1], seed)
    f = range_uniform("focal", (batch,), 300.0, 800.0, seed)
    cx = range_uniform("cx", (batch,), 107.5, 147.5, seed)
    cy = range_uniform("cy", (batch,), 107.5, 147.5, seed)
    K = torch.zeros(batch, 3, 3)
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = cx
    K[:, 1, 2] = cy
    K[:, 2, 2] = 1.0
    return x_reg, x_root, k_value, K


def fk_inputs(robot_type: str, batch: int, seed: int = 2):
    """q ~ U[JOINT_BOUNDS], rot ~ zero-mean 6-vector, trans = (x, y, z in [0.8,2.5]) -- SURVEY.md section 8d."""
    from .tables import JOINT_BOUNDS
    b = np.asarray(JOINT_BOUNDS[robot_type], dtype=np.float32)
    u = uniform01("fk_q_" + robot_type, batch * b.shape[0], seed).reshape(batch, -1)
    q = torch.from_numpy(u * (b[:, 1] - b[:, 0]) + b[:, 0])
    rot = sym_uniform("fk_rot_" + robot_type, (batch, 6), 1.0, seed)
    trans = torch.cat([sym_uniform("fk_txy_" + robot_type, (batch, 2), 0.3, seed),
                       range_uniform("fk_tz_" + robot_type, (batch, 1), 0.8, 2.5, seed)], dim=1)
    return q, rot, trans


# fixed bounding boxes of the crop fixtures (wmin, hmin, wmax, hmax) in a 640 x 480 frame: wide / tall / exactly the
# 256-pixel target (identity branch, augmentations.py:174-176) / small (up-sampling) / touching the frame edges /
# odd sizes / the whole frame height
CROP_BOXES = [(100, 50, 500, 300), (250, 10, 400, 470), (192, 112, 448, 368), (300, 200, 357, 243), (0, 0, 640, 480),
              (5, 3, 262, 480), (383, 223, 640, 480), (10, 100, 266, 300), (200, 100, 300, 356), (17, 31, 600, 333)]


def crop_inputs(batch: int, seed: int = 5, frame_hw=(480, 640)):
    """Frames (B,H,W,3) uint8 (triangle-wave gradient + noise, so the bilinear resize is exercised on non-trivial data),
    bboxes (B,4) int32 -- CROP_BOXES first, then seeded random boxes --, K (B,3,3) float64, k_bbox (B,4) float32."""
    H, W = frame_hw
    n = batch * H * W * 3
    noise = uniform01("crop_frames", n, seed).reshape(batch, H, W, 3)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64), indexing="ij")
    tri = np.abs(((xx * 5 + yy * 3) % 512) - 256).astype(np.float32)[None, :, :, None]  # integer triangle wave, 0..256
    frames = np.clip(np.float32(0.6) * tri + np.float32(0.4 * 256.0) * noise, 0, 255).astype(np.uint8)
    boxes = []
    u = uniform01("crop_boxes", batch * 4, seed).reshape(batch, 4)
    for i in range(batch):
        if i < len(CROP_BOXES):
            boxes.append(CROP_BOXES[i])
            continue
        w = 40 + int(u[i, 0] * (W - 40))
        h = 40 + int(u[i, 1] * (H - 40))
        x0 = int(u[i, 2] * (W - w + 1))
        y0 = int(u[i, 3] * (H - h + 1))
        boxes.append((x0, y0, x0 + w, y0 + h))
    boxes = np.asarray(boxes, dtype=np.int32)
    K = np.zeros((batch, 3, 3), dtype=np.float64)
    f = uniform01("crop_f", batch, seed).astype(np.float64) * 100.0 + 560.0
    K[:, 0, 0] = f
    K[:, 1, 1] = f + 0.25
    K[:, 0, 2] = 320.0 + uniform01("crop_cx", batch, seed).astype(np.float64) * 8.0
    K[:, 1, 2] = 240.0 + uniform01("crop_cy", batch, seed).astype(np.float64) * 8.0
    K[:, 2, 2] = 1.0
    k_bbox = boxes.astype(np.float32) + uniform01("crop_kb", batch * 4, seed).reshape(batch, 4) * 3.0
    return frames, boxes, K, k_bbox.astype(np.float32)


def metric_inputs(robot_type: str, batch: int, seed: int = 6):
    """Predictions and ground truth for the metric kernels: FK inputs (pred) and a perturbed copy (gt), the camera
    matrix of the 640 x 480 frame, gt 2-D keypoints partly outside the frame (exercises the valid mask)."""
    q, rot, trans = fk_inputs(robot_type, batch, seed)
    from .tables import LINK_NAMES
    nkpt = {"panda": 7, "kuka": 8, "baxter": 17}[robot_type]
    assert robot_type == "baxter" or len(LINK_NAMES[robot_type]) == nkpt
    gt_q = q + sym_uniform("m_dq_" + robot_type, tuple(q.shape), 0.05, seed)
    gt3 = torch.cat([sym_uniform("m_xy_" + robot_type, (batch, nkpt, 2), 0.4, seed),
                     range_uniform("m_z_" + robot_type, (batch, nkpt, 1), 0.8, 2.5, seed)], dim=2)
    K = torch.zeros(batch, 3, 3)
    K[:, 0, 0] = range_uniform("m_fx", (batch,), 560.0, 660.0, seed)
    K[:, 1, 1] = range_uniform("m_fy", (batch,), 560.0, 660.0, seed)
    K[:, 0, 2] = 320.0
    K[:, 1, 2] = 240.0
    K[:, 2, 2] = 1.0
    gt2 = torch.stack([range_uniform("m_u_" + robot_type, (batch, nkpt), -60.0, 700.0, seed),
                       range_uniform("m_v_" + robot_type, (batch, nkpt), -60.0, 540.0, seed)], dim=2)
    return q, rot, trans, gt_q, gt3, gt2, K


def pnp_inputs(robot_type: str, batch: int, seed: int = 8, noise_px: float = 0.5):
    """2-D / 3-D correspondences for the PnP kernel: base-frame FK keypoints of random joint states (computed by the
    caller from `q`), a random camera pose in front of the camera, pixel noise on the projections.
    Returns q (B,dof), rvec (B,3) angle-axis with |r| in [0.1, 3], t (B,3), K (3,3), noise (B,nkpt,2)."""
    from .tables import JOINT_BOUNDS
    b = np.asarray(JOINT_BOUNDS[robot_type], dtype=np.float32)
    nkpt = {"panda": 7, "kuka": 8, "baxter": 17}[robot_type]
    u = uniform01("pnp_q_" + robot_type, batch * b.shape[0], seed).reshape(batch, -1)
    q = torch.from_numpy(u * (b[:, 1] - b[:, 0]) + b[:, 0])
    d = sym_uniform("pnp_dir_" + robot_type, (batch, 3), 1.0, seed)
    d = d / d.norm(dim=1, keepdim=True).clamp_min(1e-3)
    rvec = d * range_uniform("pnp_ang_" + robot_type, (batch, 1), 0.1, 3.0, seed)
    t = torch.cat([sym_uniform("pnp_txy_" + robot_type, (batch, 2), 0.2, seed),
                   range_uniform("pnp_tz_" + robot_type, (batch, 1), 1.2, 2.5, seed)], dim=1)
    K = torch.tensor([[615.0, 0.0, 320.0], [0.0, 615.5, 240.0], [0.0, 0.0, 1.0]])
    noise = sym_uniform("pnp_noise_" + robot_type, (batch, nkpt, 2), noise_px, seed)
    return q, rvec, t, K, noise
