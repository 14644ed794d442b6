"""Generate the committed fixtures from the REAL reference (run in the build container only).

  python tests/golden/make_golden.py --calibrate   # fixtures/bn_stats.npz (BN running stats of the synthetic recipe)
  python tests/golden/make_golden.py               # tests/golden/*.npz (reference outputs on seeded inputs)

Inputs and weights are regenerated from seeds by horopose_b200.synth on every machine; only the small OUTPUT
tensors of the reference are stored.  The reference is imported from /root/reference through
oracle/ref_harness.py (sandbox + stubs, no reference source is copied).
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import horopose_b200  # noqa: E402,F401
from horopose_b200 import arch, synth  # noqa: E402
from oracle import horopose_oracle as O  # noqa: E402
from oracle import eval_oracle as EO  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLDEN = Path(__file__).resolve().parent
torch.set_grad_enabled(False)

HEATMAP_STRESS_GAIN = 10.0  # "peaky heatmap" variant


def heatmap_logits(robot_type: str, batch: int, seed: int = 3, gain: float = 1.0) -> torch.Tensor:
    nkpt = arch.ROBOTS[robot_type][1]
    u = synth.uniform01("heatmap_" + robot_type, batch * nkpt * 64 * 64 * 64, seed)
    x = torch.from_numpy((u - np.float32(0.5)) * np.float32(6.0 * gain))
    return x.reshape(batch, nkpt * 64, 64, 64)


def calibrate():
    """BN running statistics = batch statistics of the synthetic-weight nets on 8 seeded random images, layer by
    layer (equivalent to the train-mode calibration pass of SURVEY.md fact 7), rounded to fp16-representable
    values so the fixture is compact and exact."""
    sd = synth.full_state_dict("panda", with_bn_stats=False)
    robot = O.OracleRobot("panda", str(synth.URDF_PATHS["panda"]))
    x_reg, x_root, k, K = synth.inputs(8, seed=7)
    stats = {}

    def calib(name, mean, var):
        mean = mean.half().float()
        var = var.clamp_min(1e-3).half().float()
        key = name.replace("rootnet_backbone.", "B.", 1)
        stats[key + ".running_mean"] = mean.numpy().astype(np.float16)
        stats[key + ".running_var"] = var.numpy().astype(np.float16)
        return mean, var

    out = O.full_forward(sd, robot, x_reg, x_root, k, K, calib=calib)
    np.savez_compressed(synth.BN_STATS_PATH, **stats)
    print("wrote", synth.BN_STATS_PATH, len(stats), "arrays,", synth.BN_STATS_PATH.stat().st_size, "bytes")
    print("calibration-pass depth:", out[4].flatten().tolist())


def to_np(t):
    return t.detach().cpu().numpy()


def golden_full(ns, robot_type: str, batch: int = 2):
    sd = synth.full_state_dict(robot_type)
    model = ref_harness.build_full_model(ns, robot_type)
    model.load_state_dict(sd, strict=True)   # proves the key/shape schema of horopose_b200.arch
    x_reg, x_root, k, K = synth.inputs(batch, seed=11)
    outs = model(x_reg, x_root, k, K)
    names = ["pose", "rot", "trans", "root_uv", "depth", "uvd", "xyz_int", "xyz_fk"]
    data = {n: to_np(o) for n, o in zip(names, outs)}
    # activation summaries for fault localisation (mean |x| of a few stages through forward hooks)
    taps = {}

    def hook(name):
        def fn(_m, _i, o):
            taps[name] = float(o.abs().mean())
        return fn

    hs = [model.reg_backbone.layer1.register_forward_hook(hook("reg_backbone.layer1")),
          model.reg_backbone.layer4.register_forward_hook(hook("reg_backbone.layer4")),
          model.rootnet_backbone.layer1.register_forward_hook(hook("rootnet_backbone.layer1")),
          model.rootnet_backbone.final_feat_layer.register_forward_hook(hook("rootnet_backbone.final_feat")),
          model.deconv_layers.register_forward_hook(hook("deconv")),
          model.final_layer.register_forward_hook(hook("heatmap"))]
    model(x_reg, x_root, k, K)
    for h in hs:
        h.remove()
    for n, v in taps.items():
        data["tap_absmean." + n] = np.float32(v)
    np.savez(GOLDEN / f"full_{robot_type}.npz", **data)
    # oracle agreement, printed for the record
    robot = O.OracleRobot(robot_type, str(synth.URDF_PATHS[robot_type]))
    o = O.full_forward(sd, robot, x_reg, x_root, k, K)
    print(f"[full {robot_type}] oracle vs reference max|diff|:",
          {n: float((a - b).abs().max()) for n, a, b in zip(names, o, outs)})
    print("   depth", data["depth"].flatten(), "pose[0]", data["pose"][0][:4], "taps", taps)


def golden_depthnet(ns, batch: int = 2):
    sd = synth.depthnet_state_dict()
    model = ref_harness.build_depthnet(ns)
    model.load_state_dict(sd, strict=True)
    _, x_root, k, _ = synth.inputs(batch, seed=11)
    out = model(x_root, k)
    np.savez(GOLDEN / "depthnet.npz", depth_mm=to_np(out))
    o = O.depthnet_forward(sd, x_root, k)
    print("[depthnet] oracle vs reference max|diff| (mm):", float((o - out).abs().max()), to_np(out).flatten())


def golden_fk(ns, robot_type: str, batch: int = 64):
    robot = ns.urdf_robot.URDFRobot(robot_type)
    q, rot, trans = synth.fk_inputs(robot_type, batch)
    _, _, _, K = synth.inputs(batch, seed=5)
    data = {
        "link_names": np.array(robot.link_names),
        "actuated_joint_names": np.array([j.name for j in robot.robot.actuated_joints]),
        "offsets": to_np(robot.offsets.squeeze(0).squeeze(-1)),
        "keypoints": to_np(robot.get_keypoints(q, rot, trans)),
        "keypoints_only_fk": to_np(robot.get_keypoints_only_fk(q)),
    }
    nk = len(robot.link_names)
    for root in sorted({0, 3, nk - 1}):
        data[f"keypoints_root{root}"] = to_np(robot.get_keypoints_root(q, rot, trans, root=root))
        data[f"only_fk_root{root}"] = to_np(robot.get_keypoints_only_fk_at_specific_root(q, root=root))
        data[f"rotation_root{root}"] = to_np(robot.get_rotation_at_specific_root(q, rot, trans, root=root))
    fk = robot.robot.link_fk_batch(q, use_names=True)
    data["all_link_names"] = np.array(list(fk.keys()))
    data["all_link_fk"] = to_np(torch.stack(list(fk.values()), dim=1))
    data["proj_tensor"] = to_np(ns.transforms.point_projection_from_3d_tensor(K, torch.from_numpy(data["keypoints"])))
    data["proj_numpy"] = ns.transforms.point_projection_from_3d(to_np(K), data["keypoints"]).astype(np.float32)
    # quaternion rotation input (urdf_robot.py:89-90)
    quat = synth.sym_uniform("fk_quat_" + robot_type, (batch, 4), 1.0, 2)
    data["keypoints_quat"] = to_np(robot.get_keypoints(q, quat, trans))
    np.savez(GOLDEN / f"fk_{robot_type}.npz", **data)
    orob = O.OracleRobot(robot_type, str(synth.URDF_PATHS[robot_type]))
    d = (orob.get_keypoints_root(q, rot, trans, root=3) - torch.from_numpy(data["keypoints_root3"])).abs().max()
    print(f"[fk {robot_type}] oracle vs reference max|diff|: {float(d):.3g}; links {list(robot.link_names)[:3]}...")


def golden_integral(ns, robot_type: str, batch: int = 2):
    dof, nkpt, ref = arch.ROBOTS[robot_type]
    data = {}
    for tag, gain in (("", 1.0), ("_peaky", HEATMAP_STRESS_GAIN)):
        hm = heatmap_logits(robot_type, batch, gain=gain)
        _, _, k, K = synth.inputs(batch, seed=13)
        root_trans = torch.zeros(batch, 3)
        root_trans[:, 2] = synth.range_uniform("root_z", (batch,), 0.8, 2.5, 13)
        layer = ns.integral.HeatmapIntegralPose(backbone="resnet50", num_joints=nkpt, depth_dim=64, height_dim=64,
                                                width_dim=64, norm_type="softmax", image_size=256.0,
                                                bbox_3d_shape=[1300, 1300, 1300], rootid=ref, fixroot=True)
        uvd, xyz = layer(hm, root_trans=root_trans, K=K)
        data["uvd" + tag], data["xyz" + tag] = to_np(uvd), to_np(xyz)
        root_uv = (uvd[:, ref, :2] + 0.5) * 256.0
        data["trans" + tag] = to_np(ns.transforms.uvz2xyz_singlepoint(root_uv, root_trans[:, 2:3], K))
        ou, ox = O.heatmap_integral(hm, nkpt, K, root_trans, ref)
        print(f"[integral {robot_type}{tag}] oracle vs reference:", float((ou - uvd).abs().max()),
              float((ox - xyz).abs().max()))
    np.savez(GOLDEN / f"integral_{robot_type}.npz", **data)


GRAD_SAMPLES = 4096


def grad_digest(g: torch.Tensor) -> dict:
    """Small, order-independent digest of a (B, nkpt*64, 64, 64) gradient tensor: the three axis marginals, the largest
    magnitude and GRAD_SAMPLES seeded entries (the full tensor is 7.3 MB per image)."""
    B = g.shape[0]
    flat = g.reshape(B, -1)
    idx = torch.from_numpy((synth.uniform01("grad_idx", GRAD_SAMPLES, 1).astype(np.float64) * flat.shape[1]).astype(np.int64))
    return {"sum_hw": to_np(g.sum(dim=(2, 3))), "sum_cw": to_np(g.sum(dim=(1, 3))), "sum_ch": to_np(g.sum(dim=(1, 2))),
            "absmax": np.float32(g.abs().max()), "samples": to_np(flat[:, idx])}


def golden_integral_backward(ns, robot_type: str, batch: int = 2):
    """f4 (first piece): autograd of the REFERENCE's HeatmapIntegralPose w.r.t. its bf16-exact logits, for the loss
    sum(uvd * G1) + sum(xyz * G2); stored as a digest.  The oracle's autograd must reproduce it."""
    dof, nkpt, ref = arch.ROBOTS[robot_type]
    _, _, k, K = synth.inputs(batch, seed=13)
    root_trans = torch.zeros(batch, 3)
    root_trans[:, 2] = synth.range_uniform("root_z", (batch,), 0.8, 2.5, 13)
    G1 = synth.sym_uniform("g_uvd", (batch, nkpt, 3), 1.0, 3)
    G2 = synth.sym_uniform("g_xyz", (batch, nkpt, 3), 1.0, 4)
    layer = ns.integral.HeatmapIntegralPose(backbone="resnet50", num_joints=nkpt, depth_dim=64, height_dim=64,
                                            width_dim=64, norm_type="softmax", image_size=256.0,
                                            bbox_3d_shape=[1300, 1300, 1300], rootid=ref, fixroot=True)
    data = {}
    for tag, gain in (("", 1.0), ("_peaky", HEATMAP_STRESS_GAIN)):
        logits = heatmap_logits(robot_type, batch, gain=gain).bfloat16().float()
        with torch.enable_grad():
            x = logits.clone().requires_grad_(True)
            uvd, xyz = layer(x, root_trans=root_trans, K=K)
            ((uvd * G1).sum() + (xyz * G2).sum()).backward()
            xo = logits.clone().requires_grad_(True)
            ou, ox = O.heatmap_integral(xo, nkpt, K, root_trans, ref, fixroot=True, image_size=256.0,
                                        depth_factor=float(layer.depth_factor))
            ((ou * G1).sum() + (ox * G2).sum()).backward()
        assert torch.equal(x.grad, xo.grad), f"oracle autograd differs from the reference ({robot_type}{tag})"
        for kk, v in grad_digest(x.grad).items():
            data[kk + tag] = v
    np.savez(GOLDEN / f"integral_backward_{robot_type}.npz", **data)


def golden_fk_backward(ns, robot_type: str, batch: int = 16):
    """f4 (second piece): autograd of the REFERENCE's URDFRobot.get_keypoints_root / get_keypoints_only_fk_at_specific_root
    and point_projection_from_3d_tensor for the loss sum(pts * G) (+ sum(uv * G2)); the gradients themselves are stored
    (a few hundred floats).  The oracle's autograd is checked against them here."""
    robot = ns.urdf_robot.URDFRobot(robot_type)
    q0, rot0, trans0 = synth.fk_inputs(robot_type, batch, seed=9)
    _, _, _, K = synth.inputs(batch, seed=5)
    nk = len(robot.link_names)
    G = synth.sym_uniform("g_fk_" + robot_type, (batch, nk, 3), 1.0, 5)
    G2 = synth.sym_uniform("g_uv_" + robot_type, (batch, nk, 2), 1.0, 6)
    orob = O.OracleRobot(robot_type, str(synth.URDF_PATHS[robot_type]))
    data = {}
    for root in sorted({0, 3, nk - 1}):
        grads = []
        for rb, proj in ((robot, ns.transforms.point_projection_from_3d_tensor), (orob, O.point_projection_from_3d)):
            with torch.enable_grad():
                q, rot, trans = (t.clone().requires_grad_(True) for t in (q0, rot0, trans0))
                pts = rb.get_keypoints_root(q, rot, trans, root=root)
                uv = proj(K, pts)
                ((pts * G).sum() + 1e-3 * (uv * G2).sum()).backward()
                grads.append((q.grad.clone(), rot.grad.clone(), trans.grad.clone()))
        for a, b_, nm in zip(grads[0], grads[1], ("q", "rot", "trans")):
            err = float((a - b_).abs().max() / a.abs().max().clamp_min(1e-12))
            assert err < 2e-5, f"oracle autograd differs from the reference ({robot_type} root {root} {nm}: {err})"
        data[f"grad_q_root{root}"], data[f"grad_rot_root{root}"], data[f"grad_trans_root{root}"] = (to_np(g) for g in grads[0])
        with torch.enable_grad():
            q = q0.clone().requires_grad_(True)
            pts = robot.get_keypoints_only_fk_at_specific_root(q, root=root) if root > 0 else robot.get_keypoints_only_fk(q)
            (pts * G).sum().backward()
        data[f"grad_q_only_fk_root{root}"] = to_np(q.grad)
    np.savez(GOLDEN / f"fk_backward_{robot_type}.npz", **data)
    print(f"[fk backward {robot_type}] wrote", {k: v.shape for k, v in data.items()})


def golden_geometry(ns):
    r6 = synth.sym_uniform("rot6d", (64, 6), 1.0, 4)
    R = ns.geometries.rot6d_to_rotmat(r6)
    quat = synth.sym_uniform("quat", (64, 4), 1.0, 4)
    np.savez(GOLDEN / "geometry.npz", rot6d_to_rotmat=to_np(R), rotmat_to_rot6d=to_np(ns.geometries.rotmat_to_rot6d(R)),
             quat_to_rotmat=to_np(ns.geometries.quat_to_rotmat(quat)))


CROP_BATCH = 16
METRIC_BATCH = 48


def golden_crop(ns):
    """f1: the reference's own dataset chain (roboutils.resize_image -> CropResizeToAspectAugmentation) per image."""
    import copy
    import dataset.augmentations as aug
    import dataset.roboutils as ru
    frames, boxes, K, k_bbox = synth.crop_inputs(CROP_BATCH)
    imgs, Ks = [], []
    for b in range(CROP_BATCH):
        state = {"camera": {"K": K[b].copy(), "resolution": (640, 480)},
                 "objects": [{"keypoints_2d": [np.array([320.0, 240.0]), np.array([10.0, 20.0])],
                              "TCO_keypoints_3d": np.array([[0.1, 0.0, 1.5], [0.0, 0.2, 1.2]])}]}
        mask = np.zeros((480, 640), dtype=np.uint8)
        rgb, mask, state = ru.resize_image(frames[b], tuple(int(v) for v in boxes[b]), mask, state)
        crop = aug.CropResizeToAspectAugmentation(resize=(256, 256))
        side = rgb.shape[0]
        rgb, _, state = crop(rgb, np.zeros((side, side), dtype=np.uint8), state)
        rgb = aug.to_torch_uint8(rgb).permute(2, 0, 1)
        imgs.append(np.asarray(torch.FloatTensor(np.asarray(rgb))).astype(np.uint8))   # dream.py:309-310
        Ks.append(torch.FloatTensor(np.asarray(state["camera"]["K"])).numpy())            # dream.py:311
        o_img, o_K = EO.crop_resize(frames[b], boxes[b], K[b])
        assert np.array_equal(o_img.numpy(), imgs[-1]), f"oracle crop {b} differs from the reference"
        assert np.array_equal(o_K.numpy(), Ks[-1]), f"oracle K {b} differs from the reference"
    # scripts/test.py:141-152 (the expression is inline in farward_loss; evaluated here verbatim on the same tensors)
    Kt = torch.as_tensor(K).float()
    bboxes = torch.as_tensor(k_bbox)
    real_bbox = torch.tensor([1000.0, 1000.0]).to(torch.float32)
    fx, fy = Kt[:, 0, 0], Kt[:, 1, 1]
    area = torch.max(torch.abs(bboxes[:, 2] - bboxes[:, 0]), torch.abs(bboxes[:, 3] - bboxes[:, 1])) ** 2
    kv = torch.tensor([torch.sqrt(fx[n] * fy[n] * real_bbox[0] * real_bbox[1] / area[n]) for n in range(CROP_BATCH)]).to(torch.float32)
    assert torch.equal(kv, EO.k_value(fx, fy, bboxes))
    np.savez_compressed(GOLDEN / "crop.npz", images=np.stack(imgs), K=np.stack(Ks), k_value=kv.numpy())


def golden_metrics(ns, robot_type: str):
    """f2: the reference's compute_metrics_batch / summary_add_pck on seeded predictions."""
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):   # plotting imports of metrics.py:4-5 (unused here)
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import utils.metrics as M
    robot = ns.urdf_robot.URDFRobot(robot_type)
    ref_id = arch.ROBOTS[robot_type][2]
    q, rot, trans, gt_q, gt3, gt2, K = synth.metric_inputs(robot_type, METRIC_BATCH)
    out = M.compute_metrics_batch(robot=robot, gt_keypoints3d=gt3, gt_keypoints2d=gt2, K_original=K, gt_joint=gt_q,
                                  pred_joint=q, pred_rot=rot, pred_trans=trans, pred_depth=None, pred_xy=None,
                                  pred_xyz_integral=None, reference_keypoint_id=ref_id)
    names = ["error3d", "error2d", "dis3d", "dis2d", "l1_jointerror", "mean_jointerror", "error_depth",
             "batch_error_relative", "error3d_relative"]
    res = {n: np.asarray(v, dtype=np.float32) for n, v in zip(names, out)}
    if ref_id == 0:
        kp = robot.get_keypoints(q, rot, trans)
    else:
        kp = robot.get_keypoints_root(q, rot, trans, root=ref_id)
    mine = EO.metrics_batch(to_np(kp), to_np(gt3), to_np(gt2), to_np(K), to_np(q), to_np(gt_q), ref_id, robot_type == "panda")
    for n in names:
        np.testing.assert_array_equal(np.asarray(mine[n], dtype=np.float32), res[n], err_msg=n)
    # summary over per-image errors spread across the AUC range (scaled copies of the batch errors)
    d3 = np.concatenate([res["error3d"] * s for s in (0.02, 0.05, 0.1, 0.2)]).astype(np.float32)
    d2 = np.concatenate([res["error2d"] * s for s in (0.01, 0.02, 0.04, 0.08)]).astype(np.float32)
    d2 = np.nan_to_num(d2, nan=25.0, posinf=25.0)
    summ = M.summary_add_pck({"dis3d": list(d3), "dis2d": list(d2)})
    mine_s = EO.summary_add_pck(d3, d2)
    for k, v in summ.items():
        assert float(mine_s[k]) == float(v), (k, mine_s[k], v)
    np.savez(GOLDEN / f"metrics_{robot_type}.npz", pred_kp3d=to_np(kp), sum_dis3d=d3, sum_dis2d=d2,
             sum_keys=np.array(list(summ.keys())), sum_vals=np.array([float(v) for v in summ.values()], dtype=np.float64),
             **res)


PNP_BATCH = 32


def pnp_case(robot, robot_type: str, batch: int = PNP_BATCH):
    """pts2d / pts3d of one PnP fixture from seeds: FK keypoints in the base frame (world_3d_pts of test.py:121), moved
    by a seeded camera pose, projected with K, plus pixel noise.  `robot` needs get_keypoints_only_fk."""
    q, rvec, t, K, noise = synth.pnp_inputs(robot_type, batch)
    pts3d = robot.get_keypoints_only_fk(q).float().cpu()
    th = rvec.norm(dim=1, keepdim=True)
    k = rvec / th
    Kx = torch.zeros(batch, 3, 3)
    Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0], Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    R = torch.eye(3)[None] + torch.sin(th)[:, :, None] * Kx + (1 - torch.cos(th))[:, :, None] * (Kx @ Kx)
    cam = pts3d @ R.transpose(1, 2) + t[:, None, :]
    uvw = cam @ K.T
    pts2d = uvw[:, :, :2] / uvw[:, :, 2:3] + noise
    return pts2d.contiguous(), pts3d.contiguous(), K, rvec, t


def golden_pnp(ns, robot_type: str):
    """f3: the reference's BPnP_m3d.apply + angle_axis_to_rotation_matrix + rotmat_to_rot6d (scripts/test.py:121-124)."""
    import types
    if "utils.BPnP" not in sys.modules:
        sys.modules.setdefault("kornia", types.ModuleType("kornia"))       # BPnP.py:6 (used by the backward only)
        real_tensor = torch.tensor
        torch.tensor = lambda *a, **k: real_tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"})  # BPnP.py:2
        try:
            import utils.BPnP  # noqa: F401
        finally:
            torch.tensor = real_tensor
    BP = sys.modules["utils.BPnP"]
    robot = ns.urdf_robot.URDFRobot(robot_type)
    pts2d, pts3d, K, rvec, t = pnp_case(robot, robot_type)
    out = BP.BPnP_m3d.apply(pts2d, pts3d, K)
    rot = ns.geometries.rotmat_to_rot6d(ns.geometries.angle_axis_to_rotation_matrix(out[:, 0:3])[:, :3, :3])
    mine = EO.pnp_m3d(pts2d, pts3d, K)
    assert torch.equal(mine, out), "oracle PnP differs from the reference"
    assert torch.equal(EO.angle_axis_to_rot6d(out[:, :3]), rot), "oracle rot6d differs from the reference"
    np.savez(GOLDEN / f"pnp_{robot_type}.npz", pose6=to_np(out), rot6d=to_np(rot), pts2d=to_np(pts2d), pts3d=to_np(pts3d))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calibrate", action="store_true")
    ap.add_argument("--only-eval", action="store_true", help="only the crop / metrics fixtures (rows f1, f2)")
    ap.add_argument("--only-fk-backward", action="store_true", help="only the FK / projection autograd fixtures (row f4)")
    args = ap.parse_args()
    if args.calibrate:
        calibrate()
        return
    ns = ref_harness.setup({k: str(v) for k, v in synth.URDF_PATHS.items()})
    if args.only_fk_backward:
        for r in ("panda", "kuka", "baxter"):
            golden_fk_backward(ns, r)
        return
    golden_crop(ns)
    for r in ("panda", "kuka", "baxter"):
        golden_metrics(ns, r)
        golden_pnp(ns, r)
    for r in ("panda", "baxter"):
        golden_integral_backward(ns, r)
    for r in ("panda", "kuka", "baxter"):
        golden_fk_backward(ns, r)
    if args.only_eval:
        return
    golden_geometry(ns)
    for r in ("panda", "kuka", "baxter"):
        golden_fk(ns, r)
        golden_integral(ns, r)
    golden_depthnet(ns)
    for r in ("panda", "kuka", "baxter"):
        golden_full(ns, r)


if __name__ == "__main__":
    main()
