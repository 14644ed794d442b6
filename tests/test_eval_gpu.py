"""CUDA input pipeline (row f1) and on-device metrics (row f2) through the C ABI, against the fixtures generated
from the real reference and against the CPU oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pytestmark = pytest.mark.gpu
GOLDEN = ROOT / "tests" / "golden"
CROP_BATCH, METRIC_BATCH = 16, 48


def test_crop_resize_bit_exact_vs_reference_golden(hrp_lib):
    """Byte work: the crops must equal the reference's bytes; K and k_value its fp32 values (bit-exact)."""
    from horopose_b200 import synth
    from horopose_b200.preprocess import crop_resize_batch
    g = np.load(GOLDEN / "crop.npz")
    frames, boxes, K, k_bbox = synth.crop_inputs(CROP_BATCH)
    img, Kn, kv = crop_resize_batch(torch.from_numpy(frames).cuda(), torch.from_numpy(boxes).cuda(),
                                    torch.from_numpy(K).cuda(), k_bbox=torch.from_numpy(k_bbox).cuda())
    img, Kn, kv = img.cpu().numpy(), Kn.cpu().numpy(), kv.cpu().numpy()
    for b in range(CROP_BATCH):
        diff = np.argwhere(img[b] != g["images"][b])
        assert diff.shape[0] == 0, (b, boxes[b], diff[:5], img[b][tuple(diff[0])], g["images"][b][tuple(diff[0])])
    assert np.array_equal(Kn, g["K"])
    assert np.array_equal(kv, g["k_value"])


def test_crop_resize_vs_oracle_random_boxes_and_ragged_frames(hrp_lib):
    """More boxes than the fixture holds (incl. 1-pixel-wide strips and a non-640x480 frame) against the live oracle."""
    from horopose_b200 import synth
    from horopose_b200.preprocess import crop_resize_batch
    from oracle import eval_oracle as EO
    frames, boxes, K, k_bbox = synth.crop_inputs(40, seed=9)
    boxes[10] = (7, 9, 8, 300)      # one pixel wide
    boxes[11] = (0, 470, 640, 471)  # one pixel tall
    boxes[12] = (100, 100, 103, 103)
    img, Kn, kv = crop_resize_batch(torch.from_numpy(frames).cuda(), torch.from_numpy(boxes).cuda(),
                                    torch.from_numpy(K).cuda(), k_bbox=torch.from_numpy(k_bbox).cuda(), k_from_crop_K=True)
    def same_bytes(a, o):
        # the live oracle is torch's CPU bilinear on THIS host: a different vector ISA build may flip the truncation of a
        # value within one fp32 ulp of an integer (see tests/test_eval_oracle.py); the committed fixture is compared exactly
        d = np.abs(a.astype(np.int16) - o.astype(np.int16))
        return d.max() <= 1 and (d != 0).mean() < 1e-4

    for b in range(40):
        o_img, o_K = EO.crop_resize(frames[b], boxes[b], K[b])
        assert same_bytes(img[b].cpu().numpy(), o_img.numpy()), (b, boxes[b])
        assert np.array_equal(Kn[b].cpu().numpy(), o_K.numpy()), (b, boxes[b])
    assert np.array_equal(kv.cpu().numpy(), EO.k_value(Kn[:, 0, 0].cpu(), Kn[:, 1, 1].cpu(), k_bbox).numpy())
    # odd frame size
    f2 = np.ascontiguousarray(frames[:3, :333, :517])
    b2 = np.array([[0, 0, 517, 333], [13, 17, 269, 273], [500, 300, 517, 333]], dtype=np.int32)
    img2, K2 = crop_resize_batch(torch.from_numpy(f2).cuda(), torch.from_numpy(b2).cuda(), torch.from_numpy(K[:3]).cuda())
    for b in range(3):
        o_img, o_K = EO.crop_resize(f2[b], b2[b], K[b])
        assert same_bytes(img2[b].cpu().numpy(), o_img.numpy()), b
        assert np.array_equal(K2[b].cpu().numpy(), o_K.numpy()), b
    with pytest.raises(ValueError):
        crop_resize_batch(torch.from_numpy(f2).cuda(), torch.tensor([[0, 0, 518, 10]] * 3).cuda(), torch.from_numpy(K[:3]).cuda())


def test_crop_feeds_the_model(hrp_lib):
    """frame -> crop kernel -> uint8 forward equals the forward on the oracle's crop bytes (drop-in for the dataset)."""
    from horopose_b200 import synth
    from horopose_b200.models import get_rootNetwithRegInt_model
    from horopose_b200.preprocess import crop_resize_batch
    from oracle import eval_oracle as EO
    frames, boxes, K, k_bbox = synth.crop_inputs(4)
    img, Kn, kv = crop_resize_batch(torch.from_numpy(frames).cuda(), torch.from_numpy(boxes).cuda(),
                                    torch.from_numpy(K).cuda(), k_bbox=torch.from_numpy(k_bbox).cuda())
    args = dict(backbone_name="resnet50", rootnet_backbone_name="hrnet32", n_iter=4, other_image_size=256.0,
                bbox_3d_shape=[1300, 1300, 1300], reference_keypoint_id=3, fix_root=True, rotation_dim=6)
    model = get_rootNetwithRegInt_model({"robot_type": "panda", "pose_params": None, "cam_params": np.eye(4),
                                         "init_pose_from_mean": True}, args)
    model.load_state_dict(synth.full_state_dict("panda"), strict=True)
    out_a = model(img, img, kv, Kn)
    o_img = torch.stack([EO.crop_resize(frames[b], boxes[b], K[b])[0] for b in range(4)]).cuda()
    out_b = model(o_img, o_img, kv, Kn)
    for a, b in zip(out_a, out_b):
        assert torch.equal(a, b)


@pytest.mark.parametrize("rt", ["panda", "kuka", "baxter"])
def test_metrics_batch_vs_reference_golden(rt, hrp_lib):
    """compute_metrics_batch: fp32 within 1e-5 relative of the reference's numpy results (FK on the GPU included)."""
    from horopose_b200 import arch, synth
    from horopose_b200.metrics import compute_metrics_batch
    from horopose_b200.robot import URDFRobot
    g = np.load(GOLDEN / f"metrics_{rt}.npz")
    robot = URDFRobot(rt)
    q, rot, trans, gt_q, gt3, gt2, K = (t.cuda() for t in synth.metric_inputs(rt, METRIC_BATCH))
    out = compute_metrics_batch(robot, gt3, gt2, K, gt_q, pred_joint=q, pred_rot=rot, pred_trans=trans, pred_depth=None,
                                pred_xy=None, pred_xyz_integral=None, reference_keypoint_id=arch.ROBOTS[rt][2])
    names = ["error3d", "error2d", "dis3d", "dis2d", "l1_jointerror", "mean_jointerror", "error_depth",
             "batch_error_relative", "error3d_relative"]
    TOL = 1e-5
    for n, v in zip(names, out):
        ref = g[n]
        got = v.cpu().numpy()
        assert got.shape == ref.shape, n
        assert np.array_equal(np.isnan(got), np.isnan(ref)), n
        m = ~np.isnan(ref)
        # error_depth is a difference of two ~1.5 m depths: its tolerance is relative to the depths, not to itself
        scale = 2.5 if n in ("error_depth", "batch_error_relative") else max(np.abs(ref[m]).max(), 1e-12)
        assert np.abs(got[m] - ref[m]).max() / scale < TOL, (n, np.abs(got[m] - ref[m]).max(), scale)
    # xyz-integral-only variant (metrics.py:23-26): no joints
    out2 = compute_metrics_batch(robot, gt3, gt2, K, gt_q, pred_joint=None, pred_rot=None, pred_trans=None,
                                 pred_xyz_integral=torch.from_numpy(g["pred_kp3d"]).cuda(),
                                 reference_keypoint_id=arch.ROBOTS[rt][2])
    assert np.abs(out2[0].cpu().numpy() - g["error3d"]).max() < TOL * np.abs(g["error3d"]).max()
    assert float(out2[4].abs().max()) == 0.0 and float(out2[5].abs().max()) == 0.0


@pytest.mark.parametrize("rt", ["panda", "baxter"])
def test_metrics_summary_vs_reference_golden(rt, hrp_lib):
    """summary_add_pck: counts are integers (threshold fractions and the AUC curve are exact up to fp64 summation
    order), medians exact, means within fp32 round-off."""
    from horopose_b200.metrics import MetricAccumulator, summary_add_pck
    g = np.load(GOLDEN / f"metrics_{rt}.npz")
    d3, d2 = torch.from_numpy(g["sum_dis3d"]).cuda(), torch.from_numpy(g["sum_dis2d"]).cuda()
    s = summary_add_pck({"dis3d": d3, "dis2d": d2})
    ref = dict(zip((str(k) for k in g["sum_keys"]), g["sum_vals"]))
    assert set(s) == set(ref)
    for k, v in ref.items():
        if k.endswith("median"):
            assert s[k] == v, k
        elif k.endswith("mean"):
            assert s[k] == pytest.approx(v, rel=1e-6), k
        else:
            assert s[k] == pytest.approx(v, rel=1e-12, abs=1e-15), k
    # odd n (single middle element), chunked accumulation and the oracle on fresh data
    from oracle import eval_oracle as EO
    acc = MetricAccumulator()
    e3 = torch.rand(1001, generator=torch.Generator().manual_seed(3)) * 0.12
    e2 = torch.rand(1001, generator=torch.Generator().manual_seed(4)) * 25.0
    for i in range(0, 1001, 77):
        acc.add((e3[i:i + 77].cuda(), e2[i:i + 77].cuda(), None, None, None, None, None, None, e3[i:i + 77].cuda()))
    s2, o2 = acc.summary(), EO.summary_add_pck(e3.numpy(), e2.numpy())
    for k, v in o2.items():
        assert s2[k] == pytest.approx(float(v), rel=1e-6 if k.endswith("mean") else 1e-12, abs=1e-15), k
    assert s2["ADD/median"] == float(o2["ADD/median"]) and s2["ADD_2D/median"] == float(o2["ADD_2D/median"])


def test_metrics_summary_nan_and_host_lists(hrp_lib):
    """An image without in-frame ground-truth keypoints has error2d = 0/0 (metrics.py:67-70): np.mean / np.median then
    give NaN, the threshold counts treat it as 'above' -- the kernel must do the same.  And `alldis` may be the list of
    python floats that scripts/test.py:199-232 builds."""
    import math
    from horopose_b200.metrics import summary_add_pck
    from oracle import eval_oracle as EO
    g = torch.Generator().manual_seed(9)
    e3 = (torch.rand(400, generator=g) * 0.12).numpy()
    e2 = (torch.rand(400, generator=g) * 25.0).numpy()
    e2[17] = float("nan")
    e2[203] = float("nan")
    got = summary_add_pck({"dis3d": [float(v) for v in e3], "dis2d": [float(v) for v in e2]})   # host python floats
    want = EO.summary_add_pck(e3, e2)
    for k, v in want.items():
        v = float(v)
        if math.isnan(v):
            assert math.isnan(got[k]), (k, got[k])
        else:
            assert got[k] == pytest.approx(v, rel=1e-6 if k.endswith("mean") else 1e-12, abs=1e-15), k
    assert math.isnan(got["ADD_2D/median"]) and math.isnan(got["ADD_2D/mean"]) and not math.isnan(got["ADD/median"])
    # mixed list of CUDA tensors and numpy chunks
    got2 = summary_add_pck({"dis3d": [torch.from_numpy(e3[:100]).cuda(), e3[100:]], "dis2d": [e2[:100], torch.from_numpy(e2[100:]).cuda()]})
    assert got2["ADD/median"] == got["ADD/median"] and got2["PCK/AUC"] == got["PCK/AUC"]


def test_eval_loop_frames_to_summary_vs_oracle_chain(hrp_lib):
    """System parity of the widened path: frames + boxes -> crop -> network -> metrics -> summary on the GPU against the
    same chain built from the CPU oracles (crop oracle -> fp32 reference forward -> metrics oracle).  The crop is exact.
    The network is bf16 against fp32 and these inputs are image-like (constant padding bars, smooth gradients), where
    bf16 rounding errors are spatially correlated and do not average out in the pooled features the way they do on the
    config's U[0,1) noise inputs: PyTorch's own autocast-bf16 forward of the reference network is 2-15 mm / 0.4-3 px off
    the fp32 forward here (vs 0.2-0.5 mm / < 0.1 px on noise) -- weights and activations rounded to bf16 contribute
    equally (profiles/r02_exp_precision_sources.txt), so no storage trick short of a wider type removes it.  The
    north-star bars (stated on the config's synthetic inputs) are NOT met on these inputs by any bf16 evaluation; what is
    asserted over 24 frames is that this path is statistically no worse than PyTorch's own bf16: its RMS error against
    the fp32 chain must not exceed 1.25 x the RMS error of torch-autocast-bf16 (the maxima are logged beside them), and
    the joint angles (which do not depend on the root depth) meet their 1e-2 rad bar outright."""
    from horopose_b200 import arch, synth
    from horopose_b200.models import get_rootNetwithRegInt_model
    from horopose_b200.pipeline import EvalPipeline
    from horopose_b200.robot import URDFRobot
    from oracle import eval_oracle as EO
    from oracle import horopose_oracle as O
    rt, B = "panda", 24
    ref_id = arch.ROBOTS[rt][2]
    frames, boxes, Kf, k_bbox = synth.crop_inputs(B, seed=5)
    # k = f * 1000 / side (scripts/test.py:141-152): boxes of 400..640 px keep the synthetic depth (gamma ~ 2) * k / 1000
    # at 2-3 m (a 60-pixel box would put the robot 20 m away)
    k_bbox = np.stack([np.array([0.0, 0.0, 400.0 + 40.0 * (b % 6), 400.0 + 40.0 * (b % 6)], dtype=np.float32) for b in range(B)])
    q, rot, trans, gt_q, gt3, gt2, _ = synth.metric_inputs(rt, B)
    args = dict(backbone_name="resnet50", rootnet_backbone_name="hrnet32", n_iter=4, other_image_size=256.0,
                bbox_3d_shape=[1300, 1300, 1300], reference_keypoint_id=ref_id, fix_root=True, rotation_dim=6)
    model = get_rootNetwithRegInt_model({"robot_type": rt, "pose_params": None, "cam_params": np.eye(4),
                                         "init_pose_from_mean": True}, args)
    sd = synth.full_state_dict(rt)
    model.load_state_dict(sd, strict=True)
    pipe = EvalPipeline(model, URDFRobot(rt), ref_id)
    for lo, hi in ((0, 4), (4, 6), (6, 24)):  # ragged batches through the accumulator
        pipe.step(torch.from_numpy(frames[lo:hi]), torch.from_numpy(boxes[lo:hi]), torch.from_numpy(Kf[lo:hi]),
                  torch.from_numpy(k_bbox[lo:hi]), gt3[lo:hi], gt2[lo:hi], gt_q[lo:hi])
    got = pipe.summary()
    e3 = torch.cat(pipe.acc.dis3d).cpu().numpy()
    e2 = torch.cat(pipe.acc.dis2d).cpu().numpy()
    ej = torch.cat(pipe.joint_err).cpu().numpy()
    # oracle chain (fp32) and the torch-autocast-bf16 evaluation of the same chain
    crops = [EO.crop_resize(frames[b], boxes[b], Kf[b]) for b in range(B)]
    x = torch.stack([c[0] for c in crops]).float() / 255.0
    Kc = torch.stack([c[1] for c in crops])
    Kf32 = torch.as_tensor(Kf).float()
    kv = EO.k_value(Kf32[:, 0, 0], Kf32[:, 1, 1], k_bbox)
    orob = O.OracleRobot(rt, str(synth.URDF_PATHS[rt]))

    def chain_metrics(outs):
        kp = orob.get_keypoints_root(outs[0], outs[1], outs[2], root=ref_id)
        return EO.metrics_batch(kp.numpy(), gt3.numpy(), gt2.numpy(), Kf32.numpy(), outs[0].numpy(), gt_q.numpy(), ref_id, True)

    with torch.no_grad():
        m = chain_metrics(O.full_forward(sd, orob, x, x, kv, Kc))
        sd_cuda = {k_: v.cuda() for k_, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            feat, x_out, heat = O.full_features(sd_cuda, x.cuda(), x.cuda())
        mf = chain_metrics(O.full_head(sd, orob, feat.float().cpu(), x_out.float().cpu(), heat.float().cpu(), kv, Kc))
    want = EO.summary_add_pck(m["error3d"], m["error2d"])
    ok = ~np.isnan(m["error2d"])
    checks = [("error3d [m]", e3, m["error3d"], mf["error3d"], 1e-3),
              ("error2d [px]", e2[ok], m["error2d"][ok], mf["error2d"][ok], 0.5),
              ("mean_jointerror [rad]", ej, m["mean_jointerror"], mf["mean_jointerror"], 1e-2)]
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    with open(out_dir / "model_parity.txt", "a") as f:
        f.write(f"[eval loop] frames -> crop -> network -> metrics, image-like crops, panda B={B}\n")
        for name, a, r, fl, bar in checks:
            err, floor = float(np.abs(a - r).max()), float(np.abs(fl - r).max())
            rms, rms_floor = float(np.sqrt(np.mean((a - r) ** 2))), float(np.sqrt(np.mean((fl - r) ** 2)))
            f.write(f"  {name:22s} vs fp32 oracle chain: max {err:.3e} rms {rms:.3e}  (north-star bar {bar:.1e}; "
                    f"torch-autocast-bf16 floor: max {floor:.3e} rms {rms_floor:.3e})\n")
    for name, a, r, fl, bar in checks:
        rms, rms_floor = float(np.sqrt(np.mean((a - r) ** 2))), float(np.sqrt(np.mean((fl - r) ** 2)))
        if name.startswith("mean_jointerror"):
            assert float(np.abs(a - r).max()) < bar, (name, float(np.abs(a - r).max()), bar)
        else:
            assert rms <= 1.25 * rms_floor + 1e-9, (name, rms, rms_floor)
    tol3 = max(1e-3, 1.5 * float(np.abs(mf["error3d"] - m["error3d"]).max()))
    assert abs(got["ADD/mean"] - float(want["ADD/mean"])) < tol3
    assert abs(got["Depth_l1_error/mean_m"] - float(m["error_depth"].mean())) < tol3
    assert set(want) <= set(got)


def test_uint8_forward_equals_float_forward(hrp_lib):
    """`model(u8)` must equal `model(u8.float() / 255)` bit for bit: the fused `/255` 4-pixel packing kernel against
    the fp32 packing kernel (same fp32 quotient, same bf16 rounding)."""
    from horopose_b200 import synth
    from horopose_b200.models import get_rootNetwithRegInt_model
    x_reg, x_root, k, K = synth.inputs(3, seed=17)
    u8 = lambda t: (t * 255.0).round().clamp(0, 255).to(torch.uint8).cuda()
    a, b = u8(x_reg), u8(x_root)
    args = dict(backbone_name="resnet50", rootnet_backbone_name="hrnet32", n_iter=4, other_image_size=256.0,
                bbox_3d_shape=[1300, 1300, 1300], reference_keypoint_id=3, fix_root=True, rotation_dim=6)
    model = get_rootNetwithRegInt_model({"robot_type": "panda", "pose_params": None, "cam_params": np.eye(4),
                                         "init_pose_from_mean": True}, args)
    model.load_state_dict(synth.full_state_dict("panda"), strict=True)
    o8 = model(a, b, k.cuda(), K.cuda())
    of = model(a.float() / 255.0, b.float() / 255.0, k.cuda(), K.cuda())
    for x, y in zip(o8, of):
        assert torch.equal(x, y)


@pytest.mark.parametrize("rt", ["panda", "kuka", "baxter"])
def test_pnp_vs_reference_golden(rt, hrp_lib):
    """f3: the GPU solver (DLT -> Gauss-Newton, fp64) against the reference's BPnP_m3d (cv2 EPnP -> LM): both reach the
    same minimiser of the reprojection error; fp32 outputs within 1e-5 (rad / m), rot6d within 1e-5."""
    from horopose_b200.pnp import BPnP_m3d, pnp_rot6d
    g = np.load(GOLDEN / f"pnp_{rt}.npz")
    K = torch.tensor([[615.0, 0.0, 320.0], [0.0, 615.5, 240.0], [0.0, 0.0, 1.0]]).cuda()
    p2, p3 = torch.from_numpy(g["pts2d"]).cuda(), torch.from_numpy(g["pts3d"]).cuda()
    out = BPnP_m3d.apply(p2, p3, K).cpu().numpy()
    TOL = 1e-5
    assert np.abs(out - g["pose6"]).max() < TOL, np.abs(out - g["pose6"]).max()
    r6 = pnp_rot6d(p2, p3, K[None].repeat(p2.shape[0], 1, 1)).cpu().numpy()   # batched-K variant
    assert np.abs(r6 - g["rot6d"]).max() < TOL, np.abs(r6 - g["rot6d"]).max()


def test_pnp_recovers_exact_pose_and_rejects_bad_input(hrp_lib):
    """Noise-free projections: the seeded pose comes back to fp32 round-off, including rotations close to pi."""
    from horopose_b200 import synth
    from horopose_b200.pnp import BPnP_m3d
    from horopose_b200.robot import URDFRobot
    B = 256
    q, rvec, t, K, _ = synth.pnp_inputs("baxter", B, seed=21)
    rvec[:8] = rvec[:8] / rvec[:8].norm(dim=1, keepdim=True) * 3.1405   # near pi
    p3 = URDFRobot("baxter").get_keypoints_only_fk(q.cuda()).double().cpu()
    th = rvec.double().norm(dim=1, keepdim=True)
    k = rvec.double() / th
    Kx = torch.zeros(B, 3, 3, dtype=torch.float64)
    Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0], Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    R = torch.eye(3, dtype=torch.float64)[None] + torch.sin(th)[:, :, None] * Kx + (1 - torch.cos(th))[:, :, None] * (Kx @ Kx)
    cam = p3 @ R.transpose(1, 2) + t.double()[:, None, :]
    uvw = cam @ K.double().T
    p2 = uvw[:, :, :2] / uvw[:, :, 2:3]
    out = BPnP_m3d.apply(p2.float().cuda(), p3.float().cuda(), K.cuda()).cpu().double()
    # compare as rotation matrices (angle-axis is double-valued at pi) and translations
    th2 = out[:, :3].norm(dim=1, keepdim=True)
    k2 = out[:, :3] / th2
    Kx2 = torch.zeros(B, 3, 3, dtype=torch.float64)
    Kx2[:, 0, 1], Kx2[:, 0, 2], Kx2[:, 1, 0], Kx2[:, 1, 2], Kx2[:, 2, 0], Kx2[:, 2, 1] = -k2[:, 2], k2[:, 1], k2[:, 2], -k2[:, 0], -k2[:, 1], k2[:, 0]
    R2 = torch.eye(3, dtype=torch.float64)[None] + torch.sin(th2)[:, :, None] * Kx2 + (1 - torch.cos(th2))[:, :, None] * (Kx2 @ Kx2)
    # inputs were rounded to fp32 (pixels ~1e-5 px, points ~1e-7 m): the recovered pose moves by ~1e-5 at most
    assert float((R2 - R).abs().max()) < 5e-5, float((R2 - R).abs().max())
    assert float((out[:, 3:] - t.double()).abs().max()) < 5e-5, float((out[:, 3:] - t.double()).abs().max())
    with pytest.raises(Exception):
        BPnP_m3d.apply(p2[:, :5].float().cuda(), p3[:, :5].float().cuda(), K.cuda())   # < 6 points
    with pytest.raises(NotImplementedError):
        BPnP_m3d.apply(p2.float().cuda().requires_grad_(), p3.float().cuda(), K.cuda())
