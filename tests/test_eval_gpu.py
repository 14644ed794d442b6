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
    for b in range(40):
        o_img, o_K = EO.crop_resize(frames[b], boxes[b], K[b])
        assert np.array_equal(img[b].cpu().numpy(), o_img.numpy()), (b, boxes[b])
        assert np.array_equal(Kn[b].cpu().numpy(), o_K.numpy()), (b, boxes[b])
    assert np.array_equal(kv.cpu().numpy(), EO.k_value(Kn[:, 0, 0].cpu(), Kn[:, 1, 1].cpu(), k_bbox).numpy())
    # odd frame size
    f2 = np.ascontiguousarray(frames[:3, :333, :517])
    b2 = np.array([[0, 0, 517, 333], [13, 17, 269, 273], [500, 300, 517, 333]], dtype=np.int32)
    img2, K2 = crop_resize_batch(torch.from_numpy(f2).cuda(), torch.from_numpy(b2).cuda(), torch.from_numpy(K[:3]).cuda())
    for b in range(3):
        o_img, o_K = EO.crop_resize(f2[b], b2[b], K[b])
        assert np.array_equal(img2[b].cpu().numpy(), o_img.numpy()), b
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
