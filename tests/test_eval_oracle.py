"""CPU: the f1 / f2 oracle (oracle/eval_oracle.py) against the fixtures generated from the REAL reference
(tests/golden/make_golden.py: golden_crop, golden_metrics), plus the shims' argument checks (no GPU)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import horopose_b200  # noqa: E402,F401
from horopose_b200 import arch, synth  # noqa: E402
from oracle import eval_oracle as EO  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
CROP_BATCH, METRIC_BATCH = 16, 48


def test_crop_oracle_matches_reference_bytes():
    g = np.load(GOLDEN / "crop.npz")
    frames, boxes, K, k_bbox = synth.crop_inputs(CROP_BATCH)
    assert any(int(max(b[2] - b[0], b[3] - b[1])) == 256 for b in boxes)  # identity branch is covered
    for b in range(CROP_BATCH):
        img, Kn = EO.crop_resize(frames[b], boxes[b], K[b])
        # byte work: the fixture holds the reference's bytes.  The oracle calls torch's CPU bilinear kernel, whose
        # AVX2 / AVX-512 builds may contract multiply-adds differently; a different host may therefore flip the
        # truncation of a value that sits within one fp32 ulp of an integer: allow single-LSB flips on < 0.01 % of the
        # bytes here (0 observed in the generating container).  The CUDA kernel is compared EXACTLY to the fixture.
        diff = np.abs(img.numpy().astype(np.int16) - g["images"][b].astype(np.int16))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-4, f"crop {b}: max diff {diff.max()}, {(diff != 0).sum()} bytes"
        assert np.array_equal(Kn.numpy(), g["K"][b]), f"K {b}"
    Kt = torch.as_tensor(K).float()
    kv = EO.k_value(Kt[:, 0, 0], Kt[:, 1, 1], k_bbox)
    assert np.array_equal(kv.numpy(), g["k_value"])


def test_crop_properties():
    # a box that already has the target size passes the pixels through; K only moves the principal point
    frames, _, K, _ = synth.crop_inputs(1)
    img, Kn = EO.crop_resize(frames[0], (100, 60, 356, 316), K[0])
    assert np.array_equal(img.permute(1, 2, 0).numpy(), frames[0][60:316, 100:356])
    assert Kn[0, 0] == np.float32(K[0, 0, 0]) and Kn[0, 2] == np.float32(K[0, 0, 2] - 100)
    # a 2x smaller square box doubles the focal length; the principal point follows get_K_crop_resize's convention
    # cx' = s * (cx - wmin) - 0.5 (geometries.py:377-395 composed)
    img, Kn = EO.crop_resize(frames[0], (100, 60, 228, 188), K[0])
    assert abs(float(Kn[0, 0]) - 2 * K[0, 0, 0]) < 1e-3
    assert abs(float(Kn[0, 2]) - (2 * (K[0, 0, 2] - 100) - 0.5)) < 1e-3
    assert abs(float(Kn[1, 2]) - (2 * (K[0, 1, 2] - 60) - 0.5)) < 1e-3


@pytest.mark.parametrize("rt", ["panda", "kuka", "baxter"])
def test_metrics_oracle_matches_reference(rt):
    g = np.load(GOLDEN / f"metrics_{rt}.npz")
    q, rot, trans, gt_q, gt3, gt2, K = synth.metric_inputs(rt, METRIC_BATCH)
    ref_id = arch.ROBOTS[rt][2]
    res = EO.metrics_batch(g["pred_kp3d"], gt3.numpy(), gt2.numpy(), K.numpy(), q.numpy(), gt_q.numpy(), ref_id,
                           rt == "panda")
    for name, v in res.items():
        # same numpy code path as the generating container; the projection matmul may use a different BLAS kernel
        np.testing.assert_allclose(np.asarray(v, dtype=np.float32), g[name], rtol=2e-6, atol=1e-7, err_msg=name,
                                   equal_nan=True)
    summ = EO.summary_add_pck(g["sum_dis3d"], g["sum_dis2d"])
    for k, v in zip(g["sum_keys"], g["sum_vals"]):
        assert float(summ[str(k)]) == pytest.approx(float(v), rel=1e-12, abs=0), k


def test_summary_threshold_grid():
    # the device kernel bins with thresholds i * delta: identical to np.arange's values and count
    for limit, delta in ((0.1, 0.00001), (20.0, 0.01)):
        thr = np.arange(0.0, limit, delta)
        assert len(thr) == int(np.ceil(limit / delta))
        assert np.array_equal(thr, np.arange(len(thr)) * delta)


def test_shims_refuse_cpu_tensors():
    from horopose_b200 import metrics, preprocess
    from horopose_b200._lib import HrpError
    frames, boxes, K, _ = synth.crop_inputs(1)
    with pytest.raises(HrpError):
        preprocess.crop_resize_batch(torch.from_numpy(frames), torch.from_numpy(boxes), torch.from_numpy(K))
    with pytest.raises(HrpError):
        metrics.summary_add_pck({"dis3d": torch.zeros(4), "dis2d": torch.zeros(4)})


@pytest.mark.parametrize("rt", ["panda", "kuka", "baxter"])
def test_pnp_oracle_matches_reference(rt):
    """f3: cv2.solvePnP (EPnP -> iterative) as the reference's BPnP_m3d.forward calls it, vs the fixture written by the
    reference itself; plus the property that the noisy correspondences recover the seeded pose to ~1e-2."""
    cv2 = pytest.importorskip("cv2")  # noqa: F841  (the reference's own dependency; present in this image)
    g = np.load(GOLDEN / f"pnp_{rt}.npz")
    K = torch.tensor([[615.0, 0.0, 320.0], [0.0, 615.5, 240.0], [0.0, 0.0, 1.0]])
    out = EO.pnp_m3d(torch.from_numpy(g["pts2d"]), torch.from_numpy(g["pts3d"]), K)
    np.testing.assert_allclose(out.numpy(), g["pose6"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(EO.angle_axis_to_rot6d(out[:, :3]).numpy(), g["rot6d"], rtol=0, atol=2e-6)
    q, rvec, t, _, _ = synth.pnp_inputs(rt, 32)
    assert np.abs(g["pose6"][:, 3:] - t.numpy()).max() < 0.1       # 0.5 px noise on 7-17 points at 1-2.5 m
