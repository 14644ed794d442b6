"""CUDA head path (FK, projection, fused soft-argmax head) vs the committed reference outputs and the CPU oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
pytestmark = pytest.mark.gpu
GOLDEN = ROOT / "tests" / "golden"
ROBOT_TYPES = ["panda", "kuka", "baxter"]


def _rel(a, b):
    """max |a-b| / max |b| -- the 'relative' of the north-star FK / projection bar (1e-5)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


@pytest.mark.parametrize("rt", ROBOT_TYPES)
def test_fk_and_projection_vs_reference_golden(rt, hrp_lib):
    """FK + projection in fp32 within 1e-5 relative of the reference's own outputs (BASELINE.json north_star)."""
    from horopose_b200 import synth
    from horopose_b200.robot import URDFRobot, point_projection_from_3d, point_projection_from_3d_tensor
    g = np.load(GOLDEN / f"fk_{rt}.npz")
    robot = URDFRobot(rt)
    assert robot.link_names == list(g["link_names"])
    np.testing.assert_allclose(robot.offsets.squeeze(0).squeeze(-1).numpy(), g["offsets"], rtol=1e-7)
    q, rot, trans = (t.cuda() for t in synth.fk_inputs(rt, 64))
    TOL = 1e-5
    assert _rel(robot.get_keypoints(q, rot, trans).cpu(), g["keypoints"]) < TOL
    assert _rel(robot.get_keypoints_only_fk(q).cpu(), g["keypoints_only_fk"]) < TOL
    nk = len(robot.link_names)
    for root in sorted({0, 3, nk - 1}):
        assert _rel(robot.get_keypoints_root(q, rot, trans, root=root).cpu(), g[f"keypoints_root{root}"]) < TOL, root
        assert _rel(robot.get_keypoints_only_fk_at_specific_root(q, root=root).cpu(), g[f"only_fk_root{root}"]) < TOL
        assert _rel(robot.get_rotation_at_specific_root(q, rot, trans, root=root).cpu(), g[f"rotation_root{root}"]) < TOL
    quat = synth.sym_uniform("fk_quat_" + rt, (64, 4), 1.0, 2).cuda()
    assert _rel(robot.get_keypoints(q, quat, trans).cpu(), g["keypoints_quat"]) < TOL
    # projection: keypoints with z < 0.2 m are excluded (pure conditioning)
    _, _, _, K = synth.inputs(64, seed=5)
    pts = torch.from_numpy(g["keypoints"])
    uv = point_projection_from_3d_tensor(K.cuda(), pts.cuda()).cpu().numpy()
    ok = g["keypoints"][..., 2] > 0.2
    assert ok.sum() > 0.5 * ok.size
    assert _rel(uv[ok], g["proj_tensor"][ok]) < TOL
    uv_np = point_projection_from_3d(K.numpy(), g["keypoints"])
    assert _rel(uv_np[ok], g["proj_numpy"][ok]) < TOL
    # properties: root=0 equals get_keypoints; the root keypoint maps to the translation
    assert torch.equal(robot.get_keypoints_root(q, rot, trans, root=0), robot.get_keypoints(q, rot, trans))
    if rt != "baxter":  # zero keypoint offsets: the root link maps to the identity
        kr = robot.get_keypoints_root(q, rot, trans, root=3)
        np.testing.assert_allclose(kr[:, 3].cpu().numpy(), trans.cpu().numpy(), atol=2e-6)


def test_fk_large_batch_vs_oracle(hrp_lib):
    """B=4096 (the FK micro-KAT size) against the oracle; ragged tail (B not a multiple of the CTA size)."""
    from horopose_b200 import synth
    from horopose_b200.robot import URDFRobot
    from oracle import horopose_oracle as O
    for rt, B in (("panda", 4096), ("baxter", 1000 + 13)):
        q, rot, trans = synth.fk_inputs(rt, B, seed=21)
        robot = URDFRobot(rt)
        ref = O.OracleRobot(rt, str(synth.URDF_PATHS[rt])).get_keypoints_root(q, rot, trans, root=3)
        got = robot.get_keypoints_root(q.cuda(), rot.cuda(), trans.cuda(), root=3).cpu()
        assert _rel(got, ref) < 1e-5


def test_fk_rejects_unsupported(hrp_lib):
    from horopose_b200 import _lib, synth
    from horopose_b200.robot import URDFRobot
    robot = URDFRobot("panda")
    q, rot, trans = (t.cuda() for t in synth.fk_inputs("panda", 4))
    with pytest.raises(_lib.HrpError):
        robot.get_keypoints(q, torch.zeros(4, 9).cuda(), trans)   # 9-D SVD rotations are outside the hot path
    with pytest.raises(_lib.HrpError):
        robot.get_keypoints(q.cpu(), rot.cpu(), trans.cpu())       # no CPU fallback


@pytest.mark.parametrize("logits_dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("rt", ROBOT_TYPES)
def test_heatmap_integral_vs_reference_golden(rt, logits_dtype, hrp_lib):
    """HeatmapIntegralPose drop-in against the reference's own outputs: on the caller's fp32 logits as they are (the
    default: only the order of the fp32 sums differs from the reference), and with the bf16 hand-off the network's final
    convolution uses (SURVEY.md section 9 finding 3)."""
    from horopose_b200 import arch, synth
    from horopose_b200.integral import HeatmapIntegralPose
    from make_golden import HEATMAP_STRESS_GAIN, heatmap_logits
    g = np.load(GOLDEN / f"integral_{rt}.npz")
    dof, nkpt, ref = arch.ROBOTS[rt]
    layer = HeatmapIntegralPose(backbone="resnet50", num_joints=nkpt, depth_dim=64, height_dim=64, width_dim=64,
                                norm_type="softmax", image_size=256.0, bbox_3d_shape=[1300, 1300, 1300], rootid=ref,
                                fixroot=True, logits_dtype=logits_dtype)
    for tag, gain in (("", 1.0), ("_peaky", HEATMAP_STRESS_GAIN)):
        hm = heatmap_logits(rt, 2, gain=gain)
        _, _, k, K = synth.inputs(2, seed=13)
        root_trans = torch.zeros(2, 3)
        root_trans[:, 2] = synth.range_uniform("root_z", (2,), 0.8, 2.5, 13)
        uvd, xyz = layer(hm.cuda(), root_trans=root_trans.cuda(), K=K.cuda())
        # the fp32 logits are rounded to bf16 on hand-off.  Realistic logits (|x| <= 3): 0.02 px / 0.1 mm budget
        # (survey: ~7e-4 px).  The "peaky" stress case is uniform noise in +-30, where a bf16 ulp is 0.125..0.25:
        # measured 0.07 px / 0.4 mm on B200, still far inside the 0.5 px / 1 mm end-to-end bars.
        px_tol, m_tol = (0.02, 1e-4) if tag == "" else (0.15, 1e-3)
        if logits_dtype == "fp32":   # no rounding of the logits: fp32 summation order and ex2.approx only
            px_tol, m_tol = 2e-3, 2e-5
        assert np.abs(uvd.cpu().numpy() - g["uvd" + tag]).max() * 256 < px_tol, tag
        assert np.abs(xyz.cpu().numpy() - g["xyz" + tag]).max() < m_tol, tag
        assert float(uvd[:, ref, 2].abs().max()) == 0.0


def test_fused_head_exact_on_bf16_logits(hrp_lib):
    """With logits that are exactly representable in bf16 the fused head must match the fp32 oracle to fp32
    round-off, including root translation, FK and the projections (all fp32 islands)."""
    from horopose_b200 import arch, ops, synth
    from horopose_b200.integral import run_head
    from horopose_b200.robot import URDFRobot
    from make_golden import heatmap_logits
    from oracle import horopose_oracle as O
    for rt, B in (("panda", 3), ("baxter", 2)):
        dof, nkpt, ref = arch.ROBOTS[rt]
        hm = heatmap_logits(rt, B, seed=9, gain=2.0).to(torch.bfloat16).float()
        _, _, k, K = synth.inputs(B, seed=17)
        depth = synth.range_uniform("hz", (B,), 0.8, 2.5, 17)
        q, rot, _ = synth.fk_inputs(rt, B, seed=17)
        root_trans = torch.zeros(B, 3)
        root_trans[:, 2] = depth
        uvd_ref, xyz_ref = O.heatmap_integral(hm, nkpt, K, root_trans, ref)
        root_uv_ref = (uvd_ref[:, ref, :2] + 0.5) * 256.0
        trans_ref = O.uvz2xyz_singlepoint(root_uv_ref, depth.view(-1, 1), K)
        orob = O.OracleRobot(rt, str(synth.URDF_PATHS[rt]))
        fk_ref = orob.get_keypoints_root(q, rot, trans_ref, root=ref) if ref else orob.get_keypoints(q, rot, trans_ref)
        hm_nhwc = ops.nchw_to_nhwc_bf16(hm.cuda(), cpad=nkpt * 64)
        r = run_head(hm_nhwc, K.cuda(), depth.cuda(), nkpt=nkpt, ref_kpt=ref, robot=URDFRobot(rt), pose=q.cuda(),
                     rot=rot.cuda(), want_uv=True)
        torch.cuda.synchronize()
        np.testing.assert_allclose(r["uvd"].cpu().numpy(), uvd_ref.numpy(), atol=3e-6)
        np.testing.assert_allclose(r["xyz_int"].cpu().numpy(), xyz_ref.numpy(), atol=1e-5)
        np.testing.assert_allclose(r["root_uv"].cpu().numpy(), root_uv_ref.numpy(), atol=1e-3)
        np.testing.assert_allclose(r["trans"].cpu().numpy(), trans_ref.numpy(), atol=1e-5)
        np.testing.assert_allclose(r["xyz_fk"].cpu().numpy(), fk_ref.numpy(), atol=1e-5)
        uv_int_ref = O.point_projection_from_3d(K, xyz_ref)
        np.testing.assert_allclose(r["uv_int"].cpu().numpy(), uv_int_ref.numpy(), atol=2e-3)
        # run twice: the per-image completion counters must self-reset
        r2 = run_head(hm_nhwc, K.cuda(), depth.cuda(), nkpt=nkpt, ref_kpt=ref, robot=URDFRobot(rt), pose=q.cuda(),
                      rot=rot.cuda())
        assert torch.equal(r2["uvd"], r["uvd"])


@pytest.mark.parametrize("logits_dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("rt", ["panda", "baxter"])
def test_heatmap_integral_backward_vs_autograd(rt, logits_dtype, hrp_lib):
    """Row f4, first piece: d(loss)/d(logits) through HeatmapIntegralPose (backward kernel re-using the forward's
    softmax statistics; xyz from uvd in torch) against torch autograd through the fp32 oracle on bf16-exact logits.
    Tolerance: 2e-4 of the largest gradient entry (fp32, ex2.approx vs exp)."""
    from horopose_b200 import arch, synth
    from horopose_b200.integral import HeatmapIntegralPose
    from make_golden import HEATMAP_STRESS_GAIN, heatmap_logits
    from oracle import horopose_oracle as O
    dof, nkpt, ref = arch.ROBOTS[rt]
    layer = HeatmapIntegralPose(backbone="resnet50", num_joints=nkpt, depth_dim=64, height_dim=64, width_dim=64,
                                norm_type="softmax", image_size=256.0, bbox_3d_shape=[1300, 1300, 1300], rootid=ref,
                                fixroot=True, logits_dtype=logits_dtype)
    B = 2
    _, _, k, K = synth.inputs(B, seed=13)
    root_trans = torch.zeros(B, 3)
    root_trans[:, 2] = synth.range_uniform("root_z", (B,), 0.8, 2.5, 13)
    G1 = synth.sym_uniform("g_uvd", (B, nkpt, 3), 1.0, 3)
    G2 = synth.sym_uniform("g_xyz", (B, nkpt, 3), 1.0, 4)
    # logits exactly representable in bf16 (the hand-off rounds nothing away; the reference-autograd digests were made on
    # them), and -- fp32 path only -- the raw fp32 logits against the oracle's autograd
    cases = [(1.0, True), (HEATMAP_STRESS_GAIN, True)] + ([(1.0, False)] if logits_dtype == "fp32" else [])
    for gain, exact in cases:
        logits = heatmap_logits(rt, B, gain=gain)
        if exact:
            logits = logits.bfloat16().float()
        with torch.enable_grad():   # (tests/golden/make_golden.py switches autograd off process-wide on import)
            x = logits.clone().cuda().requires_grad_(True)
            uvd, xyz = layer(x, root_trans=root_trans.cuda(), K=K.cuda())
            ((uvd * G1.cuda()).sum() + (xyz * G2.cuda()).sum()).backward()
            xo = logits.clone().requires_grad_(True)
            uvd_o, xyz_o = O.heatmap_integral(xo, nkpt, K, root_trans, ref, fixroot=True, image_size=256.0,
                                              depth_factor=float(layer.depth_factor))
            ((uvd_o * G1).sum() + (xyz_o * G2).sum()).backward()
        assert np.abs(uvd.detach().cpu().numpy() - uvd_o.detach().numpy()).max() < 1e-5
        assert np.abs(xyz.detach().cpu().numpy() - xyz_o.detach().numpy()).max() < 1e-5
        got, want = x.grad.cpu(), xo.grad
        assert got.shape == want.shape
        scale = float(want.abs().max())
        assert scale > 0
        err = float((got - want).abs().max())
        assert err < 2e-4 * scale, (gain, err, scale)
        # the reference keypoint's depth is pinned to 0 (integral.py:134): its depth expectation gets no gradient,
        # and gradients of one keypoint's logits sum to zero (softmax)
        assert float(got.double().reshape(B, nkpt, -1).sum(dim=2).abs().max()) < 5e-2 * scale   # 262144 fp32 terms
        if not exact:
            continue
        # and against the digest of the REFERENCE's own autograd (tests/golden/make_golden.py: golden_integral_backward)
        from make_golden import grad_digest
        gold = np.load(GOLDEN / f"integral_backward_{rt}.npz")
        tag = "" if gain == 1.0 else "_peaky"
        dg = grad_digest(got)
        amax = float(gold["absmax" + tag])
        assert abs(float(dg["absmax"]) - amax) < 2e-4 * amax
        assert np.abs(dg["samples"] - gold["samples" + tag]).max() < 2e-4 * amax
        for m in ("sum_hw", "sum_cw", "sum_ch"):   # sums of 4096 / nkpt*4096 entries: rounding noise adds up
            assert np.abs(dg[m] - gold[m + tag]).max() < 5e-2 * amax, m


@pytest.mark.parametrize("rt", ROBOT_TYPES)
def test_link_transforms_vs_reference_golden(rt, hrp_lib):
    """URDF.link_fk_batch over ALL links (urdf.py:3061-3149) and URDFRobot.get_TWL (urdf_robot.py:107-111) against the
    reference's own transforms (fixture `all_link_fk`), 1e-5 relative."""
    from horopose_b200 import synth
    from horopose_b200.robot import URDFRobot
    g = np.load(GOLDEN / f"fk_{rt}.npz")
    robot = URDFRobot(rt)
    q, _, _ = synth.fk_inputs(rt, 64)
    fk = robot.robot.link_fk_batch(q.cuda(), use_names=True)
    names = [str(n) for n in g["all_link_names"]]
    assert sorted(fk.keys()) == sorted(names)
    for i, n in enumerate(names):
        assert _rel(fk[n].cpu(), g["all_link_fk"][:, i]) < 1e-5, n
        assert torch.equal(fk[n][:, 3].cpu(), torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(64, 4)), n
    twl = robot.get_TWL(q.cuda())
    assert twl.shape == (64, len(robot.link_names), 4, 4)
    for k, n in enumerate(robot.link_names):
        assert _rel(twl[:, k].cpu(), g["all_link_fk"][:, names.index(n)]) < 1e-5, n
        # the pruned-tree kernel and the all-link kernel evaluate the same chain in the same order
        assert torch.equal(twl[:, k], fk[n]), n
    # keypoints = TWL * offset (urdf_robot.py:104): consistent with the FK kernel
    pts = (twl[:, :, :3, :3] @ robot.offsets.cuda() + twl[:, :, :3, 3:]).squeeze(-1)
    assert _rel(pts.cpu(), g["keypoints_only_fk"]) < 1e-5
    # ragged batch (not a multiple of the CTA size) and global_scale
    robot.global_scale = 2.0
    robot._handles = {}
    t2 = robot.get_TWL(q[:37].cuda())
    assert torch.allclose(t2[..., :3, 3], twl[:37, ..., :3, 3] * 2.0, rtol=1e-6, atol=0)
    assert torch.equal(t2[..., :3, :3], twl[:37, ..., :3, :3])


def test_geometry_operators_vs_oracle(hrp_lib):
    """Standalone uvd_to_xyz / uvz2xyz_singlepoint / get_intrinsic_matrix_batch (transforms.py:33-73,133-162,
    integral.py:56-73) against the CPU oracle: fp32, 1e-5 relative; K^-1 bit-exact (fp64 divisions, fp32 store)."""
    from horopose_b200 import synth, transforms as T
    from oracle import horopose_oracle as O
    B, N = 257, 17
    _, _, _, K = synth.inputs(B, seed=19)
    K[:, 1, 1] = K[:, 0, 0] * 1.013   # fx != fy
    uvd = synth.sym_uniform("geo_uvd", (B, N, 3), 0.25, 19)
    root = torch.cat([synth.sym_uniform("geo_txy", (B, 2), 0.3, 19), synth.range_uniform("geo_tz", (B, 1), 0.8, 2.5, 19)], 1)
    inv_ref = O.inv_intrinsics(K)
    inv = T.get_intrinsic_matrix_batch((K[:, 0, 0].cuda(), K[:, 1, 1].cuda()), (K[:, 0, 2].cuda(), K[:, 1, 2].cuda()),
                                       bsz=B, inv=True)
    assert torch.equal(inv.cpu(), inv_ref)
    fwd = T.get_intrinsic_matrix_batch((K[:, 0, 0], K[:, 1, 1]), (K[:, 0, 2], K[:, 1, 2]), bsz=B, inv=False)
    assert fwd.is_cuda and torch.equal(fwd.cpu(), K)
    xyz_ref = O.uvd_to_xyz(uvd, 256.0, inv_ref, root, 1.3)
    xyz = T.uvd_to_xyz(uvd.cuda(), 256.0, inv, root.cuda(), 1.3)
    assert _rel(xyz.cpu(), xyz_ref) < 1e-5
    rel = T.uvd_to_xyz(uvd.cuda(), 256.0, inv, root.cuda(), 1.3, return_relative=True)
    assert _rel(rel.cpu(), xyz_ref - root[:, None]) < 1e-5
    uv = (uvd[:, 0, :2] + 0.5) * 256.0
    z = root[:, 2:3]
    p_ref = O.uvz2xyz_singlepoint(uv, z, K)
    p = T.uvz2xyz_singlepoint(uv.cuda(), z.cuda(), K.cuda())
    assert _rel(p.cpu(), p_ref) < 1e-5
    with pytest.raises(Exception):
        T.uvd_to_xyz(uvd, 256.0, inv_ref, root, 1.3)   # CPU tensors: no CPU path


@pytest.mark.parametrize("rt", ROBOT_TYPES)
def test_fk_and_projection_backward_vs_reference_autograd(rt, hrp_lib):
    """Row f4: gradients of sum(pts * G) + 1e-3 * sum(uv * G2) w.r.t. (q, rot6d, trans) through get_keypoints_root and
    point_projection_from_3d_tensor, and of the only_fk variants w.r.t. q, against the gradients the REAL reference's
    autograd produced (tests/golden/fk_backward_*.npz, written by tests/golden/make_golden.py): 1e-4 relative."""
    from horopose_b200 import synth
    from horopose_b200.robot import URDFRobot, point_projection_from_3d_tensor
    g = np.load(GOLDEN / f"fk_backward_{rt}.npz")
    robot = URDFRobot(rt)
    q0, rot0, trans0 = (t.cuda() for t in synth.fk_inputs(rt, 16, seed=9))
    _, _, _, K = synth.inputs(16, seed=5)
    nk = len(robot.link_names)
    G = synth.sym_uniform("g_fk_" + rt, (16, nk, 3), 1.0, 5).cuda()
    G2 = synth.sym_uniform("g_uv_" + rt, (16, nk, 2), 1.0, 6).cuda()
    for root in sorted({0, 3, nk - 1}):
        with torch.enable_grad():   # (tests/golden/make_golden.py switches autograd off globally when imported)
            q, rot, trans = (t.clone().requires_grad_(True) for t in (q0, rot0, trans0))
            pts = robot.get_keypoints_root(q, rot, trans, root=root)
            uv = point_projection_from_3d_tensor(K.cuda(), pts)
            ((pts * G).sum() + 1e-3 * (uv * G2).sum()).backward()
        for nm, t in (("q", q), ("rot", rot), ("trans", trans)):
            assert _rel(t.grad.cpu(), g[f"grad_{nm}_root{root}"]) < 1e-4, (root, nm, _rel(t.grad.cpu(), g[f"grad_{nm}_root{root}"]))
        with torch.enable_grad():
            q = q0.clone().requires_grad_(True)
            pts = robot.get_keypoints_only_fk_at_specific_root(q, root=root) if root > 0 else robot.get_keypoints_only_fk(q)
            (pts * G).sum().backward()
        assert _rel(q.grad.cpu(), g[f"grad_q_only_fk_root{root}"]) < 1e-4, root
    # no gradient requested: the plain kernels run (no autograd graph)
    assert not robot.get_keypoints_root(q0, rot0, trans0, root=3).requires_grad
