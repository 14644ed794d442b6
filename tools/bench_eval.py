"""Rows f1-f3 measured on the GPU with the reference's CPU path (oracle port) timed beside them.
CUDA events, 3 warm-ups, inputs resident in HBM; the CPU legs run on a bounded sample and are scaled per image."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import synth as _synth
_synth.use_synthetic_urdfs()
from horopose_b200 import synth
from horopose_b200.metrics import compute_metrics_batch, summary_add_pck
from horopose_b200.pnp import BPnP_m3d
from horopose_b200.preprocess import crop_resize_batch
from horopose_b200.robot import URDFRobot
from oracle import eval_oracle as EO


def gpu_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def cpu_time(fn, runs=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(runs):
        fn()
    return (time.perf_counter() - t0) / runs


torch.set_num_threads(os.cpu_count() or 1)
B = 512
print(f"host cores: {os.cpu_count()}  device: {torch.cuda.get_device_name(0)}")
# ---- f1: crop + resize ----
frames, boxes, K, k_bbox = synth.crop_inputs(64, seed=5)
rep = B // 64
fr = torch.from_numpy(frames).cuda().repeat(rep, 1, 1, 1)
bx = torch.from_numpy(boxes).repeat(rep, 1)
Kd = torch.from_numpy(K).cuda().repeat(rep, 1, 1)
kb = torch.from_numpy(k_bbox).cuda().repeat(rep, 1)
t = gpu_time(lambda: crop_resize_batch(fr, bx, Kd, k_bbox=kb))
side = np.maximum(boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]).astype(np.int64)
src_bytes = int(((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) * 3).sum()) * rep
out_bytes = B * 3 * 256 * 256
tc = cpu_time(lambda: [EO.crop_resize(frames[b], boxes[b], K[b]) for b in range(16)]) / 16
print(f"f1 crop_resize      B={B}: {t * 1e6:8.1f} us  ({B / t:10.0f} img/s, {(src_bytes + out_bytes) / t / 1e9:7.1f} GB/s of bbox-in + crop-out bytes; "
      f"includes the host-side bbox validation of the shim)   CPU oracle: {tc * 1e3:6.2f} ms/img ({1 / tc:7.0f} img/s)")
# ---- f2: metrics ----
robot = URDFRobot("baxter")
q, rot, trans, gt_q, gt3, gt2, Km = (x.cuda() for x in synth.metric_inputs("baxter", B))
kw = dict(pred_joint=q, pred_rot=rot, pred_trans=trans, pred_depth=None, pred_xy=None, pred_xyz_integral=None,
          reference_keypoint_id=0)
t = gpu_time(lambda: compute_metrics_batch(robot, gt3, gt2, Km, gt_q, **kw))
kp = robot.get_keypoints(q, rot, trans).cpu().numpy()
args = [a.cpu().numpy() for a in (gt3, gt2, Km, q, gt_q)]
tc = cpu_time(lambda: EO.metrics_batch(kp, args[0], args[1], args[2], args[3], args[4], 0, False))
print(f"f2 metrics_batch    B={B} (baxter, FK + 2 metric kernels): {t * 1e6:8.1f} us   CPU oracle (numpy, FK excluded): {tc * 1e6:8.1f} us")
n = 100_000
d3 = (torch.rand(n, generator=torch.Generator().manual_seed(1)) * 0.12).cuda()
d2 = (torch.rand(n, generator=torch.Generator().manual_seed(2)) * 25.0).cuda()
t = gpu_time(lambda: summary_add_pck({"dis3d": d3, "dis2d": d2}), iters=10)
d3c, d2c = d3.cpu().numpy(), d2.cpu().numpy()
tc = cpu_time(lambda: EO.summary_add_pck(d3c, d2c), runs=1)
print(f"f2 summary_add_pck  n={n}: {t * 1e6:8.1f} us (incl. the 22-value read-back)   CPU oracle: {tc * 1e3:8.1f} ms")
# ---- f3: PnP ----
qq, rvec, tt, Kp, noise = synth.pnp_inputs("baxter", B, seed=8)
p3 = robot.get_keypoints_only_fk(qq.cuda())
th = rvec.norm(dim=1, keepdim=True)
k = rvec / th
Kx = torch.zeros(B, 3, 3)
Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0], Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
R = torch.eye(3)[None] + torch.sin(th)[:, :, None] * Kx + (1 - torch.cos(th))[:, :, None] * (Kx @ Kx)
cam = p3.cpu() @ R.transpose(1, 2) + tt[:, None, :]
uvw = cam @ Kp.T
p2 = (uvw[:, :, :2] / uvw[:, :, 2:3] + noise).cuda()
Kpc = Kp.cuda()
t = gpu_time(lambda: BPnP_m3d.apply(p2, p3, Kpc))
out = BPnP_m3d.apply(p2, p3, Kpc).cpu()
p2c, p3c = p2.cpu(), p3.cpu()
tc = cpu_time(lambda: EO.pnp_m3d(p2c[:64], p3c[:64], Kp), runs=2) / 64
ref = EO.pnp_m3d(p2c[:64], p3c[:64], Kp)
print(f"f3 BPnP_m3d         B={B} N=17: {t * 1e6:8.1f} us ({B / t:9.0f} poses/s)   CPU oracle (cv2 EPnP + LM): {tc * 1e6:7.1f} us/pose "
      f"({1 / tc:7.0f} poses/s)   max|diff| on 64 poses {float((out[:64] - ref).abs().max()):.2e}")
