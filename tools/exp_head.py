"""Fused head alone on > L2 heatmaps: time per launch and GB/s (python tools/exp_head.py; HRP_HEAD_RING=0 for the register loop)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import arch
from horopose_b200.integral import run_head
from horopose_b200.robot import URDFRobot

for robot, B in (("kuka", 512), ("panda", 512), ("baxter", 256), ("kuka", 64)):
    dof, nkpt, ref = arch.ROBOTS[robot]
    rob = URDFRobot(robot)
    hm = torch.randn(B, 64, 64, nkpt * 64, device="cuda").to(torch.bfloat16)
    K = torch.tensor([[[500.0, 0, 128], [0, 500, 128], [0, 0, 1]]], device="cuda").repeat(B, 1, 1)
    depth = torch.full((B,), 1.5, device="cuda")
    pose = torch.zeros(B, dof, device="cuda")
    rot = torch.tensor([[1.0, 0, 0, 0, 1, 0]], device="cuda").repeat(B, 1)
    f = lambda: run_head(hm, K, depth, nkpt=nkpt, ref_kpt=ref, robot=rob, pose=pose, rot=rot)
    for _ in range(3):
        out = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = f()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    byt = hm.numel() * 2
    chk = float(out[0].double().sum()) if isinstance(out, (tuple, list)) else 0.0
    print(f"ring={os.environ.get('HRP_HEAD_RING', '1')} {robot:7s} B={B:4d}: {t * 1e6:8.1f} us  {byt / t / 1e9:7.1f} GB/s  checksum {chk:.6f}")
