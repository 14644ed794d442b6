"""Run the fused head a few times on a > L2 heatmap (for ncu): python tools/prof_head.py [robot] [B]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import synth as _synth
_synth.use_synthetic_urdfs()
from horopose_b200 import arch
from horopose_b200.integral import run_head
from horopose_b200.robot import URDFRobot

robot = sys.argv[1] if len(sys.argv) > 1 else "kuka"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dof, nkpt, ref = arch.ROBOTS[robot]
rob = URDFRobot(robot)
hm = torch.randn(B, 64, 64, nkpt * 64, device="cuda").to(torch.bfloat16)
K = torch.tensor([[[500.0, 0, 128], [0, 500, 128], [0, 0, 1]]], device="cuda").repeat(B, 1, 1)
depth = torch.full((B,), 1.5, device="cuda")
pose = torch.zeros(B, dof, device="cuda")
rot = torch.tensor([[1.0, 0, 0, 0, 1, 0]], device="cuda").repeat(B, 1)
for _ in range(4):
    run_head(hm, K, depth, nkpt=nkpt, ref_kpt=ref, robot=rob, pose=pose, rot=rot)
torch.cuda.synchronize()
print("done head", robot, B)
