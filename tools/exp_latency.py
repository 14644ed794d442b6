"""Batch-1 latency of the Panda full model (p50 over 200 forwards) under the current HRP_* environment."""
import os
import statistics
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import synth as _synth
_synth.use_synthetic_urdfs()
from horopose_b200 import synth
from horopose_b200.models import get_rootNetwithRegInt_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
args = dict(backbone_name="resnet50", rootnet_backbone_name="hrnet32", n_iter=4, other_image_size=256.0,
            bbox_3d_shape=[1300, 1300, 1300], reference_keypoint_id=3, fix_root=True, rotation_dim=6)
pm = get_rootNetwithRegInt_model({"robot_type": "panda", "pose_params": None, "cam_params": np.eye(4),
                                  "init_pose_from_mean": True}, args)
pm.chunk, pm.inflight = B, 1
pm.load_state_dict(synth.full_state_dict("panda"), strict=True)
x_reg, x_root, k, K = (t.cuda() for t in synth.inputs(B, seed=11))
x_reg, x_root = (x_reg * 255).to(torch.uint8), (x_root * 255).to(torch.uint8)
for _ in range(10):
    out = pm(x_reg, x_root, k, K)
torch.cuda.synchronize()
lat = []
for _ in range(200):
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    out = pm(x_reg, x_root, k, K)
    s1.record()
    s1.synchronize()
    lat.append(s0.elapsed_time(s1))
env = {k: v for k, v in os.environ.items() if k.startswith("HRP_")}
print(f"B={B} env={env}: p50 {statistics.median(lat):.3f} ms  p90 {sorted(lat)[179]:.3f} ms  min {min(lat):.3f} ms  "
      f"pose[0,:3]={out[0][0, :3].tolist()}")
