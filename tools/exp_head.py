"""Fused head alone on > L2 heatmaps (forward, and the heatmap-integral backward): time per launch and GB/s."""
import os
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
    print(f"forward  {robot:7s} B={B:4d}: {t * 1e6:8.1f} us  {byt / t / 1e9:7.1f} GB/s  checksum {chk:.6f}")

# ---- backward of the heatmap integral (row f4): bf16 gradient out, heatmap read + gradient write ----
import ctypes as C
from horopose_b200 import _lib
for robot, B in (("kuka", 512), ("baxter", 256)):
    dof, nkpt, ref = arch.ROBOTS[robot]
    hm = torch.randn(B, 64, 64, nkpt * 64, device="cuda").to(torch.bfloat16)
    K = torch.tensor([[[500.0, 0, 128], [0, 500, 128], [0, 0, 1]]], device="cuda").repeat(B, 1, 1)
    depth = torch.full((B,), 1.5, device="cuda")
    need = C.c_int64(0)
    _lib.check(_lib.lib().hrp_head_workspace_bytes(B, nkpt, C.byref(need)))
    ws = torch.zeros(need.value, dtype=torch.uint8, device="cuda")
    r = run_head(hm, K, depth, nkpt=nkpt, ref_kpt=ref, workspace=ws)
    g = torch.randn(B, nkpt, 3, device="cuda")
    out = torch.empty_like(hm)
    f = lambda: _lib.check(_lib.lib().hrp_head_backward_heatmap(
        C.c_void_p(hm.data_ptr()), C.c_void_p(r["uvd"].data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(ws.data_ptr()),
        C.c_int64(ws.numel()), C.c_int32(B), C.c_int32(nkpt), C.c_int32(ref), C.c_int32(1), C.c_int32(0),
        C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    print(f"backward {robot:7s} B={B:4d}: {t * 1e6:8.1f} us  {2 * hm.numel() * 2 / t / 1e9:7.1f} GB/s (read + write)")
