"""Per-layer timing of the network plan (hrp_model_profile) + chunk / inflight sweep of the whole forward."""
import ctypes as C
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import horopose_b200  # noqa
from horopose_b200 import synth as _synth
_synth.use_synthetic_urdfs()
from horopose_b200 import _lib, arch, synth
from horopose_b200.models import get_rootNetwithRegInt_model


def make(robot, chunk, inflight):
    ref = arch.ROBOTS[robot][2]
    margs = dict(backbone_name="resnet50", rootnet_backbone_name="hrnet32", n_iter=4, other_image_size=256.0,
                 bbox_3d_shape=[1300, 1300, 1300], reference_keypoint_id=ref, fix_root=True, rotation_dim=6)
    m = get_rootNetwithRegInt_model({"robot_type": robot, "pose_params": None, "cam_params": np.eye(4),
                                     "init_pose_from_mean": True}, margs)
    m.chunk, m.inflight = chunk, inflight
    m.load_state_dict(synth.full_state_dict(robot), strict=True)
    return m


def category(name):
    if name.startswith("rootnet_backbone."):
        n = name[len("rootnet_backbone."):]
        if ".branches.0." in n:
            return "hr branch0 (32ch@64)"
        if ".branches.1." in n:
            return "hr branch1 (64ch@32)"
        if ".branches.2." in n:
            return "hr branch2 (128ch@16)"
        if ".branches.3." in n:
            return "hr branch3 (256ch@8)"
        if "fuse_layers" in n:
            return "hr fuse convs"
        if n.startswith("layer1") or n.startswith("conv") or n.startswith("transition"):
            return "hr stem+layer1+transitions"
        return "hr cls head"
    if name.startswith("reg_backbone."):
        n = name[len("reg_backbone."):]
        return "rn " + n.split(".")[0]
    if name.startswith("deconv"):
        return "deconv"
    if name.startswith("final_layer"):
        return "final 1x1"
    return name


def profile(robot="kuka", B=32):
    m = make(robot, B, 1)
    x_reg, x_root, k, K = (t.cuda() for t in synth.inputs(min(B, 32), seed=3))
    reps = (B + 31) // 32
    x_reg, x_root = x_reg.repeat(reps, 1, 1, 1)[:B], x_root.repeat(reps, 1, 1, 1)[:B]
    k, K = k.repeat(reps)[:B], K.repeat(reps, 1, 1)[:B]
    m(x_reg, x_root, k, K)
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 20)
    _lib.check(_lib.lib().hrp_model_profile(m._handle, B, 5, buf, len(buf)))
    rows = [r.split("\t") for r in buf.value.decode().strip().split("\n")]
    tot = sum(float(r[12]) for r in rows)
    cat = defaultdict(lambda: [0.0, 0.0, 0])
    for r in rows:
        c = cat[category(r[0])]
        c[0] += float(r[12])
        c[1] += float(r[12]) * float(r[13]) * 1e6 if r[1] == "conv" else 0.0  # flops
        c[2] += 1
    print(f"== per-op profile {robot} B={B}: {len(rows)} ops, sum {tot:.0f} us -> {B / tot * 1e6:.0f} img/s if serial")
    for name, (us, fl, n) in sorted(cat.items(), key=lambda kv: -kv[1][0]):
        print(f"  {name:30s} {n:4d} ops {us:9.1f} us {us / tot * 100:5.1f}%  {fl / us * 1e-6 if us else 0:7.1f} TFLOP/s")
    # per-layer roofline: ideal time = max(flops / tensor peak, bytes / HBM peak) with the measured peaks
    import json
    pk = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) \
        if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    ideal = t_mem = t_mma = 0.0
    for r in rows:
        if r[1] != "conv":
            continue
        tm, tb = float(r[15]) / (pk["bf16_tflops_sustained"] * 1e6), float(r[16]) / (pk["hbm_gbs"] * 1e3)
        ideal += max(tm, tb)
        t_mma += tm
        t_mem += tb
    print(f"  roofline sums over conv ops: tensor-only {t_mma:.0f} us, hbm-only {t_mem:.0f} us, per-layer max {ideal:.0f} us "
          f"-> measured/ideal = {tot / ideal:.2f}")
    out = Path(__file__).resolve().parent.parent / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / f"per_op_{robot}_{B}.tsv", "w") as f:
        f.write("name\tkind\tlane\tin\tCin\tCout\tout\ttaps\tn_tile\tepi\tvariant\tstages\tus\tTFLOPs\tGBs\tflops\tbytes\n")
        for r in rows:
            f.write("\t".join(r) + "\n")
    print("  top ops:")
    for r in sorted(rows, key=lambda r: -float(r[12]))[:25]:
        print("   ", "  ".join(r))
    return rows


def sweep(robot="kuka", B=512):
    x_reg, x_root, k, K = synth.inputs(32, seed=3)
    to_u8 = lambda t: (t * 255).round().to(torch.uint8).repeat(B // 32, 1, 1, 1).cuda()
    xr, xo = to_u8(x_reg), to_u8(x_root)
    kk, KK = k.repeat(B // 32).cuda(), K.repeat(B // 32, 1, 1).cuda()
    import os
    cfgs = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("HRP_SWEEP", "256:1,512:1").split(",")]
    for chunk, inflight in cfgs:
        m = make(robot, chunk, inflight)
        for _ in range(2):
            m(xr, xo, kk, KK)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            m(xr, xo, kk, KK)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"  chunk {chunk:4d} inflight {inflight}: {ms:7.2f} ms / {B} images -> {B / ms * 1e3:8.0f} img/s")
        del m


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "profile"):
        profile("kuka", int(sys.argv[2]) if len(sys.argv) > 2 else 32)
    if what in ("all", "sweep"):
        sweep()
