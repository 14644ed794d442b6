#!/usr/bin/env python
"""Benchmark of the HoRoPose full-network inference forward pass on B200 (driver contract: one JSON line).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[3]): Kuka full model (ResNet-50 + HRNet-w32 + deconv head + fused head), bf16
tensor-core conv stack / fp32 head, synthetic seeded weights and inputs.  One "step" = one forward over the
per-GPU batch of 512 images; under torchrun every rank runs the same per-GPU batch (weak scaling: the path shards
by image, no collective on the data path -- SURVEY.md section 8e).
  value  = images/s over all ranks with inputs already resident in HBM (uint8 crops, as the reference's loader
           delivers them), CUDA events, max over ranks.
  e2e    = the same metric through the public API with HOST buffers: pinned uint8 crops + k + K are copied to the
           device and the eight outputs are copied back inside the timed region, every step.
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "full_model_images_per_sec"
UNIT = "images/s"
ROBOT = "kuka"
PER_GPU_BATCH = 512
# algorithmic work per image: conv / deconv / linear GMACs of the Kuka full model
GFLOP_PER_IMAGE = 38.86
HEATMAP_BYTES_PER_IMAGE = {"panda": 3_670_016, "kuka": 4_194_304, "baxter": 8_912_896}


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]), tf_sust=float(d["bf16_tflops_sustained"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # samples under load = upper half of the observed clocks
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _dist():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _oracle_model(robot):
    import horopose_b200  # noqa: F401
    from horopose_b200 import synth
    from oracle import horopose_oracle as O
    return synth.full_state_dict(robot), O.OracleRobot(robot, str(synth.URDF_PATHS[robot])), O, synth


def cpu_forward_rate(robot, batch, runs, warmup=1):
    """images/s of the reference algorithm (oracle port, fp32 PyTorch on the host cores)."""
    import torch
    sd, orob, O, synth = _oracle_model(robot)
    torch.set_num_threads(os.cpu_count() or 1)
    x_reg, x_root, k, K = synth.inputs(batch, seed=31)
    times = []
    with torch.no_grad():
        for i in range(warmup + runs):
            t0 = time.perf_counter()
            O.full_forward(sd, orob, x_reg, x_root, k, K)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return batch / statistics.median(times), torch.get_num_threads(), times


def run_reference(args):
    """--impl reference: the reference's own algorithm on the box's host cores (oracle port; the GPU box has no
    /root/reference).  Each step is a bounded sample of the workload: one forward over 64 images."""
    rank, world, _ = _dist()
    if rank != 0:
        return
    sample = 64
    rate, cores, times = cpu_forward_rate(ROBOT, sample, runs=args.steps, warmup=args.warmup)
    ms = 1e3 * statistics.median(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ROBOT}_full_b{PER_GPU_BATCH}_per_gpu", "robot": ROBOT, "per_gpu_batch": PER_GPU_BATCH,
                   "step_sample_images": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} forwards of {sample} images, fp32 PyTorch oracle port on the host"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch")
    ap.add_argument("--robot", default=ROBOT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="wrap the timed device-resident steps in cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    rank, world, local = _dist()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import horopose_b200  # noqa: F401
    from horopose_b200 import _lib, arch, synth
    from horopose_b200.integral import run_head
    from horopose_b200.models import get_rootNetwithRegInt_model
    from horopose_b200.pipeline import HostPipeline

    robot = args.robot
    B = args.batch
    dof, nkpt, ref = arch.ROBOTS[robot]
    margs = dict(backbone_name="resnet50", rootnet_backbone_name="hrnet32", n_iter=4, other_image_size=256.0,
                 bbox_3d_shape=[1300, 1300, 1300], reference_keypoint_id=ref, fix_root=True, rotation_dim=6)
    init = {"robot_type": robot, "pose_params": None, "cam_params": np.eye(4), "init_pose_from_mean": True}
    model = get_rootNetwithRegInt_model(init, margs)
    model.load_state_dict(synth.full_state_dict(robot), strict=True)

    # synthetic inputs: uint8 crops as the reference's DataLoader delivers them (scripts/test.py:83-86)
    x_reg_f, x_root_f, k_value, K = synth.inputs(min(B, 64), seed=41)
    reps = (B + x_reg_f.shape[0] - 1) // x_reg_f.shape[0]
    to_u8 = lambda t: (t * 255.0).round().clamp(0, 255).to(torch.uint8).repeat(reps, 1, 1, 1)[:B].contiguous()
    x_reg_h, x_root_h = to_u8(x_reg_f).pin_memory(), to_u8(x_root_f).pin_memory()
    k_h = k_value.repeat(reps)[:B].contiguous().pin_memory()
    K_h = K.repeat(reps, 1, 1)[:B].contiguous().pin_memory()
    x_reg_d, x_root_d, k_d, K_d = (t.to(dev) for t in (x_reg_h, x_root_h, k_h, K_h))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput ("value") ----------------
    for _ in range(args.warmup):
        outs = model(x_reg_d, x_root_d, k_d, K_d)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profile_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        outs = model(x_reg_d, x_root_d, k_d, K_d)
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---------------- end-to-end through the host-buffer API ("e2e") ----------------
    # One step = one host batch: pinned uint8 crops + k + K are copied host->device and the eight outputs are copied
    # back, every step, inside the timed region.  HostPipeline.run_stream uploads batch i+1 while batch i computes
    # (what a DataLoader-fed eval loop does); the first upload and the last download are exposed and counted.
    pipe = HostPipeline(model)
    host_batch = (x_reg_h, x_root_h, k_h, K_h)
    for host_out in pipe.run_stream(host_batch for _ in range(args.warmup)):
        pass
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.perf_counter()
    t0.record()
    for host_out in pipe.run_stream(host_batch for _ in range(args.steps)):
        pass
    t1.record()
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    e2e_ms = max_over_ranks(max(t0.elapsed_time(t1), wall_ms))
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    h2d = x_reg_h.numel() + x_root_h.numel() + k_h.numel() * 4 + K_h.numel() * 4
    d2h = sum(t.numel() * 4 for t in host_out)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    # ---------------- roofline of the dominant kernel family (tcgen05 conv stack) ----------------
    gflop_img = {"panda": 38.72, "kuka": 38.86, "baxter": 40.07}[robot]
    achieved_tf = gflop_img * 1e9 * B / (ms_step * 1e-3) / 1e12
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["tf_sust"], "traffic": None,
                "kernel": "hrp::conv_gemm_kernel (all conv / deconv launches of one step)",
                "note": f"algorithmic {gflop_img} GFLOP/image x {B} images / measured step time "
                        f"(whole step incl. packing, pooling and head kernels: a lower bound); peak = "
                        f"{peaks['src']} sustained bf16"}
    # per-layer view, measured live: every op of the plan timed with CUDA events (3 back-to-back launches each, on the
    # plan's stream) through hrp_model_profile; each conv is placed on ITS roofline (max of FLOPs / tensor peak and
    # algorithmic bytes / HBM peak).  `traffic` of the dominant kernel comes from the committed ncu --set full capture.
    layers = None
    try:
        import ctypes as C
        buf = C.create_string_buffer(1 << 20)
        _lib.check(_lib.lib().hrp_model_profile(model._handle, min(B, model.chunk), 3, buf, len(buf)))
        rows = [r.split("\t") for r in buf.value.decode().strip().split("\n")]
        convs = [r for r in rows if r[1] == "conv"]
        t_meas = sum(float(r[12]) for r in rows)
        t_ideal = sum(max(float(r[15]) / (peaks["tf_sust"] * 1e6), float(r[16]) / (peaks["hbm"] * 1e3)) for r in convs)
        top = sorted(convs, key=lambda r: -float(r[12]))[:5]
        layers = {"ops": len(rows), "sum_us": t_meas, "per_layer_roofline_us": t_ideal,
                  "frac_of_per_layer_roofline": t_ideal / t_meas,
                  "top": [{"layer": r[0], "kernel": r[10], "us": float(r[12]), "tflops": float(r[13]), "gbs": float(r[14]),
                           "bound": "tensor" if float(r[15]) / (peaks["tf_sust"] * 1e6) > float(r[16]) / (peaks["hbm"] * 1e3) else "hbm",
                           "frac": max(float(r[13]) / peaks["tf_sust"], float(r[14]) / peaks["hbm"])} for r in top]}
    except Exception as e:  # the per-layer view is diagnostic: never fail the bench line on it
        layers = {"error": str(e)}
    tr = ROOT / "profiles" / "r01_ncu_traffic.json"
    if tr.exists():
        try:
            t = json.loads(tr.read_text())
            if t.get("robot") == robot and int(t.get("batch", 0)) > 0:
                roofline["traffic"] = float(t["dram_bytes_per_step"]) * B / int(t["batch"])
                roofline["traffic_note"] = (f"dram__bytes_read+write summed over the {t.get('kernels')} kernels of one "
                                            f"{t['batch']}-image step (ncu, {t.get('source')}), scaled to {B} images; "
                                            f"algorithmic bytes of the same step: {t.get('algorithmic_bytes_per_step')}")
        except Exception:
            pass
    # fused head alone (memory-bound): standalone launches of the same kernel on a > L2 heatmap, CUDA events
    hb = min(B, 512)
    hm = torch.randn(hb, 64, 64, nkpt * 64, device=dev).to(torch.bfloat16)
    depth = torch.full((hb,), 1.5, device=dev)
    pose = torch.zeros(hb, dof, device=dev)
    rot = torch.tensor([[1.0, 0, 0, 0, 1, 0]], device=dev).repeat(hb, 1)
    for _ in range(3):
        run_head(hm, K_d[:hb], depth, nkpt=nkpt, ref_kpt=ref, robot=model.robot, pose=pose, rot=rot)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_head = 20
    h0.record()
    for _ in range(n_head):
        run_head(hm, K_d[:hb], depth, nkpt=nkpt, ref_kpt=ref, robot=model.robot, pose=pose, rot=rot)
    h1.record()
    torch.cuda.synchronize()
    head_ms = h0.elapsed_time(h1) / n_head
    head_gbs = HEATMAP_BYTES_PER_IMAGE[robot] * hb / (head_ms * 1e-3) / 1e9
    roofline_head = {"bound": "hbm", "achieved": head_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": head_gbs / peaks["hbm"], "traffic": None, "kernel": "hrp::head_kernel",
                     "note": f"{HEATMAP_BYTES_PER_IMAGE[robot]} B/image x {hb} images per launch "
                             f"({hm.numel() * 2 / 1e6:.0f} MB > L2), {head_ms * 1e3:.1f} us per launch"}
    ht = ROOT / "profiles" / "r01_ncu_head_traffic.json"
    if ht.exists():
        try:
            t = json.loads(ht.read_text())
            if t.get("robot") == robot and int(t.get("batch", 0)) > 0:
                roofline_head["traffic"] = float(t["dram_bytes_per_launch"]) * hb / int(t["batch"])
                roofline_head["traffic_note"] = f"{t.get('source')}, scaled to {hb} images; algorithmic bytes per launch: " \
                                                f"{HEATMAP_BYTES_PER_IMAGE[robot] * hb}"
        except Exception:
            pass
    del hm

    # ---------------- batch-1 p50 latency (BASELINE.json configs[2], Panda) ----------------
    latency = None
    if not args.no_latency:
        pm = get_rootNetwithRegInt_model({"robot_type": "panda", "pose_params": None, "cam_params": np.eye(4),
                                          "init_pose_from_mean": True}, dict(margs, reference_keypoint_id=3))
        pm.chunk, pm.inflight = 1, 1
        pm.load_state_dict(synth.full_state_dict("panda"), strict=True)
        a, b, c, d = x_reg_d[:1], x_root_d[:1], k_d[:1], K_d[:1]
        for _ in range(10):
            pm(a, b, c, d)
        torch.cuda.synchronize()
        lat = []
        for _ in range(200):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            pm(a, b, c, d)
            s1.record()
            s1.synchronize()
            lat.append(s0.elapsed_time(s1))
        latency = {"workload": "panda_full_b1", "p50_ms": statistics.median(lat), "p90_ms": sorted(lat)[179],
                   "iters": 200}
        del pm

    # ---------------- the other BASELINE.json configs (device-resident, informational; N = 1 only) ----------------
    # configs[1] Panda depthnet B=256, configs[2] Panda full B=64, configs[4] Baxter full B=256 per GPU
    other = None
    if not args.no_other_configs and world == 1:
        def rate(fn, n_img, iters=5):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            return n_img * iters / (a.elapsed_time(b) * 1e-3)

        def rep_to(t, n):
            r = (n + t.shape[0] - 1) // t.shape[0]
            return t.repeat(r, *([1] * (t.dim() - 1)))[:n].contiguous()

        other = {}
        try:
            from horopose_b200.models import get_rootnet
            dn = get_rootnet("hrnet32")
            dn.chunk, dn.inflight = 256, 1
            dn.load_state_dict(synth.depthnet_state_dict(), strict=True)
            xd, kd = rep_to(x_root_d, 256).float() / 255.0, rep_to(k_d, 256)
            other["panda_depthnet_b256"] = {"images_per_sec": rate(lambda: dn(xd, kd), 256), "gflop_per_image": 23.30}
            del dn, xd
            for name, rb, nb, gf in (("panda_full_b64", "panda", 64, 38.72), ("baxter_full_b256", "baxter", 256, 40.07)):
                mm = get_rootNetwithRegInt_model({"robot_type": rb, "pose_params": None, "cam_params": np.eye(4),
                                                  "init_pose_from_mean": True},
                                                 dict(margs, reference_keypoint_id=arch.ROBOTS[rb][2]))
                mm.chunk, mm.inflight = nb, 1
                mm.load_state_dict(synth.full_state_dict(rb), strict=True)
                a_, b_, c_, d_ = rep_to(x_reg_d, nb), rep_to(x_root_d, nb), rep_to(k_d, nb), rep_to(K_d, nb)
                other[name] = {"images_per_sec": rate(lambda: mm(a_, b_, c_, d_), nb), "gflop_per_image": gf}
                del mm
        except Exception as e:  # informational: never fail the bench line on it
            other["error"] = str(e)

    # ---------------- reference algorithm on the host cores (bounded sample) ----------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only
        rate, cores, times = cpu_forward_rate(robot, 64, runs=3, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 forwards of 64 images ({sum(times):.1f} s), fp32 PyTorch oracle port of the reference forward"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{robot}_full_b{B}_per_gpu", "robot": robot, "per_gpu_batch": B,
                   "global_batch": B * world, "chunk": model.chunk, "inflight": model.inflight,
                   "input": "uint8 crops 2x(B,3,256,256)",
                   "l2": f"inputs {2 * B * 3 * 256 * 256 / 1e6:.0f} MB/step and activations are larger than the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_head": roofline_head,
        "roofline_layers": layers,
        "cpu_baseline": cpu,
        "latency_b1": latency,
        "other_configs": other,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
