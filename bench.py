#!/usr/bin/env python
"""Benchmark of the HoRoPose full-network inference forward pass on B200 (driver contract: one JSON line).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[3]): Kuka full model (ResNet-50 + HRNet-w32 + deconv head + fused head), bf16
tensor-core conv stack / fp32 head, synthetic seeded weights and inputs, GLOBAL batch 512 "batch-sharded at 1/2/4/8
B200": one "step" = one forward over the 512 images, rank r of N owning the contiguous shard
`shard.shard_range(512, r, N)` (512 / 256 / 128 / 64 images per GPU) -- STRONG scaling, the reference's analogue being
nn.DataParallel's scatter (scripts/test.py:149).  The path shards by image: no collective on the data path (SURVEY.md
section 8e); the per-rank results are gathered once after the timed region to check the assembly.  The same line also
carries the weak-scaling figure (512 images on every GPU, `weak_scaling`) and, at N = 8, BASELINE.json configs[4]
(Baxter, 2048 images over 8 GPUs).  `--scaling weak` makes the weak figure the headline instead.
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
GLOBAL_BATCH = 512      # BASELINE.json configs[3]: Kuka, batch 512, sharded over the GPUs
PER_GPU_BATCH = 512     # weak-scaling variant: this many images on every GPU
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
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ROBOT}_full_b{GLOBAL_BATCH}_global", "robot": ROBOT, "global_batch": GLOBAL_BATCH,
                   "step_sample_images": sample,
                   "note": "host CPU cores only (one process, rank 0): each step is a 64-image sample of the 512-image "
                           "workload, images/s is batch-size independent on the CPU"},
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
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: GLOBAL batch sharded over the ranks (BASELINE.json configs[3]); weak: --batch images per GPU")
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch of the weak-scaling variant")
    ap.add_argument("--overlap", type=int, default=0,
                    help="independent steps in flight per GPU (caller streams = plan replicas); 0 = auto: 3 up to 96 "
                         "images per GPU (a 64-image step cannot fill 148 SMs by itself), 2 up to 256, else 1")
    ap.add_argument("--robot", default=ROBOT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other scaling mode's figure")
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
    from horopose_b200 import _lib, arch, shard, synth
    from horopose_b200.integral import run_head
    from horopose_b200.models import get_rootNetwithRegInt_model
    from horopose_b200.pipeline import HostPipeline

    synth.use_synthetic_urdfs()  # explicit: mesh-free URDF fixtures (there is no network for the real descriptions)
    robot = args.robot
    dof, nkpt, ref = arch.ROBOTS[robot]
    margs = dict(backbone_name="resnet50", rootnet_backbone_name="hrnet32", n_iter=4, other_image_size=256.0,
                 bbox_3d_shape=[1300, 1300, 1300], reference_keypoint_id=ref, fix_root=True, rotation_dim=6)

    def build_model(rb, chunk, inflight):
        m = get_rootNetwithRegInt_model({"robot_type": rb, "pose_params": None, "cam_params": np.eye(4),
                                         "init_pose_from_mean": True},
                                        dict(margs, reference_keypoint_id=arch.ROBOTS[rb][2]))
        m.chunk, m.inflight = chunk, inflight
        m.load_state_dict(synth.full_state_dict(rb), strict=True)
        return m

    # synthetic inputs: uint8 crops as the reference's DataLoader delivers them (scripts/test.py:83-86); 64 distinct
    # seeded images, repeated to the batch size (201 MB of input per 512-image step >> the 126 MB L2 either way)
    x_reg_f, x_root_f, k_value, K = synth.inputs(64, seed=41)
    to_u8 = lambda t: (t * 255.0).round().clamp(0, 255).to(torch.uint8)
    base = (to_u8(x_reg_f), to_u8(x_root_f), k_value, K)

    def host_batch(lo, hi):
        """pinned host tensors of global images [lo, hi) (image g = distinct image g % 64)"""
        idx = torch.arange(lo, hi) % 64
        return tuple(t[idx].contiguous().pin_memory() for t in base)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_obj(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def measure(model, host, steps, warmup, overlap, profile=False, clocks=False):
        """Device-resident and host-buffer (e2e) throughput of `model` on this rank's batch `host`.  Returns per-rank
        dicts gathered over the ranks."""
        dev_in = tuple(t.to(dev) for t in host)
        n_img = host[0].shape[0]
        streams = [torch.cuda.Stream(device=dev) for _ in range(overlap)] if overlap > 1 else None

        def run_steps(n):
            outs = None
            if streams is None:
                for _ in range(n):
                    outs = model(*dev_in)
                return outs
            cur = torch.cuda.current_stream()
            for s_ in streams:
                s_.wait_stream(cur)
            for i in range(n):       # independent batches: step i+1 is enqueued on another stream / plan replica
                with torch.cuda.stream(streams[i % overlap]):
                    outs = model(*dev_in)
            for s_ in streams:
                cur.wait_stream(s_)
            return outs

        run_steps(warmup)
        barrier()
        sampler = ClockSampler(local) if clocks else None
        if sampler:
            sampler.start()
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if profile:
            torch.cuda.profiler.start()
        e0.record()
        outs = run_steps(steps)
        e1.record()
        barrier()
        if profile:
            torch.cuda.profiler.stop()
        ms_dev = e0.elapsed_time(e1)
        launches = _lib.launch_count() - launches0
        clk = None
        if sampler:
            # a short timed region (20 steps of a 64-image shard are 0.1 s) can end before nvidia-smi has delivered a
            # sample: keep the SAME load running, untimed, until a few samples exist
            extra, t_end = 0, time.time() + 4.0
            while len(sampler.lines) < 3 and time.time() < t_end:
                run_steps(max(2, overlap))
                torch.cuda.synchronize()
                extra += 1
            clk = sampler.stop()
            if extra:
                clk["note"] = ("timed region shorter than the sampling period: sampled while the same steps were "
                               "repeated right after it")
        # end to end through the host-buffer API: pinned uint8 crops + k + K are copied host->device and the eight outputs
        # are copied back, every step, inside the timed region.  HostPipeline.run_stream uploads batch i+1 while batch i
        # computes (what a DataLoader-fed eval loop does); the first upload and the last download are exposed and counted.
        pipe = HostPipeline(model)
        for host_out in pipe.run_stream(host for _ in range(warmup)):
            pass
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        wall0 = time.perf_counter()
        t0.record()
        for host_out in pipe.run_stream(host for _ in range(steps)):
            pass
        t1.record()
        barrier()
        ms_e2e = max(t0.elapsed_time(t1), (time.perf_counter() - wall0) * 1e3)
        h2d = host[0].numel() + host[1].numel() + host[2].numel() * 4 + host[3].numel() * 4
        d2h = sum(t.numel() * 4 for t in host_out)
        mine = {"rank": rank, "images": n_img, "ms_per_step": ms_dev / steps, "e2e_ms_per_step": ms_e2e / steps,
                "launches": int(launches), "h2d": int(h2d), "d2h": int(d2h),
                "sm_mhz": clk["sm_mhz"] if clk else None, "reasons": clk["reasons"] if clk else None}
        return gather_obj(mine), clk, outs, dev_in

    def summarise(per_rank, steps):
        imgs = sum(r["images"] for r in per_rank)
        ms = max(r["ms_per_step"] for r in per_rank)          # device-timed, max over ranks
        ms_e = max(r["e2e_ms_per_step"] for r in per_rank)
        return imgs, ms, imgs / (ms * 1e-3), ms_e, imgs / (ms_e * 1e-3)

    def default_overlap(b):
        """Independent steps kept in flight per GPU (caller streams = plan replicas).  Measured
        (profiles/r02_exp_small_shards.txt): 2-3 steps of single-lane PDL graphs beat one 5-lane graph by 6-9 % at 64-128
        images and 2 steps by 2-3 % at 256; at 512 every kernel fills the GPU and one step in flight is as fast."""
        return 3 if b <= 96 else (2 if b <= 256 else 1)   # (128 images: 2 and 3 tie on the device, 2 is better end to end)

    def batch_of(mode):
        if mode == "strong":
            lo, hi = shard.shard_range(args.global_batch, rank, world)
            return lo, hi
        return rank * args.batch, (rank + 1) * args.batch

    # ---------------- headline: the scaling mode asked for ----------------
    lo, hi = batch_of(args.scaling)
    B = hi - lo
    overlap = args.overlap if args.overlap > 0 else default_overlap(B)
    model = build_model(robot, B, overlap)
    per_rank, clocks, outs, dev_in = measure(model, host_batch(lo, hi), args.steps, args.warmup, overlap,
                                             profile=args.profile_range, clocks=True)
    total_imgs, ms_step, value, e2e_ms, e2e_value = summarise(per_rank, args.steps)
    x_reg_d, x_root_d, k_d, K_d = dev_in
    # result assembly (outside the timed region): every rank's keypoints gathered in shard order, as DataParallel's gather
    gathered = shard.gather_results(outs[7].float(), device=dev)
    assert gathered.shape[0] == total_imgs, (gathered.shape, total_imgs)

    # ---------------- the other scaling mode, fewer steps (identical to the headline at N = 1) ----------------
    secondary = None
    other_mode = "weak" if args.scaling == "strong" else "strong"
    if world > 1 and not args.no_secondary:
        lo2, hi2 = batch_of(other_mode)
        ov2 = args.overlap if args.overlap > 0 else default_overlap(hi2 - lo2)
        m2 = model if (hi2 - lo2 == B and ov2 == overlap) else build_model(robot, hi2 - lo2, ov2)
        pr2, _, _, _ = measure(m2, host_batch(lo2, hi2), max(5, args.steps // 2), args.warmup, ov2)
        imgs2, ms2, v2, e2e_ms2, e2e_v2 = summarise(pr2, max(5, args.steps // 2))
        secondary = {"scaling": other_mode, "value": v2, "unit": UNIT, "ms_per_step": ms2, "global_batch": imgs2,
                     "per_gpu_batch": hi2 - lo2, "overlap": ov2, "e2e": {"value": e2e_v2, "ms_per_step": e2e_ms2},
                     "per_rank_ms_per_step": [round(r["ms_per_step"], 4) for r in pr2]}
        if m2 is not model:
            del m2
    # ---------------- BASELINE.json configs[4]: Baxter, 2048 images over 8 GPUs (256 per GPU), N = 8 only ----------------
    baxter = None
    if world == 8 and not args.no_other_configs:
        bm = build_model("baxter", 256, 1)
        prb, _, _, _ = measure(bm, host_batch(rank * 256, (rank + 1) * 256), 5, args.warmup, 1)
        imgsb, msb, vb, e2e_msb, e2e_vb = summarise(prb, 5)
        baxter = {"workload": "baxter_full_b2048_over_8", "value": vb, "unit": UNIT, "ms_per_step": msb,
                  "global_batch": imgsb, "gflop_per_image": 40.07, "e2e": {"value": e2e_vb, "ms_per_step": e2e_msb},
                  "per_rank_ms_per_step": [round(r["ms_per_step"], 4) for r in prb]}
        del bm

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    launches = sum(r["launches"] for r in per_rank)                # kernels of libhrp_b200.so, all ranks, timed region
    h2d, d2h = sum(r["h2d"] for r in per_rank), sum(r["d2h"] for r in per_rank)   # bytes per (global) step
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
    tr = ROOT / "profiles" / "r02_ncu_traffic.json"
    if not tr.exists():
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
    hb = 512   # always the 512-image heatmap (2.1 GB > L2), whatever this rank's shard of the network batch is
    K_hd = K_d.repeat((hb + B - 1) // B, 1, 1)[:hb].contiguous()
    hm = torch.randn(hb, 64, 64, nkpt * 64, device=dev).to(torch.bfloat16)
    depth = torch.full((hb,), 1.5, device=dev)
    pose = torch.zeros(hb, dof, device=dev)
    rot = torch.tensor([[1.0, 0, 0, 0, 1, 0]], device=dev).repeat(hb, 1)
    for _ in range(3):
        run_head(hm, K_hd, depth, nkpt=nkpt, ref_kpt=ref, robot=model.robot, pose=pose, rot=rot)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_head = 20
    h0.record()
    for _ in range(n_head):
        run_head(hm, K_hd, depth, nkpt=nkpt, ref_kpt=ref, robot=model.robot, pose=pose, rot=rot)
    h1.record()
    torch.cuda.synchronize()
    head_ms = h0.elapsed_time(h1) / n_head
    head_gbs = HEATMAP_BYTES_PER_IMAGE[robot] * hb / (head_ms * 1e-3) / 1e9
    roofline_head = {"bound": "hbm", "achieved": head_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": head_gbs / peaks["hbm"], "traffic": None, "kernel": "hrp::head_kernel",
                     "note": f"{HEATMAP_BYTES_PER_IMAGE[robot]} B/image x {hb} images per launch "
                             f"({hm.numel() * 2 / 1e6:.0f} MB > L2), {head_ms * 1e3:.1f} us per launch"}
    ht = ROOT / "profiles" / "r02_ncu_head_traffic.json"
    if not ht.exists():
        ht = ROOT / "profiles" / "r01_ncu_head_traffic.json"   # (head_kernel is unchanged since that capture)
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
        # BASELINE.json configs[0]: Panda full model, batch 1, fp32 PyTorch on the CPU (beside latency_b1)
        _, _, t1 = cpu_forward_rate("panda", 1, runs=5, warmup=1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 forwards of 64 images ({sum(times):.1f} s), fp32 PyTorch oracle port of the reference forward",
               "panda_b1_latency_ms": 1e3 * statistics.median(t1)}

    act_gb = B * 91e6 / 1e9   # ~47 GB of intermediate tensors per 512 images (hrp_model_plan_memory)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"{robot}_full_b{total_imgs}_global" if args.scaling == "strong"
                   else f"{robot}_full_b{B}_per_gpu", "robot": robot, "global_batch": total_imgs,
                   "per_gpu_batch": [r["images"] for r in per_rank], "sharding": "contiguous batch shards, no collective",
                   "chunk": model.chunk, "steps_in_flight_per_gpu": overlap,
                   "input": "uint8 crops 2x(B,3,256,256); 64 distinct seeded images repeated to the batch",
                   "l2": f"no explicit flush: every step streams ~{act_gb:.0f} GB of activations per GPU through the "
                         f"126 MB L2 between two reads of its {2 * B * 3 * 256 * 256 / 1e6:.0f} MB of inputs"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms},
        "per_rank": [{k_: (round(v_, 4) if isinstance(v_, float) else v_) for k_, v_ in r.items()
                      if k_ in ("rank", "images", "ms_per_step", "e2e_ms_per_step", "sm_mhz", "reasons")} for r in per_rank],
        "weak_scaling" if other_mode == "weak" else "strong_scaling": secondary,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_head": roofline_head,
        "roofline_layers": layers,
        "cpu_baseline": cpu,
        "latency_b1": latency,
        "other_configs": other,
        "baxter_b2048_over_8": baxter,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
