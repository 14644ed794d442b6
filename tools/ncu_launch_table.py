"""Summarise an ncu launch list (gpu__time_duration, sm__pipe_tensor_cycles_active, dram bytes per launch) of one bench
step: per-kernel-family share of the step, time-weighted tensor-pipe utilisation (whole step / conv kernels only) and
DRAM traffic.    python tools/ncu_launch_table.py launches.csv [out.json [traffic.json per_op.tsv source-label]]
(traffic.json is the file bench.py reads for roofline.traffic: the step's DRAM bytes next to its algorithmic bytes)"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}
per, names = collections.defaultdict(dict), {}
for r in rows:
    per[r[0]][r[-3]] = float(r[-1].replace(",", "")) * UNIT.get(r[-2], 1.0)
    names[r[0]] = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("hrp::", "")
T, P, D = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_"
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, m in per.items():
    a = agg[names[i]]
    a[0] += 1
    a[1] += m.get(T, 0.0)
    a[2] += m.get(T, 0.0) * m.get(P, 0.0)
    a[3] += m.get(D + "read.sum", 0.0) + m.get(D + "write.sum", 0.0)
tot_t = sum(a[1] for a in agg.values())
tot_p = sum(a[2] for a in agg.values())
tot_b = sum(a[3] for a in agg.values())
conv = {k: a for k, a in agg.items() if k.startswith("conv_")}
ct, cp = sum(a[1] for a in conv.values()), sum(a[2] for a in conv.values())
print(f"{len(per)} launches, {tot_t / 1e6:.2f} ms serialised under ncu, {tot_b / 1e9:.2f} GB DRAM traffic")
print(f"time-weighted sm__pipe_tensor_cycles_active: whole step {tot_p / tot_t:.1f} %, conv kernels only {cp / max(ct, 1):.1f} %"
      f" (conv share of the step {ct / tot_t * 100:.1f} %)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:58s} {a[0]:4d} launches {a[1] / 1e3:9.1f} us {a[1] / tot_t * 100:5.1f}%  tensor pipe {a[2] / max(a[1], 1):5.1f} %  "
          f"{a[3] / 1e6:9.1f} MB")
if len(sys.argv) > 2:
    json.dump({"kernels": len(per), "serialised_ms": tot_t / 1e6, "dram_bytes_per_step": tot_b,
               "tensor_pipe_pct_step": tot_p / tot_t, "tensor_pipe_pct_conv": cp / max(ct, 1),
               "families": {k: {"launches": a[0], "us": a[1] / 1e3, "tensor_pipe_pct": a[2] / max(a[1], 1), "dram_mb": a[3] / 1e6}
                            for k, a in agg.items()}}, open(sys.argv[2], "w"), indent=1)
if len(sys.argv) > 4:
    # per-network split: the launch list of a single-lane plan is in program order, i.e. launch 2 + i is op i of the per-op
    # table (2 input-packing kernels first, the head kernel last); checked through the kernel family of every op
    ops = list(csv.DictReader(open(sys.argv[4]), delimiter="\t"))
    order = []
    for r in rows:
        if r[0] not in order:
            order.append(r[0])
    first = next(i for i, k in enumerate(order) if "pack_input" not in names[k])
    fam = {"tile": "conv_gemm_kernel", "persist": "conv_gemm_persistent", "halo": "conv_halo_kernel"}
    ok = len(order) - first - 1 == len(ops) and all(
        names[order[first + j]].startswith(fam.get(o["variant"], "")) for j, o in enumerate(ops) if o["kind"] == "conv")
    if ok:
        groups = collections.OrderedDict((g, [0.0, 0.0]) for g in ("rootnet_backbone (HRNet-w32)", "reg_backbone (ResNet-50)",
                                                                     "deconv head + final layer"))
        for j, o in enumerate(ops):
            m = per[order[first + j]]
            g = ("rootnet_backbone (HRNet-w32)" if o["name"].startswith("rootnet_backbone") else
                 "reg_backbone (ResNet-50)" if o["name"].startswith("reg_backbone") else "deconv head + final layer")
            groups[g][0] += m.get(T, 0.0)
            groups[g][1] += m.get(T, 0.0) * m.get(P, 0.0)
        print("time-weighted tensor-pipe activity by network (launch order = program order of the single-lane plan):")
        for g, (t_, p_) in groups.items():
            print(f"  {g:32s} {t_ / 1e3:9.1f} us  {t_ / tot_t * 100:5.1f}% of the step   tensor pipe {p_ / max(t_, 1):5.1f} %")
    else:
        print("(launch list does not align with the per-op table: no per-network split)")
    alg = sum(float(r["bytes"]) for r in csv.DictReader(open(sys.argv[4]), delimiter="\t") if r["kind"] == "conv")
    json.dump({"robot": "kuka", "batch": 512, "kernels": len(per), "dram_bytes_per_step": tot_b,
               "algorithmic_bytes_per_step": alg, "source": sys.argv[5] if len(sys.argv) > 5 else sys.argv[1]},
              open(sys.argv[3], "w"))
