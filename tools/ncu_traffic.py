"""Summarise the ncu launch list of one bench step (tools/gpu_round.sh): per-kernel time share and DRAM traffic.
Writes profiles/r01_ncu_traffic.json (read by bench.py for roofline.traffic) and prints the share table."""
import collections
import csv
import json
import re
import sys
from pathlib import Path

src = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv")
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
per = collections.defaultdict(dict)
names = {}
for r in rows:
    per[r[0]][r[-3]] = float(r[-1]) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[-2], 1.0)
    names[r[0]] = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("hrp::", "")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for i, m in per.items():
    a = agg[names[i]]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot_t = sum(a[1] for a in agg.values())
tot_b = sum(a[2] for a in agg.values())
print(f"{len(per)} launches, {tot_t / 1e6:.2f} ms serialised under ncu, {tot_b / 1e9:.2f} GB of DRAM traffic")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:60s} {a[0]:4d} launches {a[1] / 1e3:9.1f} us {a[1] / tot_t * 100:5.1f}%  {a[2] / 1e6:9.1f} MB  "
          f"{a[2] / max(a[1], 1) :6.2f} GB/ms")
alg = None
tsv = Path(f"gpurun_out/per_op_kuka_{batch}.tsv")
if tsv.exists():
    alg = sum(float(r["bytes"]) for r in csv.DictReader(open(tsv), delimiter="\t") if r["kind"] == "conv")
out = {"robot": "kuka", "batch": batch, "kernels": len(per), "dram_bytes_per_step": tot_b,
       "algorithmic_bytes_per_step": alg, "source": "profiles/r01_ncu_launches_bench.csv"}
Path("profiles/r01_ncu_traffic.json").write_text(json.dumps(out))
print(json.dumps(out))
