#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -4 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'head',d['roofline_head']['frac'],'lat',d['latency_b1'],'clk',d['clocks'])
print('layers sum_us',d['roofline_layers']['sum_us'])
P
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/profile_kuka512.txt 2>&1
