#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_head_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "head or reference or fold or bitwise or integral or fused" 2>&1 | tail -3 )
timeout 600 python bench.py --no-cpu-baseline --no-other-configs --no-secondary > gpurun_out/_b.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/_b.json')); print('bench', round(d['ms_per_step'],3), round(d['value'],1), 'lat p50', d['latency_b1']['p50_ms'], 'head', round(d['roofline_head']['achieved'],1), d['roofline_head']['note'][-22:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:head_from_partials -c 2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-other-configs --no-secondary 2>&1 | grep -E "gpu__time_duration|head_from" | head -6
