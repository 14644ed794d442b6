#!/bin/bash
for cfg in "X=1" "HRP_PDL=1" "X=1" "HRP_PDL=1"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary > gpurun_out/_b.json 2>/dev/null
  python - "$cfg" <<'PY'
import json,sys
d=json.load(open('gpurun_out/_b.json')); print(sys.argv[1], round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'img/s  e2e', round(d['e2e']['value'],1), 'clk', d['clocks']['sm_mhz'])
PY
done
