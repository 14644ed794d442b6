#!/bin/bash
mkdir -p gpurun_out
for gb in 64 128; do for ov in 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary --global-batch $gb --overlap $ov > gpurun_out/_b.json 2>/dev/null
  python - "$gb" "$ov" <<'PY'
import json,sys
d=json.load(open('gpurun_out/_b.json')); print('shard', sys.argv[1], 'overlap', sys.argv[2], round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'img/s  e2e', round(d['e2e']['value'],1))
PY
done; done
