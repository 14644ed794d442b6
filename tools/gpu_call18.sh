#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -4 gpurun_out/bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[0])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'head',round(d['roofline_head']['frac'],3),'lat',d['latency_b1']['p50_ms'])
print(d['other_configs'])
print(d['roofline_head'].get('traffic'), d['roofline'].get('traffic'))
P
