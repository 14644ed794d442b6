#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "conv or reference or simt or fold or bitwise or persistent" 2>&1 | tail -3 )
timeout 600 python bench.py --no-cpu-baseline --no-latency --no-secondary > gpurun_out/bench_old_table.json 2> gpurun_out/bench_iter.err
timeout 900 python tools/make_tuning.py gpurun_out/tuning_b200.txt > gpurun_out/make_tuning.log 2>&1
tail -2 gpurun_out/make_tuning.log
HRP_TUNING=gpurun_out/tuning_b200.txt timeout 600 python bench.py --no-cpu-baseline --no-latency --no-secondary > gpurun_out/bench_new_table.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
for f in ('old','new'):
    d=json.load(open(f'gpurun_out/bench_{f}_table.json'))
    print(f, 'bench', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], {k:round(v['images_per_sec'],1) for k,v in d['other_configs'].items()})
PY
for gb in 64 128 256; do
  for t in "" "gpurun_out/tuning_b200.txt"; do
    HRP_TUNING=$t timeout 300 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary --global-batch $gb > gpurun_out/_b.json 2>/dev/null
    python - "$gb" "$t" <<'PY'
import json,sys
d=json.load(open('gpurun_out/_b.json')); print('shard', sys.argv[1], 'table', sys.argv[2] or 'committed', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'img/s  e2e', round(d['e2e']['value'],1))
PY
  done
done
