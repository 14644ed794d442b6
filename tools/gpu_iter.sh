#!/bin/bash
# iteration visit: tests + bench + per-op profiles
mkdir -p gpurun_out
rm -f gpurun_out/model_parity.txt
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/iter_tests.txt
tail -6 gpurun_out/iter_tests.txt
timeout 600 python bench.py --no-cpu-baseline --no-latency > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['other_configs'])
PY
tail -3 gpurun_out/bench_iter.err
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/profile_kuka512.txt 2>&1
grep "final_layer" gpurun_out/per_op_kuka_512.tsv | cut -c1-140
timeout 300 python tools/profile_model.py profile 64 > gpurun_out/profile_kuka64.txt 2>&1
tail -25 gpurun_out/profile_kuka64.txt
