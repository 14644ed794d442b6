#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -5 ) > gpurun_out/np_tests.txt
cat gpurun_out/np_tests.txt
rm -f gpurun_out/nprod_bench.txt gpurun_out/nprod_timelines.txt
for cfg in "X=1" "HRP_CONV_KSTAGE=1" "HRP_CONV_KSTAGE=2" "HRP_CONV_NPROD=1"; do
  echo "######## $cfg" >> gpurun_out/nprod_bench.txt
  env $cfg timeout 300 python tools/bench_conv.py 512 2>&1 | cut -c1-100 >> gpurun_out/nprod_bench.txt
done
cat gpurun_out/nprod_bench.txt
for args in "s2fuse 512" "9 512 res" "9 512" "4 512 res" "12 512 res" "15 512" "10 512 skip=2"; do
  timeout 120 python tools/timeline_persist.py $args 2>&1 | grep -E "^==|^tile period" >> gpurun_out/nprod_timelines.txt
done
cut -c1-330 gpurun_out/nprod_timelines.txt
( timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "reference or simt or fold or bitwise" 2>&1 | tail -5 )
timeout 600 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
PY
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/profile_kuka512.txt 2>&1
grep "final_layer" gpurun_out/per_op_kuka_512.tsv | cut -c1-140
