#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "reference or fold or bitwise" 2>&1 | tail -3 )
out=gpurun_out/fold_w16.txt
rm -f $out
for cfg in "X=1" "HRP_CONV_KSTAGE=2" "HRP_CONV_DBG=4" "X=1" "HRP_CONV_KSTAGE=2"; do
  echo "######## $cfg" >> $out
  env $cfg timeout 300 python tools/profile_model.py profile 512 > gpurun_out/_p.txt 2>&1
  grep "final_layer" gpurun_out/per_op_kuka_512.tsv | cut -f1,11,12,13,14 >> $out
done
cat $out
timeout 600 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])
PY
