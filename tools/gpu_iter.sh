#!/bin/bash
# iteration visit: tests + halo microbench + bench + per-op profile
mkdir -p gpurun_out
rm -f gpurun_out/model_parity.txt gpurun_out/halo_bench.txt
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/iter_tests.txt
tail -4 gpurun_out/iter_tests.txt
for cfg in "" "HRP_HALO_NRING=5" "HRP_HALO_NRING=6" "HRP_HALO_T=2 HRP_HALO_NRING=4" "HRP_HALO_T=2 HRP_HALO_NRING=3"; do
  for args in "32 64 512 2" "32 64 512 2 res" "64 32 512 2" "64 32 512 2 res"; do
    echo -n "[$cfg] " >> gpurun_out/halo_bench.txt
    env $cfg timeout 120 python tools/bench_one_conv.py $args 2>&1 | tail -1 >> gpurun_out/halo_bench.txt
  done
done
cat gpurun_out/halo_bench.txt
timeout 600 python bench.py --no-cpu-baseline --no-latency > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['other_configs'])
PY
tail -3 gpurun_out/bench_iter.err
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/profile_kuka512.txt 2>&1
grep "final_layer" gpurun_out/per_op_kuka_512.tsv | cut -c1-140
