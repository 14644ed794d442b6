#!/bin/bash
# lanes vs single lane + PDL, small-M N-tile splitting, over shard sizes
mkdir -p gpurun_out
out=gpurun_out/small_shards2.txt
rm -f $out
run() {
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary "$@" > gpurun_out/_b.json 2> gpurun_out/_b.err
  python - "$label" "$@" <<'PY' >> gpurun_out/small_shards2.txt
import json,sys
try:
    d=json.load(open('gpurun_out/_b.json'))
    print("%-36s %-22s %8.3f ms  %9.1f img/s  e2e %9.1f  clk %s" % (sys.argv[1], ' '.join(sys.argv[2:]), d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/_b.err').read()[-400:])
PY
}
for gb in 1 2 4 8 16 32 64 128 256; do
  run "lanes" X=1 -- --global-batch $gb
  run "single+PDL" HRP_SINGLE_LANE=1 HRP_PDL=1 -- --global-batch $gb
  run "lanes, smallm2" HRP_CONV_SMALLM=2 -- --global-batch $gb
  run "single+PDL, smallm2" HRP_SINGLE_LANE=1 HRP_PDL=1 HRP_CONV_SMALLM=2 -- --global-batch $gb
done
cat $out
HRP_CONV_SMALLM=2 timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "golden or reference or simt or fold" 2>&1 | tail -5
