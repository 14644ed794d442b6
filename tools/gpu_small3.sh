#!/bin/bash
# steps in flight (plan replicas on caller streams) x lanes / single lane + PDL, small shards
mkdir -p gpurun_out
out=gpurun_out/small_shards3.txt
rm -f $out
run() {
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary "$@" > gpurun_out/_b.json 2> gpurun_out/_b.err
  python - "$label" "$@" <<'PY' >> gpurun_out/small_shards3.txt
import json,sys
try:
    d=json.load(open('gpurun_out/_b.json'))
    print("%-20s %-34s %8.3f ms  %9.1f img/s  e2e %9.1f  clk %s" % (sys.argv[1], ' '.join(sys.argv[2:]), d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/_b.err').read()[-400:])
PY
}
for gb in 16 32 64 128 256; do
  for ov in 1 2 3 4; do
    run "lanes" X=1 -- --global-batch $gb --overlap $ov
    run "single" HRP_SINGLE_LANE=1 -- --global-batch $gb --overlap $ov
    run "single+PDL" HRP_SINGLE_LANE=1 HRP_PDL=1 -- --global-batch $gb --overlap $ov
  done
done
cat $out
