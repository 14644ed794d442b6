#!/bin/bash
# iteration visit: full GPU tests + bench at the shard sizes of the strong-scaling curve
mkdir -p gpurun_out
rm -f gpurun_out/model_parity.txt
( timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/iter_tests.txt
tail -6 gpurun_out/iter_tests.txt
out=gpurun_out/shard_curve.txt
rm -f $out
run() {
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 400 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary "$@" > gpurun_out/_b.json 2> gpurun_out/_b.err
  python - "$label" "$@" <<'PY' >> gpurun_out/shard_curve.txt
import json,sys
try:
    d=json.load(open('gpurun_out/_b.json'))
    print("%-14s %-34s %8.3f ms  %9.1f img/s  e2e %9.1f  inflight %s clk %s" % (sys.argv[1], ' '.join(sys.argv[2:]), d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['steps_in_flight_per_gpu'], d['clocks']['sm_mhz']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/_b.err').read()[-400:])
PY
}
for gb in 64 128 256 512; do
  run "default" X=1 -- --global-batch $gb
done
run "ov2" X=1 -- --global-batch 512 --overlap 2
run "ov2+PDL" HRP_PDL=1 -- --global-batch 512 --overlap 2
run "ov1+PDL" HRP_PDL=1 -- --global-batch 512 --overlap 1
cat $out
