#!/bin/bash
# small-shard experiments (the per-GPU batch of the strong-scaling bench at N = 8 / 4 / 2)
mkdir -p gpurun_out
out=gpurun_out/small_shards.txt
rm -f $out
run() {  # label, env..., -- args
  local label="$1"; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary "$@" > gpurun_out/_b.json 2> gpurun_out/_b.err
  python - "$label" "$@" <<'PY' >> gpurun_out/small_shards.txt
import json,sys
try:
    d=json.load(open('gpurun_out/_b.json'))
    print("%-44s %-28s %8.3f ms  %9.1f img/s  e2e %9.1f  clk %s" % (sys.argv[1], ' '.join(sys.argv[2:]), d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/_b.err').read()[-400:])
PY
}
for gb in 64 128 256; do
  run "default" X=1 -- --global-batch $gb
  run "single lane" HRP_SINGLE_LANE=1 -- --global-batch $gb
  run "single lane + PDL" HRP_SINGLE_LANE=1 HRP_PDL=1 -- --global-batch $gb
  run "heuristics (no table)" HRP_TUNING=0 -- --global-batch $gb
  run "all tile kernels" HRP_CONV_VARIANT=tile -- --global-batch $gb
  run "overlap 2" X=1 -- --global-batch $gb --overlap 2
done
cat $out
timeout 300 python tools/profile_model.py profile 64 > gpurun_out/profile_kuka64.txt 2>&1
tail -5 gpurun_out/profile_kuka64.txt
