#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/halo_ablation.txt
rm -f $out
for dbg in 0 1 2 3; do
  for args in "32 64 512 2" "32 64 512 2 res" "64 32 512 2" "64 32 512 2 res"; do
    echo -n "[HRP_HALO_DBG=$dbg] " >> $out
    HRP_HALO_DBG=$dbg timeout 120 python tools/bench_one_conv.py $args 2>&1 | tail -1 >> $out
  done
done
cat $out
