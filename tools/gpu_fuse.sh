#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/fuse_bench.txt
for w in up2 up2pre s2; do
  for cfg in "" "HRP_CONV_STAGED=0"; do
    echo "[$w $cfg]" >> gpurun_out/fuse_bench.txt
    env $cfg timeout 120 python tools/bench_fuse_conv.py $w 512 1 tl >> gpurun_out/fuse_bench.txt 2>&1
    env $cfg timeout 120 python tools/bench_fuse_conv.py $w 512 0 >> gpurun_out/fuse_bench.txt 2>&1
  done
done
cat gpurun_out/fuse_bench.txt
