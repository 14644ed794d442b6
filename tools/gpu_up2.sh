#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/up2_now.txt
rm -f $out
for cfg in "X=1" "HRP_CONV_STAGED=1" "HRP_CONV_STAGED=0"; do
  for w in up2 up2pre; do
    for v in 0 1; do
      echo -n "[$cfg] " >> $out
      env $cfg timeout 120 python tools/bench_fuse_conv.py $w 512 $v 2>&1 | tail -1 >> $out
    done
  done
done
cat $out
