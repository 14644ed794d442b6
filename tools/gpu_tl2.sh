#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/persist_timelines.txt
rm -f $out
for args in "s2fuse 512" "9 512 res" "9 512" "2 512 res" "2 512" "4 512 res" "10 512" "12 512 res" "18 512" "15 512"; do
  timeout 120 python tools/timeline_persist.py $args >> $out 2>&1
done
cat $out
