#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -5 ) > gpurun_out/np_tests.txt
cat gpurun_out/np_tests.txt
out=gpurun_out/nprod_timelines.txt
rm -f $out
for cfg in "HRP_CONV_NPROD=1" "HRP_CONV_NPROD=2" "HRP_CONV_NPROD=1 HRP_CONV_KSTAGE=2" "HRP_CONV_NPROD=2 HRP_CONV_KSTAGE=2"; do
  echo "######## $cfg" >> $out
  for args in "s2fuse 512" "9 512 res" "9 512" "4 512 res" "12 512 res" "15 512"; do
    env $cfg timeout 120 python tools/timeline_persist.py $args 2>&1 | grep -E "^==|^tile period" >> $out
  done
  echo "######## $cfg" >> gpurun_out/nprod_bench.txt
  env $cfg timeout 300 python tools/bench_conv.py 512 >> gpurun_out/nprod_bench.txt 2>&1
done
cat $out | cut -c1-330
cat gpurun_out/nprod_bench.txt
