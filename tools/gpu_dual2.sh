#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -3 )
rm -f gpurun_out/nprod_bench.txt
for cfg in "HRP_CONV_DUAL=0" "X=1"; do
  echo "######## $cfg" >> gpurun_out/nprod_bench.txt
  env $cfg timeout 300 python tools/bench_conv.py 512 2>&1 | cut -c1-100 >> gpurun_out/nprod_bench.txt
done
cat gpurun_out/nprod_bench.txt
