#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/halo_order.txt
rm -f $out
for ord in 0 1 0 1; do
  for args in "32 64 512 2" "32 64 512 2 res"; do
    echo -n "[HRP_HALO_PAIR_ORDER=$ord] " >> $out
    HRP_HALO_PAIR_ORDER=$ord timeout 60 python tools/bench_one_conv.py $args 2>&1 | tail -1 >> $out
  done
done
cat $out
( timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -x -k "halo" 2>&1 | tail -3 )
