#!/bin/bash
# halo kernel iteration: tests + microbench + timelines at the bench batch
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu -k "halo" 2>&1 | tail -15 ) > gpurun_out/halo_tests.txt
cat gpurun_out/halo_tests.txt
rm -f gpurun_out/halo_bench.txt gpurun_out/halo_timeline.txt
for args in "32 64 512 2" "32 64 512 2 res" "64 32 512 2" "64 32 512 2 res" "64 64 512 2" "64 64 512 2 res"; do
  timeout 120 python tools/bench_one_conv.py $args >> gpurun_out/halo_bench.txt 2>&1
done
for extra in "$@"; do
  for args in "32 64 512 2" "32 64 512 2 res" "64 32 512 2" "64 32 512 2 res"; do
    echo -n "[$extra] " >> gpurun_out/halo_bench.txt
    env $extra timeout 120 python tools/bench_one_conv.py $args >> gpurun_out/halo_bench.txt 2>&1
  done
done
cat gpurun_out/halo_bench.txt
for args in "32 64 512" "32 64 512 res" "64 32 512" "64 32 512 res"; do
  timeout 120 python tools/timeline_halo.py $args 2>&1 | grep -v "^launch [12]" >> gpurun_out/halo_timeline.txt
done
cat gpurun_out/halo_timeline.txt
