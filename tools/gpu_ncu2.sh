#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k "regex:persistent<.int.64, .int.4>" -c 1 -o gpurun_out/r02_final_fold_b512 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-other-configs > gpurun_out/ncu_f.log 2>&1
grep -c "PROF" gpurun_out/ncu_f.log; ls -la gpurun_out/r02_final_fold_b512.ncu-rep
rm -f gpurun_out/halo_timeline.txt
for args in "32 64 512" "32 64 512 res" "64 32 512" "64 32 512 res"; do
  timeout 120 python tools/timeline_halo.py $args 2>&1 | grep -v "^launch [12]" >> gpurun_out/halo_timeline.txt
done
cat gpurun_out/halo_timeline.txt
