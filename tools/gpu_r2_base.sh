#!/bin/bash
# Round-2 baseline visit (round-1 code): launch list of one 512-image bench step with the tensor-pipe counter,
# halo-kernel timelines / variants at the bench batch, --set full captures of the top kernels at batch 512.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 ncu --profile-from-start off \
  --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/r02_base_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
  --no-latency --no-other-configs --profile-range > gpurun_out/r02_base_ncu_launch.log 2>&1
wc -l gpurun_out/r02_base_launches.csv
for args in "32 64 512" "32 64 512 res" "64 32 512" "64 32 512 res"; do
  timeout 120 python tools/timeline_halo.py $args >> gpurun_out/r02_base_halo_timeline.txt 2>&1
  HRP_HALO_PAIR=0 timeout 120 python tools/bench_one_conv.py ${args/ 512/ 512 2} >> gpurun_out/r02_base_halo_nopair.txt 2>&1
  timeout 120 python tools/bench_one_conv.py ${args/ 512/ 512 2} >> gpurun_out/r02_base_halo_pair.txt 2>&1
done
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:conv_halo -s 3 -c 1 -o gpurun_out/r02_base_halo_32_b512_res -f python tools/prof_conv.py 0 512 res > gpurun_out/ncu_a.log 2>&1
timeout 300 $NCU -k regex:conv_halo -s 3 -c 1 -o gpurun_out/r02_base_halo_64_b512_res -f python tools/prof_conv.py 1 512 res > gpurun_out/ncu_b.log 2>&1
timeout 300 $NCU -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/r02_base_1x1_64to256_b512_res -f python tools/prof_conv.py 4 512 res > gpurun_out/ncu_c.log 2>&1
timeout 300 $NCU -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/r02_base_128x128k3_b512_res -f python tools/prof_conv.py 2 512 res > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out
