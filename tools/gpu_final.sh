#!/bin/bash
# Round-end visit: tests, bench, per-op profile, conv microbench, ncu launch list (+DRAM traffic), ncu --set full of the
# head and the dominant conv kernels, rows f1-f3 microbench.
bash tools/gpu_round.sh
NCU="ncu --set full --clock-control none --import-source on"
timeout 240 $NCU -k regex:head_kernel -s 3 -c 1 -o gpurun_out/r01_head_kuka512 -f python tools/prof_head.py kuka 512 > gpurun_out/ncu4.log 2>&1
timeout 240 $NCU -k regex:conv_halo -s 3 -c 1 -o gpurun_out/r01_conv_halo_32x32k3_64 -f python tools/prof_conv.py 0 256 res > gpurun_out/ncu2.log 2>&1
timeout 240 $NCU -k regex:conv_gemm_kernel -s 3 -c 1 -o gpurun_out/r01_conv_tile_256x256k3_16 -f python tools/prof_conv.py 8 256 > gpurun_out/ncu1.log 2>&1
timeout 300 python tools/bench_eval.py > gpurun_out/bench_eval.txt 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
ls -la gpurun_out | head -40
