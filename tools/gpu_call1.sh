#!/bin/bash
# Session-5 visit 1: verify HEAD (tests, bench, per-op profile, launch list) + ncu --set full of the dominant kernels.
bash tools/gpu_round.sh
NCU="ncu --set full --clock-control none --import-source on"
timeout 240 $NCU -k regex:conv_halo -s 3 -c 1 -o gpurun_out/r01_conv_halo_32x32k3_64 -f python tools/prof_conv.py 0 256 res > gpurun_out/ncu2.log 2>&1
timeout 240 $NCU -k regex:conv_gemm -s 3 -c 1 -o gpurun_out/r01_conv_1x1_64to256_res -f python tools/prof_conv.py 4 256 res > gpurun_out/ncu3.log 2>&1
timeout 240 $NCU -k regex:conv_gemm_kernel -s 3 -c 1 -o gpurun_out/r01_conv_tile_256x256k3_16 -f python tools/prof_conv.py 8 256 > gpurun_out/ncu1.log 2>&1
timeout 240 $NCU -k regex:head_kernel -s 3 -c 1 -o gpurun_out/r01_head_kuka512 -f python tools/prof_head.py kuka 512 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out
