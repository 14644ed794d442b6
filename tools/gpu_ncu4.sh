#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/r02_full_1x1_64to256_persist_b512_res -f python tools/prof_conv.py 4 512 res variant=1 > gpurun_out/ncu_h1.log 2>&1
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/r02_full_1x1_256to1024_persist_b512_res -f python tools/prof_conv.py 9 512 res variant=1 > gpurun_out/ncu_h2.log 2>&1
tail -1 gpurun_out/ncu_h1.log gpurun_out/ncu_h2.log
