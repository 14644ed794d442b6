#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 300 $NCU -k regex:conv_gemm_persistent -s 5 -c 1 -o gpurun_out/r02_full_s2fuse_b512 -f python tools/bench_fuse_conv.py s2 512 1 > gpurun_out/ncu_g1.log 2>&1
tail -2 gpurun_out/ncu_g1.log
ls -la gpurun_out/r02_full_s2fuse_b512.ncu-rep
