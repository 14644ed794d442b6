#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_conv.py 512 > gpurun_out/bench_conv512.txt 2>&1
cat gpurun_out/bench_conv512.txt
HRP_CONV_WRES=0 timeout 600 python tools/bench_conv.py 512 > gpurun_out/bench_conv512_nowres.txt 2>&1
grep "final\|rn 256->1024\|rn 64->256\|rn 1024" gpurun_out/bench_conv512_nowres.txt
