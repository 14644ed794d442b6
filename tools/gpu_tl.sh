#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 4; do
HRP_CONV_DBG=$d timeout 300 python tools/profile_model.py profile 512 > gpurun_out/profile_kuka512.txt 2>&1
echo -n "dbg=$d "; grep "final_layer" gpurun_out/per_op_kuka_512.tsv | cut -c1-140
done
