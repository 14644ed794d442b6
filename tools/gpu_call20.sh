#!/bin/bash
mkdir -p gpurun_out
{
for b in 32 64 128 256; do
HRP_SINGLE_LANE=1 python tools/exp_latency.py $b
HRP_SINGLE_LANE=0 python tools/exp_latency.py $b
done
} > gpurun_out/exp_lanes_midbatch.txt 2>&1
cut -c1-110 gpurun_out/exp_lanes_midbatch.txt
