#!/bin/bash
mkdir -p gpurun_out
{
HRP_SINGLE_LANE=1 python tools/exp_latency.py 512
HRP_SINGLE_LANE=0 python tools/exp_latency.py 512
} > gpurun_out/exp_lanes_512.txt 2>&1
cut -c1-110 gpurun_out/exp_lanes_512.txt
