#!/bin/bash
# batch-1 (and batch-8) latency under the launch-structure toggles
mkdir -p gpurun_out
{
python tools/exp_latency.py 1
HRP_PDL=1 python tools/exp_latency.py 1
HRP_SINGLE_LANE=1 python tools/exp_latency.py 1
HRP_SINGLE_LANE=1 HRP_PDL=1 python tools/exp_latency.py 1
HRP_AUTOTUNE=0 python tools/exp_latency.py 1
python tools/exp_latency.py 8
HRP_PDL=1 python tools/exp_latency.py 8
} > gpurun_out/exp_latency.txt 2>&1
cat gpurun_out/exp_latency.txt
