#!/bin/bash
mkdir -p gpurun_out
{
python tools/exp_latency.py 1
HRP_CONV_SMALLM=0 python tools/exp_latency.py 1
python tools/exp_latency.py 4
HRP_CONV_SMALLM=0 python tools/exp_latency.py 4
python tools/exp_latency.py 16
HRP_CONV_SMALLM=0 python tools/exp_latency.py 16
} > gpurun_out/exp_smallm.txt 2>&1
cat gpurun_out/exp_smallm.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt
