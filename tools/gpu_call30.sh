#!/bin/bash
mkdir -p gpurun_out
( time timeout 800 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 120 python tools/exp_head.py 2>&1 | tail -7 | tee gpurun_out/exp_head_final.txt
