#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.txt
bash tools/gpu_call18.sh
