#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_head_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
timeout 200 python tools/exp_head.py > gpurun_out/exp_head_flat.txt 2>&1; cat gpurun_out/exp_head_flat.txt
