#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_eval_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/eval_tests.txt; cat gpurun_out/eval_tests.txt
