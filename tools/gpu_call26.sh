#!/bin/bash
mkdir -p gpurun_out
( time timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_eval_gpu.py -q -k "uint8_forward" -x ) 2>&1 | tail -12 > gpurun_out/sanitizer_model.txt
cat gpurun_out/sanitizer_model.txt
