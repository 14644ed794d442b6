#!/bin/bash
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_eval_gpu.py tests/test_head_gpu.py -q -k "pnp or crop_resize or metrics or fk or head" -x 2>&1 | tail -12 > gpurun_out/sanitizer_eval_head.txt
cat gpurun_out/sanitizer_eval_head.txt
