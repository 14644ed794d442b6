#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.txt
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/profile_kuka512.txt 2>&1
grep -P "^fuse_add|^maxpool" gpurun_out/per_op_kuka_512.tsv | cut -f1,13
