#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "reference or fold or bitwise" 2>&1 | tail -3 )
for i in 1 2; do
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/_p.txt 2>&1
grep "final_layer" gpurun_out/per_op_kuka_512.tsv | cut -f1,11,12,13,14
done
