#!/bin/bash
# One GPU-box visit of round 2: parity tests, bench line, per-op profile, (optional) tuning table.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
rm -f gpurun_out/model_parity.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -15 gpurun_out/pytest_gpu.txt
if [ "$1" == "tune" ]; then
  timeout 900 python tools/make_tuning.py gpurun_out/tuning_b200.txt > gpurun_out/make_tuning.log 2>&1
  tail -3 gpurun_out/make_tuning.log
fi
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/profile_kuka512.txt 2>&1
