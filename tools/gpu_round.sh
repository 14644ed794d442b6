#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-op profile, ncu launch list, ncu full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/profile_kuka512.txt 2>&1
timeout 300 python tools/bench_conv.py 256 > gpurun_out/bench_conv256.txt 2>&1
# launch list of the bench command itself (one timed step of the 512-image workload), with DRAM traffic per launch
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
  --no-latency --no-other-configs --profile-range > gpurun_out/ncu_launch_bench.log 2>&1
wc -l gpurun_out/launches.csv
