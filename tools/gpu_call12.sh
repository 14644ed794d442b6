#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_eval.py > gpurun_out/bench_eval.txt 2>&1; cat gpurun_out/bench_eval.txt
