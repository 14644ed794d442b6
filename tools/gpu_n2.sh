#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json; tail -4 gpurun_out/bench_n2.err
( time timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
cat gpurun_out/bench_ref_n2.json; tail -4 gpurun_out/bench_ref_n2.err
