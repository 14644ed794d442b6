#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n8_gpus.txt
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-latency ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep '^{' gpurun_out/bench_n$N.json | cut -c1-300; tail -4 gpurun_out/bench_n$N.err
done
