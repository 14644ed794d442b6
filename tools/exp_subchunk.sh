#!/bin/bash
# per-op times at small batches (L2-resident when repeated) vs 512, and a chunk sweep of the whole forward
mkdir -p gpurun_out
python tools/profile_model.py profile 32 > gpurun_out/profile_kuka32.txt 2>&1
python tools/profile_model.py profile 64 > gpurun_out/profile_kuka64.txt 2>&1
python tools/profile_model.py profile 128 > gpurun_out/profile_kuka128.txt 2>&1
HRP_SWEEP=32:1,64:1,128:1,128:2,256:1,256:2,512:1 python tools/profile_model.py sweep > gpurun_out/sweep.txt 2>&1
cat gpurun_out/sweep.txt; head -3 gpurun_out/profile_kuka32.txt
