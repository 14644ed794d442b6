mkdir -p gpurun_out
timeout 600 python tools/probe_tma.py > gpurun_out/probe_tma2.txt 2>&1
for pr in 0 128; do echo "== HRP_TMA_L2PROMO=$pr" >> gpurun_out/probe_tma2.txt; HRP_TMA_L2PROMO=$pr timeout 600 python tools/probe_tma.py 2>&1 | sed -n 3,12p >> gpurun_out/probe_tma2.txt; done
cat gpurun_out/probe_tma2.txt
