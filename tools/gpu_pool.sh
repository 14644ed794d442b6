#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -3 )
( timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "reference or simt or fold or bitwise or depthnet_vs" 2>&1 | tail -3 )
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/_p.txt 2>&1
grep -P "final_feat_layer|layer4.2.conv3|maxpool" gpurun_out/per_op_kuka_512.tsv | cut -f1,11,13
timeout 300 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary --global-batch 16 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks'])"
timeout 400 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'])"
