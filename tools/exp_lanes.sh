#!/bin/bash
python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== multi-lane"; HRP_SWEEP=512:1 python tools/profile_model.py sweep
echo "== single-lane"; HRP_SINGLE_LANE=1 HRP_SWEEP=512:1 python tools/profile_model.py sweep
python tools/profile_model.py profile 512 | head -24
