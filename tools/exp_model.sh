#!/bin/bash
python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -2
HRP_SWEEP=${SWEEP:-512:1} python tools/profile_model.py sweep
python tools/profile_model.py profile 512 2>&1 | head -20
