#!/bin/bash
python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "halo" 2>&1 | tail -5
for res in "" res; do
  echo -n "pair $res: "; python tools/bench_one_conv.py 32 64 256 2 $res | tail -1
  echo -n "plain $res: "; HRP_HALO_PAIR=0 python tools/bench_one_conv.py 32 64 256 2 $res | tail -1
  echo -n "C64 $res: "; python tools/bench_one_conv.py 64 32 256 2 $res | tail -1
  echo -n "C64@64 $res: "; python tools/bench_one_conv.py 64 64 256 2 $res | tail -1
done
python tools/timeline_halo.py 32 64 256 2>&1 | head -4
python tools/timeline_halo.py 32 64 256 res 2>&1 | head -4
