#!/bin/bash
# Ablation of the halo kernel (what bounds a tile?): time with pieces of the epilogue removed
for dbg in 0 1 2 3 4 5 7; do
  for res in "" res; do
    echo -n "dbg=$dbg $res: "
    HRP_HALO_DBG=$dbg python tools/bench_one_conv.py 32 64 256 2 $res | tail -1
  done
done
for dbg in 0 1 2 7; do
  for res in "" res; do
    echo -n "C64 dbg=$dbg $res: "
    HRP_HALO_DBG=$dbg python tools/bench_one_conv.py 64 32 256 2 $res | tail -1
  done
done
