#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/fold_dbg.txt
rm -f $out
for cfg in "X=1" "HRP_CONV_DBG=4" "HRP_CONV_DBG=1" "HRP_CONV_DBG=4 HRP_CONV_DUAL=0" "HRP_CONV_DBG=4 HRP_CONV_KSTAGE=2"; do
  echo "######## $cfg" >> $out
  env $cfg timeout 300 python tools/profile_model.py profile 512 > gpurun_out/_p.txt 2>&1
  grep "final_layer" gpurun_out/per_op_kuka_512.tsv | cut -f1,11,12,13,14 >> $out
done
cat $out
