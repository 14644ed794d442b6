#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/kstage_model.txt
rm -f $out
for cfg in "HRP_CONV_KSTAGE=1 HRP_CONV_NPROD=1" "HRP_CONV_KSTAGE=1" "HRP_CONV_KSTAGE=2 HRP_CONV_NPROD=1" "HRP_CONV_KSTAGE=2" "X=1"; do
  echo "######## $cfg" >> $out
  env $cfg timeout 300 python tools/profile_model.py profile 512 > gpurun_out/_p.txt 2>&1
  python - >> $out <<'PY'
import csv
rows=list(csv.DictReader(open('gpurun_out/per_op_kuka_512.tsv'),delimiter='\t'))
tot=sum(float(r['us']) for r in rows)
pers=sum(float(r['us']) for r in rows if r['variant']=='persist')
print(f"sum of ops {tot:.0f} us, persistent-kernel ops {pers:.0f} us")
for r in rows:
    if r['name'] in ('final_layer','reg_backbone.layer3.1.conv3','reg_backbone.layer3.1.conv1','reg_backbone.layer4.1.conv3','reg_backbone.layer2.1.conv3','rootnet_backbone.stage3.0.fuse_layers.1.0.0.0','rootnet_backbone.stage3.0.branches.2.0.conv1','reg_backbone.layer1.1.conv3','reg_backbone.layer1.1.conv1'):
        print(f"  {r['name']:48s} {r['variant']:8s} stages {r['stages']:2s} {float(r['us']):8.1f} us")
PY
done
cat $out
