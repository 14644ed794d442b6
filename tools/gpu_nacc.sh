#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -3 )
( timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "reference or simt or fold or bitwise or single_lane" 2>&1 | tail -3 )
out=gpurun_out/nacc_model.txt
rm -f $out
for cfg in "HRP_CONV_NACC=2" "X=1" "HRP_CONV_NACC=2" "X=1"; do
  echo "######## $cfg" >> $out
  env $cfg timeout 300 python tools/profile_model.py profile 512 > gpurun_out/_p.txt 2>&1
  python - >> $out <<'PY'
import csv
rows=list(csv.DictReader(open('gpurun_out/per_op_kuka_512.tsv'),delimiter='\t'))
tot=sum(float(r['us']) for r in rows)
pers=sum(float(r['us']) for r in rows if r['variant']=='persist')
res=sum(float(r['us']) for r in rows if r['variant']=='persist' and r['epi']=='3')
gen=sum(float(r['us']) for r in rows if r['variant']=='persist' and r['epi'] in ('1','2'))
pl=sum(float(r['us']) for r in rows if r['variant']=='persist' and r['epi']=='0')
fin=[float(r['us']) for r in rows if r['name']=='final_layer'][0]
print(f"sum of ops {tot:.0f} us, persistent-kernel ops {pers:.0f} us (plain {pl:.0f}, residual {res:.0f}, generic {gen:.0f}, final {fin:.0f})")
PY
done
cat $out
for cfg in "HRP_CONV_NACC=2" "X=1"; do
for gb in 512 64; do
  env $cfg timeout 400 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary --global-batch $gb > gpurun_out/_b.json 2>/dev/null
  python - "$cfg" "$gb" <<'PY'
import json,sys
d=json.load(open('gpurun_out/_b.json')); print(sys.argv[1], 'B', sys.argv[2], round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'img/s  e2e', round(d['e2e']['value'],1), 'clk', d['clocks']['sm_mhz'])
PY
done
done
