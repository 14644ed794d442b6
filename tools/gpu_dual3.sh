#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -3 )
( timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "reference or simt or fold or bitwise" 2>&1 | tail -3 )
rm -f gpurun_out/nprod_bench.txt
for cfg in "HRP_CONV_DUAL=0 HRP_CONV_RES_STORE=0" "HRP_CONV_DUAL=0" "X=1"; do
  echo "######## $cfg (residual)" >> gpurun_out/nprod_bench.txt
  env $cfg timeout 300 python tools/bench_conv.py 512 res 2>&1 | cut -c1-100 >> gpurun_out/nprod_bench.txt
done
out=gpurun_out/kstage_model.txt
rm -f $out
for cfg in "HRP_CONV_DUAL=0 HRP_CONV_RES_STORE=0" "HRP_CONV_DUAL=0" "X=1" "HRP_CONV_DUAL=0 HRP_CONV_RES_STORE=0" "X=1"; do
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
print(f"sum of ops {tot:.0f} us, persistent-kernel ops {pers:.0f} us (plain {pl:.0f}, residual {res:.0f}, generic {gen:.0f})")
PY
done
cat $out
timeout 600 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])
PY
