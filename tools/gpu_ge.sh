#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -4 )
for v in 0 1; do timeout 120 python tools/bench_fuse_conv.py s2 512 $v 2>&1 | tail -1; done
for v in 0 1; do timeout 120 python tools/bench_fuse_conv.py up2 512 $v 2>&1 | tail -1; done
( timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "reference or simt or fold or bitwise" 2>&1 | tail -3 )
timeout 300 python tools/profile_model.py profile 512 > gpurun_out/_p.txt 2>&1
python - <<'PY'
import csv
rows=list(csv.DictReader(open('gpurun_out/per_op_kuka_512.tsv'),delimiter='\t'))
tot=sum(float(r['us']) for r in rows)
pers=sum(float(r['us']) for r in rows if r['variant']=='persist')
res=sum(float(r['us']) for r in rows if r['variant']=='persist' and r['epi']=='3')
gen=sum(float(r['us']) for r in rows if r['variant']=='persist' and r['epi'] in ('1','2'))
gent=sum(float(r['us']) for r in rows if r['variant']=='tile' and r['epi'] in ('1','2'))
print(f"sum of ops {tot:.0f} us, persistent-kernel ops {pers:.0f} us (residual epilogue {res:.0f}, generic epilogues {gen:.0f}; tile-kernel generic {gent:.0f})")
PY
timeout 600 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])
PY
