#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_conv_gpu.py -q -m gpu -x 2>&1 | tail -3 )
rm -f gpurun_out/nprod_bench.txt
for cfg in "HRP_CONV_TILE_NPROD=1" "X=1"; do
  echo "######## $cfg" >> gpurun_out/nprod_bench.txt
  env $cfg timeout 300 python tools/bench_conv.py 512 2>&1 | cut -c1-100 >> gpurun_out/nprod_bench.txt
done
( timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "reference or simt or fold or bitwise" 2>&1 | tail -3 )
for cfg in "HRP_CONV_TILE_NPROD=1" "X=1"; do
  env $cfg timeout 300 python tools/profile_model.py profile 512 > gpurun_out/_p.txt 2>&1
  python - "$cfg" <<'PY'
import csv,sys
rows=list(csv.DictReader(open('gpurun_out/per_op_kuka_512.tsv'),delimiter='\t'))
tot=sum(float(r['us']) for r in rows)
tile=sum(float(r['us']) for r in rows if r['variant']=='tile')
print(sys.argv[1], f"sum of ops {tot:.0f} us, tile-kernel ops {tile:.0f} us")
PY
done
timeout 400 python bench.py --no-cpu-baseline --no-latency --no-other-configs --no-secondary | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks'])"
