#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/model_parity.txt
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/r02_pytest_gpu.txt 2>&1
tail -4 $O/r02_pytest_gpu.txt
cp $O/model_parity.txt $O/r02_parity.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; tail -2 $O/r02_smoke.txt
timeout 900 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
cut -c1-300 $O/r02_bench.json; tail -2 $O/r02_bench.err
