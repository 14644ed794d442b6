#!/bin/bash
# Round-2 final visit: full GPU test suite, smoke, bench (both arms), per-op profile, ncu launch list of one bench step
# (time + tensor-pipe + DRAM bytes per launch), ncu --set full of the top kernels at batch 512, rows f1-f3, sanitizer.
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/r02_smi.txt
rm -f $O/model_parity.txt
( time timeout 1800 python -m pytest tests -m gpu -q ) > $O/r02_pytest_gpu.txt 2>&1
tail -4 $O/r02_pytest_gpu.txt
cp $O/model_parity.txt $O/r02_parity.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; tail -3 $O/r02_smoke.txt
timeout 900 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
cat $O/r02_bench.json | cut -c1-600; tail -3 $O/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
cat $O/r02_bench_reference.json | cut -c1-400
timeout 300 python tools/profile_model.py profile 512 > $O/r02_profile_kuka512.txt 2>&1
cp $O/per_op_kuka_512.tsv $O/r02_per_op_kuka512.tsv
timeout 1500 ncu --profile-from-start off \
  --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $O/r02_ncu_launches_b512.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
  --no-latency --no-other-configs --no-secondary --profile-range > $O/r02_ncu_launch.log 2>&1
wc -l $O/r02_ncu_launches_b512.csv
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 600 $NCU -k "regex:persistent<.int.64, .int.4>" -c 1 -o $O/r02_full_final_fold_b512 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-other-configs --no-secondary > $O/ncu_f1.log 2>&1
timeout 300 $NCU -k regex:conv_halo -s 3 -c 1 -o $O/r02_full_halo_pair_32_b512_res -f python tools/prof_conv.py 0 512 res > $O/ncu_f2.log 2>&1
timeout 300 $NCU -k regex:conv_halo -s 3 -c 1 -o $O/r02_full_halo_64_b512_res -f python tools/prof_conv.py 1 512 res > $O/ncu_f3.log 2>&1
timeout 300 $NCU -k regex:conv_gemm_kernel -s 3 -c 1 -o $O/r02_full_deconv_256_tile_b512 -f python tools/prof_conv.py 14 512 > $O/ncu_f4.log 2>&1
timeout 300 $NCU -k regex:conv_gemm -s 3 -c 1 -o $O/r02_full_1x1_256to1024_b512_res -f python tools/prof_conv.py 9 512 res > $O/ncu_f5.log 2>&1
timeout 300 $NCU -k regex:head_kernel -s 3 -c 1 -o $O/r02_full_head_kuka512 -f python tools/prof_head.py kuka 512 > $O/ncu_f6.log 2>&1
timeout 300 python tools/bench_eval.py > $O/r02_bench_eval_rows_f1_f3.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "test_full_forward_vs_reference_and_oracle and panda" > $O/r02_sanitizer.txt 2>&1
tail -5 $O/r02_sanitizer.txt
ls -la $O | grep r02_
