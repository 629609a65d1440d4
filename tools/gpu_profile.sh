#!/bin/bash
# One gpurun call: default bench (both arms), then the ncu launch list and full captures of the
# dominant kernels on a small configuration of the same command.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"
tail -c 3000 gpurun_out/bench_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc $?"
CMD="python bench.py --videos 16 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_tc05 -s 4 -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn rc $?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 25 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc $?"
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"knapsack|overlap_kernel|shot_mean" -s 3 -c 3 -o gpurun_out/prof_eval $CMD > gpurun_out/ncu_eval.log 2>&1
echo "eval rc $?"
ls -la gpurun_out
