#!/bin/bash
# ncu --set full of the dominant kernels on the DEFAULT bench workload (256 videos) -> traffic numbers
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_tc05 -s 4 -c 1 -o gpurun_out/prof_attn_full $CMD > gpurun_out/ncu_attn_full.log 2>&1
echo "attn rc $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 26 -c 5 -o gpurun_out/prof_gemm_full $CMD > gpurun_out/ncu_gemm_full.log 2>&1
echo "gemm rc $?"
