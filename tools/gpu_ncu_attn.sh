#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --videos 32 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_tc05 -s 4 -c 1 -o gpurun_out/prof_attn3 $CMD > gpurun_out/ncu_attn3.log 2>&1
echo "attn rc $?"
