#!/bin/bash
# ncu --set full of the dominant kernel (two-tile attention forward, fast pass) on the DEFAULT bench workload (256 videos)
# -> traffic / pipe numbers for profiles/ and for bench.py's roofline.traffic; plus the launch list of one bench run.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn2_tc05 -s 8 -c 1 -o gpurun_out/prof_attn_full -f $CMD > gpurun_out/ncu_attn_full.log 2>&1
echo "attn rc $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launches rc $?"
tail -n 1 gpurun_out/plain.log | cut -c1-300
