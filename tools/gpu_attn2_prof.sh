#!/bin/bash
mkdir -p gpurun_out
echo "=== timing stamps"; timeout 300 python tools/attn2_one.py --lib tools/variants/libvsum_timing.so --reps 1 > gpurun_out/a2_timing.log 2>&1; echo "exit $?"; cat gpurun_out/a2_timing.log | tail -n 30
echo "=== ncu"
python tools/attn2_one.py > gpurun_out/a2_one_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn2_tc05 -s 1 -c 1 -o gpurun_out/a2_prof -f python tools/attn2_one.py > gpurun_out/a2_ncu.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/a2_one_plain.log gpurun_out/a2_ncu.log
