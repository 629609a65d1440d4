#!/bin/bash
mkdir -p gpurun_out
echo "=== 148x2048 shipped"; timeout 120 python tools/attn2_one.py --check 2>&1 | tail -3
echo "=== timing stamps"; timeout 120 python tools/attn2_one.py --lib tools/variants/libvsum_timing.so --reps 1 2>&1 | tail -4 | cut -c1-300
for L in "[300]*150" "[8192,4000,77]" "[1]*700" "[129,128,127]*60"; do
  echo "=== verbose $L"; timeout 120 python tools/attn2_one.py --lib tools/variants/libvsum_verbose.so --lens "$L" --check --reps 1 > gpurun_out/dbg.log 2>&1; echo "exit $?"
  grep -o "block [0-9]*,0 thread [0-9]* bar [0-9]* parity [0-9]" gpurun_out/dbg.log | awk '{print "blk",$2,"warp",int($4/32),"bar",($6-197632)/8,"par",$8}' | sort | uniq -c | sort -k3,3n -k5,5n | head -30
  grep -v "mbarrier timeout" gpurun_out/dbg.log | tail -2
done
echo "=== variants"; timeout 600 python tools/attn2_bringup.py 2>&1 | grep -v "softmax warp" | tail -8
echo "=== tc05 tests"; timeout 900 python -m pytest tests/test_tc05_gpu.py -m gpu -q --maxfail=6 2>&1 | tail -4
