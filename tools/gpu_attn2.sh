#!/bin/bash
# GPU box: bring-up of the two-tile attention kernel -- quick correctness first (bounded), then all variants, then the unit tests.
mkdir -p gpurun_out
echo "=== quick"; timeout 300 python tools/attn2_bringup.py --quick --libs video-summarization_b200/vsum_b200/libvsum_b200.so > gpurun_out/a2_quick.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/a2_quick.log
if grep -q "TFLOP" gpurun_out/a2_quick.log && ! grep -q "MISMATCH\|Error\|error" gpurun_out/a2_quick.log; then
  echo "=== timing stamps"; timeout 300 python tools/attn2_one.py --lib tools/variants/libvsum_timing.so --reps 1 > gpurun_out/a2_timing.log 2>&1; echo "exit $?"; tail -n 24 gpurun_out/a2_timing.log
  echo "=== variants"; timeout 1200 python tools/attn2_bringup.py > gpurun_out/a2_variants.log 2>&1; echo "exit $?"; tail -n 20 gpurun_out/a2_variants.log
  echo "=== tc05 tests"; timeout 900 python -m pytest tests/test_tc05_gpu.py -m gpu -q --maxfail=6 > gpurun_out/a2_tc05.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/a2_tc05.log
  echo "=== bench v2"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a2_bench_v2.log 2>&1; echo "exit $?"; tail -n 2 gpurun_out/a2_bench_v2.log | cut -c1-400
fi
