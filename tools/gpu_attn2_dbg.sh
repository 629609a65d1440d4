#!/bin/bash
mkdir -p gpurun_out
echo "=== timing stamps (prescaled)"; timeout 120 python tools/attn2_one.py --prescaled --lib tools/variants/libvsum_timing.so --reps 1 2>&1 | tail -3 | cut -c1-300
echo "=== variants (prescaled)"; timeout 600 python tools/attn2_bringup.py 2>&1 | grep -v "softmax warp" | tail -9
echo "=== tc05 tests"; timeout 900 python -m pytest tests/test_tc05_gpu.py -m gpu -q --maxfail=6 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a2_bench_v2.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/a2_bench_v2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['kernel_ms_per_step'])"
