#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/$name.log | cut -c1-300; }
run a3_bringup python tools/attn2_bringup.py
run a3_tests   python -m pytest tests/test_tc05_gpu.py -m gpu -q -k "prescaled" --maxfail=5
VSUM_ATTN_KERNEL=3 run a3_bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --e2e-batches 1
