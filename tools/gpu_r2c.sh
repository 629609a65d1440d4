#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/$name.log | cut -c1-500; }
run eval_tests python -m pytest tests/test_eval_gpu.py tests/test_pipeline_gpu.py -m gpu -q --maxfail=4
run eval_plain python tools/eval_timing.py
run bench      python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train
ncu --set full --clock-control none -k regex:"knapsack" -s 12 -c 4 -o gpurun_out/r02_knap_after -f python tools/eval_timing.py > gpurun_out/ncu_knap.log 2>&1
echo "ncu rc $?"
