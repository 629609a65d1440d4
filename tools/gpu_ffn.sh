#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
run ffn_test   python -m pytest tests/test_tc05_gpu.py -m gpu -q -k "fused_ffn" --maxfail=3
run ffn_time   python tools/ffn_timing.py
run scorer     python -m pytest tests/test_scorer_bf16_gpu.py tests/test_pipeline_gpu.py -m gpu -q --maxfail=4
run bench      python bench.py --steps 10 --warmup 3 --no-cpu-baseline
