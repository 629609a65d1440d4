#!/bin/bash
# Quick GPU check after a kernel change: tcgen05 unit tests, eval + scorer parity, pipeline, bench.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/$name.log; }
run tc05     python -m pytest tests/test_tc05_gpu.py -m gpu -q --maxfail=4
run eval     python -m pytest tests/test_eval_gpu.py -m gpu -q --maxfail=4
run bf16     python -m pytest tests/test_scorer_bf16_gpu.py tests/test_pipeline_gpu.py tests/test_scorer_fp32_gpu.py -m gpu -q --maxfail=4
run bench    python bench.py --steps 10 --warmup 3 --no-cpu-baseline
