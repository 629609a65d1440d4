#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 2 gpurun_out/$name.log | cut -c1-300; }
run train_tests python -m pytest tests/test_train_gpu.py tests/test_scorer_bf16_gpu.py tests/test_scorer_fp32_gpu.py tests/test_pipeline_gpu.py -m gpu -q --maxfail=4
for i in 1 2 3; do python tools/train_bench.py --config finetune --steps 50 --warmup 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('finetune ms', round(d['ms_per_step'],3), 'launches', d['gpu_launches_per_step'])"; done
python tools/train_bench.py --config pretrain --len 2048 --steps 30 --warmup 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pretrain ms', round(d['ms_per_step'],3))"
