#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_eval_gpu.py -m gpu -q --maxfail=4 2>&1 | tail -2
python tools/eval_timing.py
python tools/eval_timing.py
