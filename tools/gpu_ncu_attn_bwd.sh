#!/bin/bash
# GPU box: ncu --set full of the attention backward kernel inside a training step (run after train_timing exits 0).
set -e
mkdir -p gpurun_out
timeout 300 python tools/train_timing.py 2048,2048,2048,2048,2048,2048,2048,2048 > gpurun_out/train_plain.log 2>&1 && \
VSUM_TRAIN_PRECISION=bf16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc05 -s 4 -c 1 \
    -o gpurun_out/r01_attn_bwd_full -f python tools/train_timing.py 2048,2048,2048,2048,2048,2048,2048,2048 > gpurun_out/ncu_bwd.log 2>&1
tail -2 gpurun_out/ncu_bwd.log
