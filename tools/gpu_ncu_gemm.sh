#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --e2e-batches 1"
ncu --set full --clock-control none -k regex:"gemm_tc05|ffn_tc05" -s 21 -c 7 -o gpurun_out/r02c_gemm_full -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "rc $?"
