#!/bin/bash
# One gpurun call: every GPU test file in its own process (a trapped kernel poisons its CUDA
# context), smoke, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n 6 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run eval        python -m pytest tests/test_eval_gpu.py -m gpu -q --maxfail=10
run fp32        python -m pytest tests/test_scorer_fp32_gpu.py -m gpu -q --maxfail=10
run tc_gemm     python -m pytest tests/test_tc05_gpu.py -m gpu -q -k "gemm_bf16" --maxfail=4
run tc_tf32     python -m pytest tests/test_tc05_gpu.py -m gpu -q -k "tf32" --maxfail=4
run tc_ln       python -m pytest tests/test_tc05_gpu.py -m gpu -q -k "layernorm" --maxfail=4
run tc_attn     python -m pytest tests/test_tc05_gpu.py -m gpu -q -k "attention" --maxfail=4
if ! grep -q "passed" gpurun_out/tc_attn.log || grep -q "failed" gpurun_out/tc_attn.log; then run diag_attn python tools/diag_tc05.py; fi
run bf16        python -m pytest tests/test_scorer_bf16_gpu.py -m gpu -q --maxfail=10
run pipeline    python -m pytest tests/test_pipeline_gpu.py -m gpu -q --maxfail=10
run smoke       python -c "import __graft_entry__ as g; g.smoke()"
run bench_small python bench.py --videos 64 --steps 3 --warmup 3 --no-cpu-baseline
echo DONE | tee -a gpurun_out/summary.txt
