#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run tc05     python -m pytest tests/test_tc05_gpu.py -m gpu -q --maxfail=4 -k "gemm or ffn or layernorm or tf32 or wgrad"
run scorer   python -m pytest tests/test_scorer_bf16_gpu.py tests/test_scorer_fp32_gpu.py tests/test_pipeline_gpu.py tests/test_train_gpu.py -m gpu -q --maxfail=4
run ffn_time python tools/ffn_timing.py
for i in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --e2e-batches 1 > gpurun_out/bench_dyn$i.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_dyn$i.log").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], "\n alone", d["kernel_ms_per_step"], "\n piped", d["kernel_ms_per_step_pipelined"])
PY
done
