#!/bin/bash
mkdir -p gpurun_out
for v in 2 3 2 3; do
  VSUM_ATTN_KERNEL=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --e2e-batches 1 > gpurun_out/ab_$v.log 2>&1
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$v.log").read().strip().splitlines()[-1])
print("kernel $v value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "attn TF", round(d["roofline"]["achieved"]), "alone", d["kernel_ms_per_step"]["attention"], "clk", d["clocks"]["sm_mhz"])
PY
done
