#!/bin/bash
# 2 GPUs: DP tests + N=1/N=2 train records; ncu --set full of the evaluation kernels and of the fused FFN; fresh attention capture
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/$name.log | cut -c1-400; }
run dp_tests  python -m pytest tests/test_dp_train_gpu.py tests/test_train_gpu.py tests/test_pipeline_gpu.py -m gpu -q --maxfail=4
run bench2    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3
export CUDA_VISIBLE_DEVICES=0
run eval_plain python tools/eval_timing.py
ncu --set full --clock-control none --import-source on -k regex:"shot_mean|knapsack|summary_mask|overlap|fscore" -s 30 -c 9 -o gpurun_out/r02_eval_full -f python tools/eval_timing.py > gpurun_out/ncu_eval.log 2>&1
echo "ncu eval rc $?"
run ffn_plain python tools/ffn_timing.py
ncu --set full --clock-control none --import-source on -k regex:ffn_tc05 -s 3 -c 1 -o gpurun_out/r02_ffn_full -f python tools/ffn_timing.py > gpurun_out/ncu_ffn.log 2>&1
echo "ncu ffn rc $?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --e2e-batches 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launches rc $?"
python - <<'PY'
import json
txt = open("gpurun_out/bench2.log").read()
line = [l for l in txt.splitlines() if '"metric"' in l][-1]
d = json.loads(line[line.index("{"):])
print("N=2", d["value"], d["e2e"]["value"], d["e2e_compact_pack"]["value"], json.dumps(d["train"])[:1600])
PY
