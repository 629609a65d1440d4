#!/bin/bash
# 2 GPUs: DP tests (flat + bucketed), bench at N=2 with the train sub-record; then N=1 train record
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
run dp_tests  python -m pytest tests/test_dp_train_gpu.py tests/test_train_gpu.py -m gpu -q --maxfail=4
run bench1    python bench.py --steps 6 --warmup 3 --no-cpu-baseline
run bench2    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3
python - <<'PY'
import json
for f in ("bench1", "bench2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["e2e_compact_pack"]["value"], d["h2d_probe"]["gb_per_s_per_gpu"], json.dumps(d["train"])[:1500])
    except Exception as e:
        print(f, "unparsed", e)
PY
