#!/bin/bash
mkdir -p gpurun_out
for flags in "" "--no-pin"; do
for n in 8 1; do
  if [ $n = 1 ]; then CMD="python bench.py --only-train $flags"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --only-train $flags"; fi
  timeout 600 $CMD 2>/dev/null | grep '"train"' | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=$n flags=[$flags] cores/rank', d['host_cores_per_rank'], {k: (round(v['ms_per_step'], 3), v.get('dp_equals_single_process', {}).get('ok')) for k, v in d['train'].items()})"
done
done
python -m pytest tests/test_dp_train_gpu.py -m gpu -q 2>&1 | tail -1
