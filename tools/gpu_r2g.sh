#!/bin/bash
for i in 1 2 3; do python tools/train_bench.py --config finetune --steps 50 --warmup 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('finetune ms', round(d['ms_per_step'],3), 'launches', d['gpu_launches_per_step'], 'loss', d['final_loss'])"; done
python tools/train_bench.py --config pretrain --len 2048 --steps 30 --warmup 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pretrain ms', round(d['ms_per_step'],3), 'loss', d['final_loss'])"
python tools/train_bench.py --config pretrain --len 2048 --steps 30 --warmup 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pretrain ms', round(d['ms_per_step'],3), 'loss', d['final_loss'])"
