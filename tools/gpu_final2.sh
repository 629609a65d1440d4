#!/bin/bash
# Round-end evidence with the shipped build: every GPU test, smoke, both bench arms, fresh ncu of the dominant kernel + launch list.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run final_tests python -m pytest tests -m gpu -q --maxfail=6
run final_smoke python -c "import __graft_entry__ as g; g.smoke()"
run final_ref   python bench.py --impl reference --steps 2 --warmup 1
run final_bench python bench.py --steps 10 --warmup 3
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --e2e-batches 1"
ncu --set full --clock-control none --import-source on -k regex:attn2_tc05 -s 8 -c 1 -o gpurun_out/r02b_attn_full -f $CMD > gpurun_out/ncu_attn_full.log 2>&1
echo "attn ncu rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launches rc $?"
