#!/bin/bash
# Round 2, first call of the session: host facts, the new loader tests, the whole GPU suite, the restructured bench.
mkdir -p gpurun_out
{ nproc; free -g; df -h /dev/shm /tmp; lscpu | grep -E "Model name|NUMA|Socket|Thread"; nvidia-smi topo -m; } > gpurun_out/host.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/$name.log | cut -c1-600; }
run loader   python -m pytest tests/test_pipeline_gpu.py tests/test_data_pack.py -q --maxfail=4
run bench    python bench.py --steps 10 --warmup 3
run refarm   python bench.py --impl reference --steps 1 --warmup 0
run all_gpu  python -m pytest tests -m gpu -q --maxfail=6
