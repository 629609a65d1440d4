#!/bin/bash
# 8 GPUs: the driver's scaling launch at N=8 (and N=4 on the same box), full bench incl. e2e legs, probe and train record
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi topo -m | head -12; } > gpurun_out/host8.txt 2>&1
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_$n.log 2>&1
  echo "N=$n exit $?"
  python - <<PY
import json
txt = open("gpurun_out/scale_$n.log").read()
ls = [l for l in txt.splitlines() if '"metric"' in l]
if ls:
    d = json.loads(ls[-1][ls[-1].index("{"):])
    print("N=$n value", round(d["value"]), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), "compact", round(d["e2e_compact_pack"]["value"]), round(d["e2e_compact_pack"]["ms_per_step"],1),
          "probe", [round(x,1) for x in d["h2d_probe"]["gb_per_s_per_gpu"]], "cores/rank", d["config"]["host_cores_per_rank"])
    print("   train", {k: (round(v["ms_per_step"],3), v.get("dp_equals_single_process", {}).get("ok")) for k, v in d["train"].items()})
else:
    print(txt[-1500:])
PY
done
