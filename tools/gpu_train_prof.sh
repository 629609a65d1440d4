#!/bin/bash
mkdir -p gpurun_out
python tools/train_bench.py --config pretrain --len 2048 --steps 20 --warmup 5 > gpurun_out/train_plain.log 2>&1; tail -1 gpurun_out/train_plain.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 190 --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py --config pretrain --len 2048 --steps 8 --warmup 5 > gpurun_out/ncu_train.log 2>&1
echo "rc $?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/train_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
idx={h:i for i,h in enumerate(rows[hdr])}
tot=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    if len(r)<=idx['Metric Value'] or r[idx['Metric Name']]!='gpu__time_duration.sum': continue
    v=float(r[idx['Metric Value']].replace(',','')); u=r[idx['Metric Unit']]
    us=v/1000 if u.startswith('n') else (v if u.startswith('u') else v*1000)
    k=r[idx['Kernel Name']][:70]; tot[k][0]+=1; tot[k][1]+=us
s=sum(v[1] for v in tot.values())
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][1])[:28]:
    print("%-72s %5d %10.1f %6.3f"%(k,v[0],v[1],v[1]/s))
print("total us", s)
PY
