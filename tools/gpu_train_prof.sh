#!/bin/bash
# launch list of one training step at csm-1b / 4096 tokens: per-kernel time shares
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2100 --csv --log-file gpurun_out/train_launches.csv \
   python tools/train_bench.py --seq 4096 --batch 1 --steps 1 --warmup 1 > gpurun_out/train_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/train_launches.csv')))
i0=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[i0]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
data=rows[i0+1:]
half=len(data)//2
tot=collections.Counter(); cnt=collections.Counter()
for r in data[half:]:
    v=float(r[mv].replace(',','')); 
    if r[mu]=='ns': v/=1e3
    elif r[mu]=='ms': v*=1e3
    name=r[kn].split('(')[0][-60:]
    tot[name]+=v; cnt[name]+=1
s=sum(tot.values())
print('launches in the second step', len(data)-half, 'total us', round(s))
for k,v in tot.most_common(25): print(f'{v:10.0f} us {100*v/s:5.1f}%  n={cnt[k]:4d}  {k}')
PY
