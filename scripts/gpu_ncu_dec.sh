#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_dec.csv python scripts/ncu_target.py decode > gpurun_out/ncu_dec.log 2>&1; echo rc=$?
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_dec.csv')) if len(r)>5]
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
    v*={'ns':1,'nsecond':1,'us':1e3,'usecond':1e3,'ms':1e6,'msecond':1e6}.get(u,1)
    k=d['Kernel Name'][:70]+' grid='+d.get('Grid Size','')
    agg[k][0]+=1; agg[k][1]+=v
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
    print(f"{v[1]/1e3:9.1f} us total  {v[1]/v[0]/1e3:8.1f} us/launch  n={v[0]:3d}  {k}")
PY
