#!/bin/bash
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_aa.csv \
   python tools/solve_c4.py --max-it 12 > gpurun_out/launches_aa.log 2>&1
tail -2 gpurun_out/launches_aa.log
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_aa.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<15: continue
    name=r[4].split('(')[0]
    agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=float(r[14])/1e6
for k,v in agg.items(): print("%-40s %3d launches  %.3f ms each"%(k,v[0],v[1]/v[0]))
PY
timeout 200 python tools/solve_c4.py --max-it 40 2>&1 | tail -1
