#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/f32_launches.csv python tools/bench_f32.py > gpurun_out/f32_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/f32_launches.csv')) if len(r)>10]
hdr=rows[0]; iK=hdr.index("Kernel Name"); iV=hdr.index("Metric Value"); iU=hdr.index("Metric Unit")
agg=collections.OrderedDict(); n=collections.Counter()
for r in rows[1:]:
    k=r[iK].split('(')[0]; v=float(r[iV].replace(',',''))
    if r[iU] in ('ns','nsecond'): v/=1e3
    elif r[iU] in ('ms','msecond'): v*=1e3
    agg[k]=agg.get(k,0)+v; n[k]+=1
tot=sum(agg.values())
for k,v in sorted(agg.items(),key=lambda kv:-kv[1])[:12]: print("%-50s n=%5d %10.1f us %5.1f%%"%(k[:50],n[k],v,100*v/tot))
PY
