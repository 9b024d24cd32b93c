#!/bin/bash
# ncu launch list (device time of every launch) of the bench command, cold-cache and serialised: compare SHARES
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-0} -c ${COUNT:-400} --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/ncu_launch.log | cut -c1-300
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10]
hdr=rows[0]; iK=hdr.index("Kernel Name"); iV=hdr.index("Metric Value"); iU=hdr.index("Metric Unit")
agg=collections.OrderedDict(); n=collections.Counter()
for r in rows[1:]:
    k=r[iK].split('(')[0]; v=float(r[iV].replace(',','')); 
    if r[iU]=='ns': v/=1e3
    elif r[iU]=='ms': v*=1e3
    agg[k]=agg.get(k,0)+v; n[k]+=1
tot=sum(agg.values())
for k,v in sorted(agg.items(),key=lambda kv:-kv[1]): print("%-60s n=%4d  %10.1f us  %5.1f%%"%(k[:60],n[k],v,100*v/tot))
PY
