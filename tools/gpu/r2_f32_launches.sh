#!/bin/bash
# launch list of ONE forward + decode of the float path (BASELINE config 5 geometry, batch 32) under ncu: per-kernel totals
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cat > /tmp/f32_one.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from tools.bench_f32_config5 import _cfg, _state
from codenet_b200.engine_f32 import EngineF32
from codenet_b200.synth import make_images
raw, g = _state()
eng = EngineF32(_cfg(), raw, device=0, gemm=os.environ.get("CODENET_F32_GEMM", "tf32x3"))
x = torch.from_numpy(np.concatenate([make_images(8, 512, seed=100)] * 4).copy()).cuda()
eng.detect(x); torch.cuda.synchronize()
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_f32_launches.csv python /tmp/f32_one.py > gpurun_out/f32_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_f32_launches.csv')) if len(r)>10]
hdr=rows[0]; iK=hdr.index("Kernel Name"); iV=hdr.index("Metric Value"); iU=hdr.index("Metric Unit")
agg=collections.OrderedDict(); n=collections.Counter()
for r in rows[1:]:
    k=r[iK].split('(')[0]; v=float(r[iV].replace(',',''))
    if r[iU] in ('ns','nsecond'): v/=1e3
    elif r[iU] in ('ms','msecond'): v*=1e3
    agg[k]=agg.get(k,0)+v; n[k]+=1
tot=sum(agg.values())
print("total %.1f us, %d launches" % (tot, sum(n.values())))
for k,v in sorted(agg.items(),key=lambda kv:-kv[1])[:14]: print("%-50s n=%5d %10.1f us %5.1f%%"%(k[:50],n[k],v,100*v/tot))
PY
