#!/bin/bash
mkdir -p gpurun_out
for f in ${FLAGS:-0 32}; do
  CODENET_DEBUG_FLAGS=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_$f.json > gpurun_out/bench_$f.log 2>&1
  python - <<PY
import json
r=json.load(open('gpurun_out/ops_$f.json'))
pw=[x for x in r if x['kind']=='pw']
print("flags=$f pw total %.3f"%sum(x['ms'] for x in pw), " ".join("%s=%.3f"%(x['op'],x['ms']) for x in pw if x['op'] in ('layer1.0.pw1','layer1.1.pw1','layer1.1.pw3','layer2.1.pw1','layer2.1.pw3','layer3.1.pw3','layer4','heads.pw1','heads.out')))
PY
done
