#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for f in 0 268435456; do
CODENET_DEBUG_FLAGS=$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-bilinear --dump-ops gpurun_out/ops_res_$f.json > gpurun_out/bench_res_$f.log 2>gpurun_out/bench_res_$f.err
python - $f <<'PY'
import json, sys
f=sys.argv[1]
r=json.load(open('gpurun_out/ops_res_%s.json'%f))
print(f, {x['op']: x['ms'] for x in r if x['op'] in ('layer1.0.pw1','layer1.1.pw1','layer2.1.pw1')})
PY
done
