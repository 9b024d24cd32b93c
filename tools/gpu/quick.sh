#!/bin/bash
# quick GPU trip: parity tests + bench (+ optional per-op dump)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x --timeout 600 -p no:cacheprovider tests ${PYTEST_ARGS} > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --dump-ops gpurun_out/ops.json > gpurun_out/bench.log 2>gpurun_out/bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/ops.json'))
agg={}
for x in r: agg.setdefault(x['kind'],[0,0]); agg[x['kind']][0]+=x['ms']; agg[x['kind']][1]+=x.get('MB',0) or 0
for k,v in agg.items(): print(k, round(v[0],3),'ms', round(v[1]/max(v[0],1e-9)/1e3,1),'GB/s')
for x in r:
    if x['kind']=='pw': print(x['op'],x['ms'],x.get('GBps'))
PY
