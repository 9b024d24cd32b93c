#!/bin/bash
# GPU trip: GEMM + engine parity tests, then the bench with per-op dump (A/B of a pw_gemm.cu build variant)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${TAG:-a}
timeout 600 python -m pytest -m gpu -q -x --timeout 300 -p no:cacheprovider tests/test_gpu_ops.py tests/test_gpu_engine.py > gpurun_out/t_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/t_gpu_$TAG.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_$TAG.json > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1])
r=json.load(open('gpurun_out/ops_$TAG.json'))
agg={}
for x in r: agg[x['kind']]=agg.get(x['kind'],0)+x['ms']
print("$TAG", j['value'], j['e2e']['value'], {k:round(v,3) for k,v in agg.items()})
PY
done
