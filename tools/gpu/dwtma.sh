#!/bin/bash
# GPU trip for dw_tma.cu: op-level + engine parity tests, then the bench with the TMA-staged kernel and with the LDG kernel (bit 9)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest -m gpu -q -x --timeout 200 -p no:cacheprovider tests/test_gpu_ops.py tests/test_gpu_engine.py > gpurun_out/t_dwtma.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -n 12 gpurun_out/t_dwtma.log
[ $rc -ne 0 ] && exit 1
for tag in tma s2ldg; do
  flags=0; [ $tag = s2ldg ] && flags=1024
  CODENET_DEBUG_FLAGS=$flags timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_$tag.json > gpurun_out/bench_$tag.log 2>gpurun_out/bench_$tag.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1])
r=json.load(open('gpurun_out/ops_$tag.json'))
print("$tag", j['value'], j['e2e']['value'], {k:v['ms'] for k,v in j['roofline']['families'].items()})
print("   ", [(x['op'],x['ms']) for x in r if x['kind']=='dw'])
PY
done
