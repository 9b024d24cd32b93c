#!/bin/bash
# GPU trip: parity tests, then the bench with the integer requantisation (default) and with the guarded fp32 one (bit 7)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest -m gpu -q -x --timeout 600 -p no:cacheprovider tests ${PYTEST_ARGS} > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_int.json > gpurun_out/bench_int.log 2>gpurun_out/bench_int.err
echo "bench int rc=$?"; tail -c 2500 gpurun_out/bench_int.log; tail -n 5 gpurun_out/bench_int.err
CODENET_DEBUG_FLAGS=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_fp.json > gpurun_out/bench_fp.log 2>gpurun_out/bench_fp.err
echo "bench fp rc=$?"; tail -c 1200 gpurun_out/bench_fp.log; tail -n 5 gpurun_out/bench_fp.err
python - <<'PY'
import json
for tag in ("int", "fp"):
    r=json.load(open('gpurun_out/ops_%s.json' % tag))
    agg={}
    for x in r: agg.setdefault(x['kind'],[0,0]); agg[x['kind']][0]+=x['ms']; agg[x['kind']][1]+=x.get('MB',0) or 0
    print(tag, {k:(round(v[0],3), round(v[1]/max(v[0],1e-9)/1e3,1)) for k,v in agg.items()})
PY
