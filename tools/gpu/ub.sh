#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
./tools/ubench_consts > gpurun_out/ubench_consts.txt 2>&1; cat gpurun_out/ubench_consts.txt
timeout 1200 python -m pytest -m gpu -q -x --timeout 600 -p no:cacheprovider tests/test_gpu_ops.py > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_int.json > gpurun_out/bench_int.log 2>gpurun_out/bench_int.err
echo "bench int rc=$?"; tail -n 5 gpurun_out/bench_int.err
CODENET_DEBUG_FLAGS=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_fp.json > gpurun_out/bench_fp.log 2>gpurun_out/bench_fp.err
echo "bench fp rc=$?"; tail -n 5 gpurun_out/bench_fp.err
python - <<'PY'
import json
for tag in ("int", "fp"):
    r=json.load(open('gpurun_out/ops_%s.json' % tag))
    agg={}
    for x in r: agg.setdefault(x['kind'],[0,0]); agg[x['kind']][0]+=x['ms']; agg[x['kind']][1]+=x.get('MB',0) or 0
    print(tag, {k:(round(v[0],3), round(v[1]/max(v[0],1e-9)/1e3,1)) for k,v in agg.items()})
    print(json.loads(open('gpurun_out/bench_%s.log' % tag).read().strip().splitlines()[-1])["value"])
PY
