#!/bin/bash
# GPU trip for heads_fused.cu: its own parity test first (short timeout: a hang must not hold the box), then the suite + bench A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 180 python -m pytest -m gpu -q -x --timeout 120 -p no:cacheprovider tests/test_gpu_engine.py -k heads_fused > gpurun_out/t_hf.log 2>&1
rc=$?; echo "hf pytest rc=$rc"; tail -n 25 gpurun_out/t_hf.log
[ $rc -ne 0 ] && exit 1
timeout 900 python -m pytest -m gpu -q -x --timeout 300 -p no:cacheprovider tests/test_gpu_engine.py tests/test_gpu_compat.py > gpurun_out/t_gpu_hf.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/t_gpu_hf.log
for tag in fused nofuse; do
  extra=""; [ $tag = nofuse ] && extra="--no-fuse-heads"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu $extra --dump-ops gpurun_out/ops_$tag.json > gpurun_out/bench_$tag.log 2>gpurun_out/bench_$tag.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/bench_$tag.log').read().strip().splitlines()[-1])
r=json.load(open('gpurun_out/ops_$tag.json'))
print("$tag", j['value'], j['e2e']['value'], {k:v['ms'] for k,v in j['roofline']['families'].items()}, [(x['op'],x['ms'],x['GBps']) for x in r if x['op'].startswith('heads')])
PY
done
