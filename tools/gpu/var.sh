#!/bin/bash
# build variants on the box and bench each: usage scripts_gpu_var.sh "<flags A>" "<flags B>" ...
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
i=0
for fl in "$@"; do
  CDN_NVCC_EXTRA="$fl" python codenet_b200/build.py --force > /dev/null 2>gpurun_out/build_$i.err || { echo "build failed: $fl"; tail -5 gpurun_out/build_$i.err; continue; }
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_var$i.log 2>gpurun_out/bench_var$i.err
  python - "$fl" gpurun_out/bench_var$i.log <<'PY'
import json, sys
j=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "|", j["value"], j["ms_per_step"], {k:v["ms"] for k,v in j["roofline"]["families"].items()})
PY
  i=$((i+1))
done
