#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for b in 16 32 64 96 128 192 256; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch $b > gpurun_out/bench_b$b.json 2>gpurun_out/bench_b$b.err
  python - $b <<'PY'
import json, sys
b=sys.argv[1]
j=json.loads(open('gpurun_out/bench_b%s.json'%b).read().strip().splitlines()[-1])
print(b, "ms/step", j["ms_per_step"], "img/s", j["value"], "e2e", j["e2e"]["value"])
PY
done
