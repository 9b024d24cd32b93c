#!/bin/bash
# bilinear-offset mode: parity tests that exercise it + one bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest -m gpu -q -x -p no:cacheprovider tests/test_gpu_ops.py tests/test_gpu_engine.py -k "deform or bilinear or reference_vectors" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --offset-mode bilinear > gpurun_out/bench_bilinear.json 2>gpurun_out/bench_bilinear.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_bilinear.json').read().strip().splitlines()[-1])
print("bilinear", j["value"], j["ms_per_step"], j["e2e"]["value"], {k:v["ms"] for k,v in j["roofline"]["families"].items()})
PY
