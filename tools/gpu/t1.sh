#!/bin/bash
# parity tests (optionally a subset: PYTEST_ARGS) + one bench line with the per-family split
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest -m gpu -q -x --timeout 600 -p no:cacheprovider tests ${PYTEST_ARGS} > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_int.json > gpurun_out/bench_int.log 2>gpurun_out/bench_int.err
echo "bench rc=$?"; tail -n 5 gpurun_out/bench_int.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_int.log').read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["e2e"]["value"], j.get("requant"), {k:v["ms"] for k,v in j["roofline"]["families"].items()}, [(l["layer"], l["ms"], l["GBps"]) for l in j["deform"]["layers"]])
PY
