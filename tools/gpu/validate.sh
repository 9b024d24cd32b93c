#!/bin/bash
# what the driver runs at round end: pytest -m gpu, smoke(), default bench (ours + reference arm)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/validate_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/validate_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/validate_bench.json 2>gpurun_out/validate_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/validate_ref.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/validate_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/validate_ref.json').read().strip().splitlines()[-1])
print("ours", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], "cpu", j["cpu_baseline"]["value"], "| reference arm", r["value"], r["unit"])
PY
