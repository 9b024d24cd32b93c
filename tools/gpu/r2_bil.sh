#!/bin/bash
# bilinear tile kernel: parity (op tests in both modes, engine, reference ext, detector) + default bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest -m gpu -q -x --timeout 900 -p no:cacheprovider tests > gpurun_out/t_bil.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/t_bil.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_bil.log 2>gpurun_out/bench_bil.err
echo "bench rc=$?"; tail -n 3 gpurun_out/bench_bil.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_bil.log').read().strip().splitlines()[-1])
print("value", j["value"], "parity", j["parity_checked"], "e2e", j["e2e"]["value"], "deform", j["deform"]["ms"], j["deform"]["frac_of_hbm_peak"])
print("bilinear", j["reference_semantics"])
PY
timeout 600 python bench.py --config e --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_e.log 2>gpurun_out/bench_e.err; echo "e rc=$?"; tail -n 4 gpurun_out/bench_e.err
python - <<PY
import json
j=json.loads(open('gpurun_out/bench_e.log').read().strip().splitlines()[-1])
print("config e value", j["value"], "parity", j["parity_checked"], j["parity"], "e2e", j["e2e"]["value"], {k:v["ms"] for k,v in j["roofline"]["families"].items()})
PY
