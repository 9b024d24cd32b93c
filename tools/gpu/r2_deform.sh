#!/bin/bash
# deform_tile.cu: op-level parity (every layer + random shapes + the reference's CUDA ext), engine parity, bench A/B against the v3 kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest -m gpu -q -x --timeout 300 -p no:cacheprovider tests -k "deform or reference_ext" > gpurun_out/t_deform.log 2>&1
echo "pytest deform rc=$?"; tail -n 15 gpurun_out/t_deform.log
timeout 900 python -m pytest -m gpu -q -x --timeout 600 -p no:cacheprovider tests/test_gpu_engine.py > gpurun_out/t_engine.log 2>&1
echo "pytest engine rc=$?"; tail -n 8 gpurun_out/t_engine.log
for flags in 0 2048 4096; do
  CODENET_DEBUG_FLAGS=$flags timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-bilinear > gpurun_out/bench_f$flags.log 2>gpurun_out/bench_f$flags.err
  echo "bench flags=$flags rc=$?"; tail -n 3 gpurun_out/bench_f$flags.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/bench_f$flags.log').read().strip().splitlines()[-1])
print("flags $flags value", j["value"], "parity", j["parity_checked"], "deform", j["deform"]["ms"], j["deform"]["frac_of_hbm_peak"], [(l["layer"], l["ms"]) for l in j["deform"]["layers"]])
PY
done
