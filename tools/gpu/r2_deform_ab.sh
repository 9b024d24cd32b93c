#!/bin/bash
# deform_tile.cu variants: parity of every variant (op tests + engine) and bench A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for flags in 0 4096 8192; do
  CODENET_DEBUG_FLAGS=$flags timeout 600 python -m pytest -m gpu -q -x --timeout 300 -p no:cacheprovider tests -k "deform or reference_ext or reference_vectors_512 or w2_maxpool" > gpurun_out/t_deform_f$flags.log 2>&1
  echo "pytest flags=$flags rc=$?"; tail -n 4 gpurun_out/t_deform_f$flags.log
done
for flags in 0 4096 8192 16384 20480 24576; do
  CODENET_DEBUG_FLAGS=$flags timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-bilinear > gpurun_out/bench_f$flags.log 2>gpurun_out/bench_f$flags.err
  echo "bench flags=$flags rc=$?"; tail -n 3 gpurun_out/bench_f$flags.err
  python - <<PY
import json
j=json.loads(open('gpurun_out/bench_f$flags.log').read().strip().splitlines()[-1])
print("flags $flags value", j["value"], "parity", j["parity_checked"], "deform", j["deform"]["ms"], j["deform"]["frac_of_hbm_peak"], [(l["layer"], l["ms"]) for l in j["deform"]["layers"]])
PY
done
