#!/bin/bash
# full GPU test suite + bench line + profiles
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest -m gpu -q -x --timeout 600 -p no:cacheprovider tests > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --dump-ops gpurun_out/ops_final.json > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err
echo "bench rc=$?"; tail -n 3 gpurun_out/bench_final.err
./tools/gpu/r2_profiles2.sh
